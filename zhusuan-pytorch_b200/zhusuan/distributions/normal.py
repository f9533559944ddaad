"""Univariate Normal — drop-in for zhusuan/distributions/normal.py of the reference.

sample  : one kernel (Philox + Box-Muller + mean + std*eps), parameters broadcast over particles
          instead of `.repeat`ed (reference :89-107 draws eps on the CPU and copies it over);
log_prob: one kernel for the log-density and its event-axis sum (reference :109-126 + base.py:175-176).
"""
import torch
from zhusuan._shapes import broadcast_shapes as _bshapes

from zhusuan.distributions.base import Distribution, DEFAULT_DEVICE, resolve_device
from zhusuan.distributions.utils import assert_same_log_float_dtype, check_broadcast
from zhusuan import _ops

__all__ = ['Normal']


class Normal(Distribution):
    """Normal(mean, std | logstd).  Exactly one of `std` / `logstd` must be given."""

    def __init__(self, mean=0., std=None, logstd=None, dtype=None, is_continuous=True, is_reparameterized=True,
                 group_ndims=0, device=DEFAULT_DEVICE, **kwargs):
        device = resolve_device(device, mean, std, logstd)
        self._ctor_args = (mean, std, logstd)  # as given: lets _is_standard() decide on the host
        self._mean = torch.as_tensor(mean, dtype=dtype).to(device)
        if (logstd is None) == (std is None):
            raise ValueError("Either `std` or `logstd` should be passed. It is not allowed "
                             "that both are specified or both are not.")
        if std is None:
            # std = exp(logstd) is what the reference stores (:56); log_prob takes its log again (:121)
            self._std = torch.exp(torch.as_tensor(logstd, dtype=dtype)).to(device)
        else:
            self._std = torch.as_tensor(std, dtype=dtype).to(device)
        check_broadcast(self._std, self._mean)
        dtype = assert_same_log_float_dtype([(self._mean, "Normal.mean"), (self._std, "Normal.std")])
        super(Normal, self).__init__(dtype=dtype, is_continuous=is_continuous,
                                     is_reparameterized=is_reparameterized, group_ndims=group_ndims,
                                     device=device, **kwargs)

    @property
    def mean(self):
        return self._mean

    @property
    def std(self):
        return self._std

    @property
    def logstd(self):
        return torch.log(self._std)

    def _batch_shape(self):
        return _bshapes(self._mean.shape, self._std.shape)

    def sample_device(self):
        return self._mean.device

    def _sample(self, n_samples=1):
        z = _ops.normal_sample(self._mean, self._std, n_samples, self.is_reparameterized)
        self.sample_cache = z
        return z

    def _sample_fused(self, n_samples, n_event):
        if not _ops.latent_supported(tuple(self._mean.shape), n_event, self._dtype, self._mean, self._std):
            return None
        return _ops.normal_sample_logq(self._mean, self._std, n_samples, self.is_reparameterized, n_event)

    def _is_standard(self):
        """True when this is Normal(0, 1) in every element and that is known WITHOUT synchronising with the device
        (Python numbers, host tensors, or CUDA tensors already checked once): _ops.is_constant."""
        mean, std, logstd = self._ctor_args
        if not _ops.is_constant(mean, 0.0):
            return False
        return _ops.is_constant(logstd, 0.0) if std is None else _ops.is_constant(std, 1.0)

    def _log_prob_event(self, given, n_event):
        given = self._given(given)
        # a sample drawn by a fused latent launch carries its standard-Normal log-density: a standard prior node
        # (the generator's p(z) in the VAE / IWAE examples, iwae.py:60-72) needs no launch of its own
        lp = _ops.std_prior_logp(given, "normal", n_event)
        if (lp is not None and lp.dtype == self._dtype and self._is_standard()
                and tuple(_bshapes(tuple(given.shape), tuple(self._batch_shape()))) == tuple(given.shape)):
            return _ops.back_home(lp, given.device)
        return _ops.normal_log_prob(given, self._mean, self._std, n_event)

    def _log_prob(self, sample=None):
        return _ops.normal_log_prob(self._given(sample), self._mean, self._std, 0)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
