"""Univariate Bernoulli — drop-in for zhusuan/distributions/bernoulli.py of the reference.

log_prob is x*log(p+1e-8) + (1-x)*log(1-p+1e-8) with the event-axis sum fused (one kernel instead of
the reference's ~8 elementwise kernels over the [K,B,X] likelihood tensor, :84-95); samples are
floats in {0,1} drawn in-kernel from Philox uniforms (:72-82).

Bernoulli(logits=...) (reference :47-50: probs = sigmoid(logits), then the same log-pmf): the kernels take the
logits and apply the sigmoid in registers, and the gradient flows to the logits directly, so a decoder that
ends in a Linear layer (no nn.Sigmoid) saves one read + one write of the [K,B,X] tensor in each direction
(SURVEY 8(f)-1).  `.probs` is still available and is computed on first use.
"""
import torch

from zhusuan._shapes import broadcast_shapes as _bshapes
from zhusuan.distributions.base import Distribution, DEFAULT_DEVICE, resolve_device
from zhusuan.distributions.utils import assert_same_log_float_dtype
from zhusuan import _ops

__all__ = ['Bernoulli']


class Bernoulli(Distribution):
    """Bernoulli(logits | probs).  Exactly one must be given; never reparameterised."""

    def __init__(self, logits=None, probs=None, dtype=None, is_continuous=False, group_ndims=0,
                 device=DEFAULT_DEVICE, **kwargs):
        device = resolve_device(device, probs, logits)
        if (logits is None) == (probs is None):
            raise ValueError("Either `probs` or `logits` should be passed. It is not allowed "
                             "that both are specified or both are not.")
        self._ctor_probs = probs  # as given: lets _is_half() decide on the host
        if logits is None:
            self._probs = torch.as_tensor(probs, dtype=dtype).to(device)
            self._logits_cache = None  # log(p/(1-p)) is built lazily: nothing on the hot path reads it
            self._from_logits = None
        else:
            _logits = torch.as_tensor(logits, dtype=dtype)
            assert_same_log_float_dtype([(_logits, "Bernoulli.logits")])
            self._logits_cache = torch.as_tensor(logits).to(device)
            self._from_logits = _logits.to(device)  # what the kernels read; sigmoid happens in registers
            self._probs = None
        dtype = assert_same_log_float_dtype([(self._from_logits if self._probs is None else self._probs,
                                              "Bernoulli.probs")])
        super(Bernoulli, self).__init__(dtype, is_continuous, is_reparameterized=False, group_ndims=group_ndims,
                                        device=device, **kwargs)

    @property
    def probs(self):
        if self._probs is None:
            self._probs = torch.sigmoid(self._from_logits)
        return self._probs

    @property
    def from_logits(self):
        """The logits tensor when the distribution was built from logits (the kernels' operand), else None."""
        return self._from_logits

    @property
    def logits(self):
        if self._logits_cache is None:
            p = self._probs
            self._logits_cache = torch.log(p / (torch.ones_like(p) - p))
        return self._logits_cache

    def _batch_shape(self):
        return (self._from_logits if self._probs is None else self._probs).shape

    def sample_device(self):
        return (self._from_logits if self._probs is None else self._probs).device

    def _sample(self, n_samples=1, **kwargs):
        s = _ops.bernoulli_sample(self.probs, n_samples)
        self.sample_cache = s
        return s

    def _sample_fused(self, n_samples, n_event):
        if self._from_logits is not None:
            return None
        if not _ops.latent_supported(tuple(self._probs.shape), n_event, self._dtype, self._probs):
            return None
        return _ops.bernoulli_sample_logq(self._probs, n_samples, n_event)

    def _is_half(self):
        """True when this is Bernoulli(0.5) in every element, known without a device synchronisation."""
        return self._ctor_probs is not None and _ops.is_constant(self._ctor_probs, 0.5)

    def _log_prob_event(self, given, n_event):
        given = self._given(given)
        lp = _ops.std_prior_logp(given, "bernoulli", n_event)  # see Normal._log_prob_event
        if (lp is not None and lp.dtype == self._dtype and self._is_half()
                and tuple(_bshapes(tuple(given.shape), tuple(self._batch_shape()))) == tuple(given.shape)):
            return _ops.back_home(lp, given.device)
        if self._from_logits is not None:
            return _ops.bernoulli_log_prob(self._given(given), self._from_logits, n_event, logits=True)
        return _ops.bernoulli_log_prob(self._given(given), self._probs, n_event)

    def _log_prob(self, sample=None):
        return self._log_prob_event(sample, 0)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
