"""Univariate Bernoulli — drop-in for zhusuan/distributions/bernoulli.py of the reference.

log_prob is x*log(p+1e-8) + (1-x)*log(1-p+1e-8) with the event-axis sum fused (one kernel instead of
the reference's ~8 elementwise kernels over the [K,B,X] likelihood tensor, :84-95); samples are
floats in {0,1} drawn in-kernel from Philox uniforms (:72-82).
"""
import torch

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
        if logits is None:
            self._probs = torch.as_tensor(probs, dtype=dtype).to(device)
            self._logits_cache = None  # log(p/(1-p)) is built lazily: nothing on the hot path reads it
        else:
            _logits = torch.as_tensor(logits, dtype=dtype)
            assert_same_log_float_dtype([(_logits, "Bernoulli.logits")])
            self._logits_cache = torch.as_tensor(logits).to(device)
            self._probs = torch.sigmoid(_logits).to(device)
        dtype = assert_same_log_float_dtype([(self._probs, "Bernoulli.probs")])
        super(Bernoulli, self).__init__(dtype, is_continuous, is_reparameterized=False, group_ndims=group_ndims,
                                        device=device, **kwargs)

    @property
    def probs(self):
        return self._probs

    @property
    def logits(self):
        if self._logits_cache is None:
            p = self._probs
            self._logits_cache = torch.log(p / (torch.ones_like(p) - p))
        return self._logits_cache

    def _batch_shape(self):
        return self._probs.shape

    def _sample(self, n_samples=1, **kwargs):
        s = _ops.bernoulli_sample(self._probs, n_samples)
        self.sample_cache = s
        return s

    def _log_prob_event(self, given, n_event):
        return _ops.bernoulli_log_prob(self._given(given), self._probs, n_event)

    def _log_prob(self, sample=None):
        return _ops.bernoulli_log_prob(self._given(sample), self._probs, 0)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
