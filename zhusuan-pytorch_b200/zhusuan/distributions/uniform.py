"""Univariate Uniform — drop-in for zhusuan/distributions/uniform.py of the reference.

sample  : u ~ U[0,1) from in-kernel Philox (the reference draws torch.distributions.Uniform(0, 1).sample() on the CPU,
          :63-66), returned as u * (high - low) + low (:68);
log_prob: -log(high - low) on [low, high), -inf outside, with the event-axis sum fused
          (torch.distributions.Uniform.log_prob, reference :70-81).
The reference's observable quirks are kept on purpose (pinned by tests/golden/uniform.npz):
  * a reparameterised draw caches the UNSCALED u as `sample_cache` (:67), so `log_prob(None)` right after a draw is
    evaluated at u, not at the returned sample;
  * a non-reparameterised draw is scaled twice: Uniform(low, high).sample() is cached and then returned as
    sample * (high - low) + low (:60-62,67-68);
  * `log_prob` builds torch.distributions.Uniform(low, high) with argument validation: ValueError unless low < high
    everywhere, ValueError for a value outside [low, high] (test_uniform.py:74-76 of the reference relies on the first).
"""
import torch

from zhusuan._shapes import broadcast_shapes as _bshapes
from zhusuan.distributions.base import Distribution, DEFAULT_DEVICE, resolve_device
from zhusuan.distributions.utils import assert_same_log_float_dtype, check_broadcast
from zhusuan import _ops, _backend as _be

__all__ = ['Uniform']


class Uniform(Distribution):
    """Uniform(low, high): `low` inclusive, `high` exclusive."""

    def __init__(self, low, high, dtype=None, is_continuous=True, is_reparameterized=True, group_ndims=0,
                 device=DEFAULT_DEVICE, **kwargs):
        device = resolve_device(device, low, high)
        self._low = torch.as_tensor(low, dtype=dtype).to(device)
        self._high = torch.as_tensor(high, dtype=dtype).to(device)
        check_broadcast(self._low, self._high)
        dtype = assert_same_log_float_dtype([(self._low, "Uniform.low"), (self._high, "Uniform.high")])
        super(Uniform, self).__init__(dtype, is_continuous, is_reparameterized, group_ndims=group_ndims, device=device,
                                      **kwargs)

    @property
    def low(self):
        return self._low

    @property
    def high(self):
        return self._high

    def _batch_shape(self):
        return _bshapes(self._low.shape, self._high.shape)

    def _unit(self, n_samples):
        """u ~ U[0,1) of shape ([n] +) low.shape -- the reference's sample has the shape of `low` (:54-62)."""
        zero, one = torch.zeros_like(self._low), torch.ones_like(self._low)
        return _ops.locscale_sample(_be.FAM_UNIFORM, zero, one, n_samples, False)

    def _sample(self, n_samples=1, **kwargs):
        span = self._high - self._low
        u = self._unit(n_samples)
        if self.is_reparameterized:
            self.sample_cache = u                      # the unscaled draw (:67)
            return u * span + self._low                # :68
        s = (u * span + self._low).detach()            # Uniform(low, high).sample() (:61)
        self.sample_cache = s
        return s * span + self._low                    # scaled again (:68)

    def _validate(self, given):
        if not bool(torch.lt(self._low, self._high).all()):
            raise ValueError("Expected parameter low < high of distribution Uniform (torch.distributions validation)")
        if not bool((torch.ge(given, self._low) & torch.le(given, self._high)).all()):
            raise ValueError("Expected value argument to be within the support [low, high] of distribution Uniform")

    def _log_prob_event(self, given, n_event):
        given = self._given(given)
        self._validate(given)
        return _ops.locscale_log_prob(_be.FAM_UNIFORM, given, self._low, self._high, n_event)

    def _log_prob(self, sample=None, **kwargs):
        return self._log_prob_event(sample, 0)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
