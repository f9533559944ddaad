"""Univariate Laplace — drop-in for zhusuan/distributions/laplace.py of the reference.

sample  : loc - scale * sign(u) * log1p(-|u|), u ~ U(-1, 1) from in-kernel Philox — the transform of
          torch.distributions.Laplace.sample, which the reference calls (:60-76); never reparameterised (:31);
log_prob: -log(2 scale) - |x - loc| / scale with the event-axis sum fused (reference :78-92).
"""
import torch

from zhusuan._shapes import broadcast_shapes as _bshapes
from zhusuan.distributions.base import Distribution, DEFAULT_DEVICE, resolve_device
from zhusuan.distributions.utils import assert_same_log_float_dtype, check_broadcast
from zhusuan import _ops, _backend as _be

__all__ = ['Laplace']


class Laplace(Distribution):
    def __init__(self, loc, scale, dtype=None, is_continuous=True, group_ndims=0, device=DEFAULT_DEVICE, **kwargs):
        device = resolve_device(device, loc, scale)
        self._loc = torch.as_tensor(loc, dtype=dtype).to(device)
        self._scale = torch.as_tensor(scale, dtype=dtype).to(device)
        check_broadcast(self._loc, self._scale)
        dtype = assert_same_log_float_dtype([(self._loc, "Laplace.loc"), (self._scale, "Laplace.scale")])
        super(Laplace, self).__init__(dtype, is_continuous, is_reparameterized=False, group_ndims=group_ndims,
                                      device=device, **kwargs)

    @property
    def loc(self):
        return self._loc

    @property
    def scale(self):
        return self._scale

    def _batch_shape(self):
        return _bshapes(self._loc.shape, self._scale.shape)

    def _sample(self, n_samples=1):
        z = _ops.locscale_sample(_be.FAM_LAPLACE, self._loc, self._scale, n_samples, False)
        self.sample_cache = z
        return z

    def _log_prob_event(self, given, n_event):
        return _ops.locscale_log_prob(_be.FAM_LAPLACE, self._given(given), self._loc, self._scale, n_event)

    def _log_prob(self, sample=None):
        return self._log_prob_event(sample, 0)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
