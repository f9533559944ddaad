"""Categorical — NOT in the reference (its name_mapping stops at Uniform, zhusuan/framework/bn.py:8-19)
but named by the hot-path specification; parity is therefore unpinned.  The interface follows the
reference's Bernoulli (logits | probs, dtype, group_ndims, device; non-reparameterised); the value is
the class index stored in the distribution's float dtype, batch_shape = logits.shape[:-1].
"""
import torch

from zhusuan.distributions.base import Distribution, DEFAULT_DEVICE, resolve_device
from zhusuan.distributions.utils import assert_same_log_float_dtype
from zhusuan import _ops

__all__ = ['Categorical']


class Categorical(Distribution):
    def __init__(self, logits=None, probs=None, dtype=None, is_continuous=False, group_ndims=0,
                 device=DEFAULT_DEVICE, **kwargs):
        device = resolve_device(device, logits, probs)
        if (logits is None) == (probs is None):
            raise ValueError("Either `probs` or `logits` should be passed. It is not allowed "
                             "that both are specified or both are not.")
        if logits is None:
            p = torch.as_tensor(probs, dtype=dtype).to(device)
            assert_same_log_float_dtype([(p, "Categorical.probs")])
            self._logits = torch.log(p)
        else:
            self._logits = torch.as_tensor(logits, dtype=dtype).to(device)
        if self._logits.dim() < 1:
            raise ValueError("Categorical logits need a trailing class axis")
        dtype = assert_same_log_float_dtype([(self._logits, "Categorical.logits")])
        super(Categorical, self).__init__(dtype, is_continuous, is_reparameterized=False, group_ndims=group_ndims,
                                          device=device, **kwargs)

    @property
    def logits(self):
        return self._logits

    @property
    def probs(self):
        return torch.softmax(self._logits, -1)

    @property
    def n_categories(self):
        return int(self._logits.shape[-1])

    def _batch_shape(self):
        return self._logits.shape[:-1]

    def _sample(self, n_samples=1, **kwargs):
        s = _ops.categorical_sample(self._logits, n_samples)
        self.sample_cache = s
        return s

    def _log_prob(self, sample=None):
        return _ops.categorical_log_prob(self._given(sample), self._logits)

    def _prob(self, given):
        return torch.exp(self._log_prob(given))
