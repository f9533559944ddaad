"""zhusuan.utils — drop-in for zhusuan/utils.py of the reference."""
import torch

from . import _ops

__all__ = ['log_mean_exp']


def log_mean_exp(x, dim=None, keepdims=False):
    """Numerically stable log(mean(exp(x))) along `dim` (reference: zhusuan/utils.py:6-21),
    one kernel launch instead of six.  `dim=None` reduces over all elements."""
    x = torch.as_tensor(x)
    if dim is None:
        out = _ops.log_mean_exp(x.reshape(-1), 0, False)
        return out.reshape([1] * x.dim()) if keepdims else out
    if isinstance(dim, (list, tuple)):
        if len(dim) != 1:
            raise TypeError("log_mean_exp: a single reduction axis is supported")
        dim = dim[0]
    return _ops.log_mean_exp(x, int(dim), bool(keepdims))
