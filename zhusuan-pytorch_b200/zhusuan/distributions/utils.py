"""dtype / broadcast checks of the distribution constructors (host-side only).
Error types and messages follow zhusuan/distributions/utils.py:10-70 of the reference, which its
tests match by regex (test_normal.py:31-35, test_bernoulli.py:29-33)."""
import torch
from zhusuan._shapes import broadcast_shapes as _bshapes

floating_dtypes = (torch.float32, torch.float16, torch.float64)
log_floating_dtypes = (torch.float32, torch.float64)
integer_dtypes = (torch.int32, torch.int16, torch.int64)


def assert_same_dtype_in(tensors_with_name, dtypes=None):
    allowed = set(dtypes) if dtypes else None
    expected = None
    for tensor, name in tensors_with_name:
        if allowed and tensor.dtype not in allowed:
            if len(dtypes) == 1:
                raise TypeError('{}({}) must have dtype {}.'.format(name, tensor.dtype, dtypes[0]))
            raise TypeError('{}({}) must have a dtype in {}.'.format(name, tensor.dtype, dtypes))
        if expected is None:
            expected = tensor.dtype
        elif expected != tensor.dtype:
            t0, n0 = tensors_with_name[0]
            raise TypeError('{}({}) must have the same dtype as {}({}).'.format(name, tensor.dtype, n0, t0.dtype))
    return expected


def assert_same_float_dtype(tensors_with_name):
    return assert_same_dtype_in(tensors_with_name, floating_dtypes)


def assert_same_log_float_dtype(tensors_with_name):
    return assert_same_dtype_in(tensors_with_name, log_floating_dtypes)


def check_broadcast(a, b):
    """Raise RuntimeError (torch's) when the shapes do not broadcast; no kernel is launched."""
    try:
        _bshapes(a.shape, b.shape)
    except RuntimeError:
        raise
