"""Lazily materialised first draw of a stochastic node.

The reference's protocol draws every latent TWICE per objective step: `stochastic_node` returns draw
#1 to the user's `forward` (`framework/bn.py:158`), which in every example discards it, and the objective then reads
`.tensor` again -- draw #2, the one both nets see (`variational/elbo.py:122`,
`importance_weighted_objective.py:85`).  Draw #1 costs a full sampling launch (and, for a host-resident model, an
8 MB device->host copy) for a tensor nobody looks at.

Inside `lazy_first_draws()` (the objectives and samplers enter it around `variational(observed)`) the value
`stochastic_node` returns is a `LazyDraw`: a `torch.Tensor` subclass with the sample's shape / dtype / device and no
storage.  The first torch operation that touches it runs the node's sampling launch and continues on the real
tensor, so a `forward` that does use its draw (a hierarchical net) sees exactly the reference's semantics; a
`forward` that ignores it never pays for it.  With injected noise (parity tests) draws stay eager so the injected
tensors are consumed in the reference's order.
"""
import contextlib

import torch
from torch.utils._pytree import tree_map

_active = [0]


@contextlib.contextmanager
def lazy_first_draws():
    _active[0] += 1
    try:
        yield
    finally:
        _active[0] -= 1


def lazy_active():
    return _active[0] > 0


def _meta_getters():
    T = torch.Tensor
    out = set()
    for name in ("shape", "dtype", "device", "ndim", "is_cuda", "layout"):
        prop = getattr(T, name, None)
        g = getattr(prop, "__get__", None)
        if g is not None:
            out.add(g)
    for name in ("size", "dim", "numel", "nelement", "ndimension", "element_size", "is_floating_point"):
        out.add(getattr(T, name))
    return out


_META = None


class LazyDraw(torch.Tensor):
    """A sample that has not been drawn yet (see the module docstring)."""

    @staticmethod
    def __new__(cls, materialize, shape, dtype, device):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(int(v) for v in shape), dtype=dtype, device=device,
                                                requires_grad=False)
        t._zs_make = materialize
        t._zs_real = None
        return t

    def _zs_materialize(self):
        if self._zs_real is None:
            make, self._zs_make = self._zs_make, None
            self._zs_real = make(self)
        return self._zs_real

    @property
    def materialized(self):
        return self._zs_real is not None

    def __repr__(self):
        if self._zs_real is None:
            with torch._C.DisableTorchFunctionSubclass():
                return "LazyDraw(shape=%s, dtype=%s, device=%s, not drawn)" % (
                    tuple(torch.Tensor.size(self)), torch.Tensor.dtype.__get__(self), torch.Tensor.device.__get__(self))
        return repr(self._zs_real)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        global _META
        if _META is None:
            _META = _meta_getters()
        kwargs = kwargs or {}
        if func in _META and len(args) >= 1 and isinstance(args[0], LazyDraw) and args[0]._zs_real is None:
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)

        return func(*tree_map(_real, args), **tree_map(_real, kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        # anything that reaches the dispatcher without passing __torch_function__ (C++ callers): draw, then go on
        return func(*tree_map(_real, args), **tree_map(_real, kwargs or {}))


def _real(a):
    return a._zs_materialize() if isinstance(a, LazyDraw) else a
