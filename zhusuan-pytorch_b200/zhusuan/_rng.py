"""Philox (seed, offset) bookkeeping and noise injection for the sampling kernels.

The reference draws noise with torch's CPU generator and copies it to the device
(normal.py:104, bernoulli.py:79, SGLD.py:51, SGHMC.py:27-34).  Here every sampling kernel is a pure
function of (seed, offset): the seed is the CUDA generator's seed (so `torch.manual_seed` governs
reproducibility) and each sampling call consumes one offset tick of that generator.

`inject(...)` replays caller-provided noise instead (parity tests feed the kernels the same tensors
the reference was fed through a patched torch.normal / torch.bernoulli).
"""
import contextlib
from collections import deque

import torch

_MASK64 = (1 << 64) - 1
_injected = {"normal": deque(), "uniform": deque()}
_fallback_offsets = {}
# multi-process runs decorrelate ranks by adding rank * stride to the offset
rank_stride = 0


def next_philox(device):
    """Return a fresh (seed, offset) for one sampling call on `device` and advance the generator."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    gen = torch.cuda.default_generators[idx]
    seed = int(gen.initial_seed()) & _MASK64
    try:
        off = int(gen.get_offset())
        gen.set_offset(off + 4)  # torch requires multiples of 4
    except (RuntimeError, AttributeError):
        off = _fallback_offsets.get((idx, seed), 0)
        _fallback_offsets[(idx, seed)] = off + 4
    return seed, (off + rank_stride) & _MASK64


def take_injected(kind):
    """Pop the next injected tensor of `kind` ('normal' | 'uniform'), or None."""
    q = _injected[kind]
    return q.popleft() if q else None


@contextlib.contextmanager
def inject(normal=(), uniform=()):
    """Within the block, sampling ops consume these tensors (in order) instead of Philox noise.

    normal : standard-normal tensors for Normal samples / already-scaled Gaussian terms for SG-MCMC
    uniform: U[0,1) tensors for Bernoulli / Categorical samples
    """
    saved = {k: deque(v) for k, v in _injected.items()}
    _injected["normal"] = deque(normal)
    _injected["uniform"] = deque(uniform)
    try:
        yield
    finally:
        _injected["normal"], _injected["uniform"] = saved["normal"], saved["uniform"]
