"""Philox stream bookkeeping and noise injection for the sampling kernels.

The reference draws noise with torch's CPU generator and copies it to the device
(normal.py:104, bernoulli.py:79, SGLD.py:51, SGHMC.py:27-34).  Here every sampling kernel draws from
Philox4x32-10 keyed by the CUDA generator's seed (so `torch.manual_seed` governs reproducibility).

Stream position.  By default the position lives ON THE DEVICE (a 16-byte `zs_rng_state` per device, see
include/zs_b200.h): a sampling launch reads it and advances it by one tick (4) when all of its CTAs have read it.
Nothing about the position is baked into the launch arguments, so a training step captured in a CUDA graph draws
fresh noise on every replay.  The state is (re)initialised from the generator whenever the generator's seed or offset
is not the one this module left behind (torch.manual_seed, or torch's own CUDA random ops in between): it takes the
generator's current offset as its base and moves the generator past a reserved block, so the two never share a
position and re-seeding reproduces a run.  `DEVICE_STATE = False` restores host-side (seed, offset) arguments, one
generator tick per sampling call.

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

DEVICE_STATE = True
# generator offsets reserved for one initialisation of the device-side state (2^30 sampling launches)
_RESERVE = 1 << 32


class _State(object):
    __slots__ = ("tensor", "seed", "marker")


_states = {}


def next_philox(device):
    """Host-side position: a fresh (seed, offset) for one sampling call on `device`; advances the generator."""
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


def device_state(device):
    """The device-side stream position of `device` (int64[2] CUDA tensor viewed as zs_rng_state) and its seed,
    (re)initialised from the CUDA generator when needed.  Never touches the generator during stream capture."""
    from . import _backend as be
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _states.get(idx)
    capturing = torch.cuda.is_current_stream_capturing()
    if st is not None and capturing:
        return st
    gen = torch.cuda.default_generators[idx]
    seed = int(gen.initial_seed()) & _MASK64
    off = int(gen.get_offset())
    if st is None or st.seed != seed or st.marker != off:
        if capturing:
            raise be.BackendError("the first sampling call on cuda:%d happened inside CUDA graph capture; run one "
                                  "warm-up step before capturing" % idx)
        if st is None:
            st = _State()
            st.tensor = torch.zeros(2, dtype=torch.int64, device=torch.device("cuda", idx))
            _states[idx] = st
        be.rng_state_init(st.tensor, off)
        st.seed = seed
        st.marker = off + _RESERVE
        gen.set_offset(st.marker)
    return st


def draw_args(device):
    """Keyword arguments (seed, offset, rng_state) of the next Philox sampling launch on `device`."""
    if DEVICE_STATE:
        st = device_state(device)
        return dict(seed=st.seed, offset=rank_stride & _MASK64, rng_state=st.tensor)
    seed, offset = next_philox(device)
    return dict(seed=seed, offset=offset, rng_state=None)


def snapshot_buffer(device):
    """A zs_rng_state-sized device buffer for a forward launch to record the position it used (its backward
    regenerates the noise from it)."""
    return torch.empty(2, dtype=torch.int64, device=device)


def take_injected(kind):
    """Pop the next injected tensor of `kind` ('normal' | 'uniform'), or None."""
    q = _injected[kind]
    return q.popleft() if q else None


def has_injected():
    return bool(_injected["normal"]) or bool(_injected["uniform"])


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
