"""Data-parallel helpers for the hot path (one process per GPU, torch.distributed).

The path shards naturally: batch columns (VAE / IWAE / VIMCO / BNN-VI) or chains
(SG-MCMC) are split over ranks, every particle-axis reduction stays inside a rank, and the only
exchange is a sum of the scalar objective (plus the usual parameter-gradient all-reduce of the user's
networks).  The reference has no distributed code at all; nothing here changes its API.
"""
import torch
import torch.distributed as dist

from . import _rng

# Philox offsets of different ranks are separated by this many ticks so their noise never overlaps
RANK_OFFSET_STRIDE = 1 << 40


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def world():
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous, balanced [start, stop) of `n` batch columns (or chains) owned by `rank`."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(int(n), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def decorrelate_rng(rank=None):
    """Give this rank its own Philox offset range (same seed on every rank, disjoint streams)."""
    r, _ = world()
    _rng.rank_stride = (r if rank is None else rank) * RANK_OFFSET_STRIDE


def global_mean_objective(local_mean_loss, n_local, n_global, group=None, async_op=False):
    """Mean objective over the GLOBAL batch from each rank's mean over its local columns:
    sum_r (loss_r * n_r) / n_global.  One all-reduce of a single scalar."""
    t = (local_mean_loss.detach() * (float(n_local) / float(n_global))).reshape(1).clone()
    if not is_initialized():
        return t.reshape(())
    work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (t.reshape(()), work) if async_op else t.reshape(())


def all_reduce_gradients(params, n_local, n_global, group=None):
    """Turn gradients of the LOCAL mean loss into gradients of the GLOBAL mean loss:
    g <- sum_r g_r * n_r / n_global, flattened into one all-reduce."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    scale = float(n_local) / float(n_global)
    if not is_initialized():
        for g in grads:
            g.mul_(scale)
        return
    flat = torch.cat([g.reshape(-1) for g in grads]) * scale
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].reshape(g.shape))
        off += n
