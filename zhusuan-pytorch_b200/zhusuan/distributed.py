"""Data-parallel helpers for the hot path (one process per GPU, torch.distributed over NCCL / NVLink).

The path shards naturally: batch columns (VAE / IWAE / VIMCO / BNN-VI) or chains
(SG-MCMC) are split over ranks, every particle-axis reduction stays inside a rank, and the only
exchange is a SUM of the scalar objective and of the replicated networks' parameter gradients.
The reference has no distributed code at all; nothing here changes its API.

How a data-parallel step is put together:

    bucket = zd.GradientBucket([decoder.parameters(), encoder.parameters()])   # once
    with zd.global_batch(B_global):            # kernels scale by 1/B_global: every rank's loss and gradients are
        loss = objective({"x": x_local})       #   its additive share of the GLOBAL mean -- no rescaling pass
        loss.backward()                        # gradients accumulate straight into the bucket's flat buffer;
    bucket.finish(loss)                        #   each segment's all-reduce is launched (async, NCCL stream) as soon
                                               #   as its last gradient is written, i.e. it overlaps the rest of backward
    loss_global = bucket.loss()                # one SUM all-reduce per segment, no torch.cat, no copy-back

`all_reduce_gradients` keeps the round-1 call signature on top of the same machinery.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from . import _backend as _be
from . import _ops, _rng

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


@contextlib.contextmanager
def global_batch(n_global):
    """Inside the block the objectives average over `n_global` batch columns instead of the local batch: the
    kernels' grad_scale becomes 1/n_global (zs_iw_objective / zs_iw_bernoulli_fused take it as an argument), the
    returned loss is this rank's share sum_b cost_b / n_global.  SUM-all-reducing losses and gradients over ranks
    then gives the global-batch mean and its gradient with no rescaling pass over any gradient."""
    prev = _ops._global_batch[0]
    _ops._global_batch[0] = int(n_global)
    try:
        yield
    finally:
        _ops._global_batch[0] = prev


def global_mean_objective(local_mean_loss, n_local, n_global, group=None, async_op=False):
    """Mean objective over the GLOBAL batch from each rank's mean over its local columns:
    sum_r (loss_r * n_r) / n_global.  One all-reduce of a single scalar."""
    t = (local_mean_loss.detach() * (float(n_local) / float(n_global))).reshape(1).clone()
    if not is_initialized():
        return t.reshape(())
    work = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (t.reshape(()), work) if async_op else t.reshape(())


class GradientBucket(object):
    """Persistent flat gradient buffer of the replicated networks, all-reduced in place.

    `param_groups`: an iterable of parameter iterables, in the order their gradients become ready during backward
    (for a VAE: decoder first, encoder second).  Each group is one contiguous SEGMENT of the flat buffer; the last
    segment carries one extra slot for the scalar objective.  Every parameter's `.grad` is a view into the buffer, so
    autograd accumulates into it directly: there is no flatten (`torch.cat`) before the collective and no copy back
    after it.  A post-accumulate hook on each parameter counts the segment down and launches its all-reduce
    (async: NCCL's own stream, ordered after the compute stream at that point) when the segment is complete, so
    the collective of the decoder's gradients overlaps the latent nodes' and the encoder's backward.
    Without torch.distributed (or with world size 1) everything degenerates to local no-ops.
    """

    def __init__(self, param_groups, device=None, dtype=None, group=None, backend="auto"):
        """backend: "peer" = this library's NVLink peer-memory all-reduce kernel (zs_allreduce_sum_peer; float32 CUDA
        buffers on one node, up to 8 ranks), "nccl" / "gloo" = torch.distributed.all_reduce, "auto" = peer where it
        is available, else torch.distributed.  `self.backend` says which one is in use."""
        groups = [[p for p in g if p.requires_grad] for g in param_groups]
        groups = [g for g in groups if g]
        self.params = [p for g in groups for p in g]
        first = self.params[0] if self.params else None
        self.device = device if device is not None else (first.device if first is not None else torch.device("cpu"))
        self.dtype = dtype if dtype is not None else (first.dtype if first is not None else torch.float32)
        self.group = group
        sizes = [sum(p.numel() for p in g) for g in groups] or [0]
        sizes[-1] += 1  # the objective's slot
        # segments start on 16-byte boundaries (the peer kernel moves float4s; it also keeps every view aligned)
        starts, total = [], 0
        for n in sizes:
            starts.append(total)
            total += (n + 3) // 4 * 4
        self._peer = None
        self.backend = "local"
        if is_initialized() and dist.get_world_size(group) > 1:
            self.backend = dist.get_backend(group)
            if backend in ("auto", "peer"):
                try:
                    self._peer = _PeerBuffer(total, self.device, self.dtype, group)
                    self._peer.tune(total, group)
                    self.backend = "peer"
                except Exception as e:  # no peer access / symmetric memory: torch.distributed does the exchange
                    if backend == "peer":
                        raise
                    self._peer_error = repr(e)
        self.flat = self._peer.flat if self._peer is not None else torch.zeros(total, dtype=self.dtype,
                                                                                  device=self.device)
        self._slot_index = starts[-1] + sizes[-1] - 1  # the objective's slot: last element of the last segment
        self.segments, self._ranges = [], []
        for st, n in zip(starts, sizes):
            self.segments.append(self.flat[st:st + n])
            self._ranges.append((st, (n + 3) // 4 * 4))
        self._comm_stream = None
        self._zero_pending = False
        self._views, self._seg_of, self._pending, self._works = {}, {}, [], []
        for si, g in enumerate(groups):
            off = starts[si]
            for p in g:
                v = self.flat[off:off + p.numel()].view_as(p)
                if p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v
                self._views[id(p)] = v
                self._seg_of[id(p)] = si
                off += p.numel()
                if hasattr(p, "register_post_accumulate_grad_hook"):
                    p.register_post_accumulate_grad_hook(self._on_grad)
                if self.flat.is_cuda:
                    p.register_hook(self._join_zero)
        self._sizes = [len(g) for g in groups] or [0]
        self.zero_grad()

    # -- per step ---------------------------------------------------------------------------------
    def zero_grad(self, overlap=False):
        """Zero every gradient (one memset of the flat buffer) and re-arm the segment counters.  Use this instead
        of optimizer.zero_grad(set_to_none=True), which would detach the parameters from the buffer.
        `overlap=True` (CUDA buffers): the memset runs on the bucket's side stream, ordered after everything the current
        stream has enqueued so far (the optimizer's reads of the gradients); the current stream joins it right before
        the first gradient is accumulated into the buffer (a tensor hook on every parameter) or before the first
        collective of the step, whichever comes first, so zeroing overlaps the forward pass.  Off by default: for a
        5 MB bucket the memset is ~3 us, and inside a captured CUDA graph the fork / join around it measured 4.5 us
        MORE than running it in line (N = 2: 113.8 against 109.3 us per step); it pays for buckets of 100s of MB."""
        self._pending = list(self._sizes)
        self._works = []
        if not (overlap and self.flat.is_cuda):
            self.flat.zero_()
            return
        cur = torch.cuda.current_stream(self.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        self._comm_stream.wait_stream(cur)
        with torch.cuda.stream(self._comm_stream):
            self.flat.zero_()
        self._zero_pending = True

    def _join_zero(self, grad=None):
        """The current stream waits for an overlapped zero_grad() (no host blocking).  Doubles as the parameters'
        tensor hook, which autograd runs BEFORE it accumulates the gradient into the buffer."""
        if self._zero_pending:
            self._zero_pending = False
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        return None

    @property
    def loss_slot(self):
        return self.segments[-1][-1:]

    def _on_grad(self, p):
        v = self._views.get(id(p))
        if v is None:
            return
        if p.grad is not v:  # someone replaced .grad (set_to_none): fold it back into the buffer
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.add_(p.grad)
            p.grad = v
        si = self._seg_of[id(p)]
        self._pending[si] -= 1
        if self._pending[si] == 0 and si < len(self.segments) - 1:
            self._launch(si)

    def _launch(self, si, extra=None):
        self._join_zero()
        if not (is_initialized() and dist.get_world_size(self.group) > 1):
            if extra is not None:
                self.loss_slot.copy_(extra)
            return
        if self._peer is None:
            if extra is not None:
                self.loss_slot.copy_(extra)
            self._works.append(dist.all_reduce(self.segments[si], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        # this library's kernel, on a side stream ordered after what the compute stream has enqueued so far: the
        # exchange overlaps whatever backward launches next; finish() joins the streams
        first, count = self._ranges[si]
        slot = self._slot_index
        # The step's LAST segment has nothing left to overlap with (finish() makes the compute stream wait for it
        # right away), so it is launched in the compute stream itself: a fork / join pair inside a captured graph
        # costs more than it hides.  Earlier segments go to the side stream and overlap the rest of backward.
        if si == len(self.segments) - 1 and os.environ.get("ZS_BUCKET_STREAM") != "side":
            if extra is not None and not (extra.is_cuda and extra.dtype == torch.float32 and extra.is_contiguous()):
                self.loss_slot.copy_(extra)
                extra = None
            self._peer.all_reduce(first, count, flag_set=si % _be.PEER_FLAG_SETS, extra=extra, extra_index=slot)
            return
        cur = torch.cuda.current_stream(self.device)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.device)
        self._comm_stream.wait_stream(cur)
        if extra is not None and not (extra.is_cuda and extra.dtype == torch.float32 and extra.is_contiguous()):
            self.loss_slot.copy_(extra)
            extra = None
        with torch.cuda.stream(self._comm_stream):
            self._peer.all_reduce(first, count, flag_set=si % _be.PEER_FLAG_SETS, extra=extra, extra_index=slot)
        self._works.append(_StreamJoin(self._comm_stream, self.device))

    def reduce_segment(self, si):
        """Launch segment `si`'s all-reduce now (for loops that write gradients without autograd hooks)."""
        self._launch(si)

    def finish(self, loss=None):
        """Store this rank's share of the objective in its slot, launch the last segment's all-reduce and make the
        current stream wait for every collective of the step."""
        # the peer kernels store the scalar into the slot themselves (no copy launch); other backends copy it
        self._launch(len(self.segments) - 1, extra=None if loss is None else loss.detach().reshape(1))
        for w in self._works:
            w.wait()  # stream-level wait: the host does not block
        self._works = []

    def loss(self):
        """The all-reduced objective (valid after finish())."""
        return self.segments[-1][-1]


class _StreamJoin(object):
    """`.wait()` of a launch on a side stream: the current stream waits for it (no host blocking)."""

    def __init__(self, stream, device):
        self.stream, self.device = stream, device

    def wait(self):
        torch.cuda.current_stream(self.device).wait_stream(self.stream)


class _PeerBuffer(object):
    """A float32 buffer in peer-mapped (symmetric) device memory plus the flag area of zs_allreduce_sum_peer.
    torch.distributed._symmetric_memory does the plumbing -- allocation and the exchange of the mappings between the
    ranks of the group; the data movement is this library's kernel."""

    def __init__(self, n, device, dtype, group):
        import torch.distributed._symmetric_memory as symm
        if dtype != torch.float32 or torch.device(device).type != "cuda":
            raise RuntimeError("the peer all-reduce takes float32 CUDA buffers")
        pg = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(pg)
        if self.world > _be.MAX_PEERS:
            raise RuntimeError("the peer all-reduce takes at most %d ranks" % _be.MAX_PEERS)
        flag_floats = (_be.allreduce_peer_flag_bytes() + 3) // 4
        n_pad = (n + 3) // 4 * 4
        self.storage = symm.empty(n_pad + flag_floats, dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.storage, pg.group_name)
        self.storage.zero_()
        self.rank = self.handle.rank
        self.buf_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.flag_ptrs = [p + 4 * n_pad for p in self.buf_ptrs]
        # NVSwitch multicast mapping of the same allocation (0 when the fabric has none): the NVLS kernel
        mc = 0
        try:
            mc = int(self.handle.multicast_ptr)
        except Exception:
            mc = 0
        if os.environ.get("ZS_PEER_NVLS", "1") == "0" or self.storage.data_ptr() != self.buf_ptrs[self.handle.rank]:
            mc = 0
        self.mc_ptr = mc
        self.variant = "nvls" if mc else "p2p"
        self.flat = self.storage[:n]
        self.device = self.storage.device
        torch.cuda.synchronize(self.device)
        dist.barrier(group=pg)  # every rank's flags are zero before anyone's first kernel signals

    def tune(self, count, group=None, iters=30):
        """Pick the faster exchange kernel for THIS world size and bucket size by timing both on the buffer (its contents
        are zero at construction, so summing them repeatedly changes nothing).  The multicast form moves ~(1 + 1/N)
        buffer volumes over each GPU's links, peer loads / stores 2(N-1)/N: at N = 2 peer access wins (measured 20.6
        against 27.8 us for the 5.4 MB bucket), from N = 4 on the switch should.  Every rank times both and the ranks
        agree on the maxima, so they make the same choice.  ZS_PEER_NVLS=0 / 1 forces one."""
        forced = os.environ.get("ZS_PEER_NVLS")
        if not self.mc_ptr or forced in ("0", "1") or count <= 0:
            return self.variant
        mc = self.mc_ptr
        times = []
        for use_mc in (0, mc):
            self.mc_ptr = use_mc
            for _ in range(5):
                self.all_reduce(0, count)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                self.all_reduce(0, count)
            e1.record()
            torch.cuda.synchronize(self.device)
            times.append(e0.elapsed_time(e1) / iters)
        t = torch.tensor(times, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        t = [float(v) for v in t]
        self.tuned_us = {"p2p": round(t[0] * 1e3, 2), "nvls": round(t[1] * 1e3, 2)}
        self.mc_ptr = mc if t[1] < t[0] else 0
        self.variant = "nvls" if self.mc_ptr else "p2p"
        return self.variant

    def all_reduce(self, first, count, flag_set=0, extra=None, extra_index=0):
        """`extra`: a 1-element float32 tensor on this device that the kernel stores at float index `extra_index` of
        this rank's buffer before the exchange (the objective's slot: no copy launch)."""
        if self.mc_ptr:
            _be.allreduce_sum_nvls(self.mc_ptr, self.buf_ptrs[self.rank], self.flag_ptrs, self.rank, first, count,
                                   flag_set, self.device, extra=extra, extra_index=extra_index)
        else:
            _be.allreduce_sum_peer(self.buf_ptrs, self.flag_ptrs, self.rank, first, count, flag_set, self.device,
                                   extra=extra, extra_index=extra_index)


_legacy_buckets = {}


def all_reduce_gradients(params, n_local=None, n_global=None, group=None):
    """Turn gradients of the LOCAL mean loss into gradients of the GLOBAL mean loss: g <- sum_r g_r * n_r / n_global.
    Kept for callers that do not use `global_batch` + `GradientBucket` directly: the gradients are moved into a
    persistent flat buffer on first use (one copy, once), scaled in place by n_local / n_global (skip the scale by
    computing the loss under `global_batch`: pass n_local=None) and SUM-all-reduced with ONE collective; `.grad`
    stays a view of the buffer, so nothing is copied back."""
    params = [p for p in params if p.grad is not None]
    if not params:
        return
    key = tuple(id(p) for p in params)
    b = _legacy_buckets.get(key)
    if b is None:
        grads = [p.grad for p in params]
        b = GradientBucket([params], group=group)
        for p, g in zip(params, grads):  # the constructor zeroed the buffer after adopting the gradients
            p.grad.copy_(g)
        _legacy_buckets.clear()
        _legacy_buckets[key] = b
    else:
        for p in params:
            v = b._views[id(p)]
            if p.grad is not v:
                v.copy_(p.grad)
                p.grad = v
    if n_local is not None and n_global is not None and float(n_local) != float(n_global):
        b.flat.mul_(float(n_local) / float(n_global))
    if is_initialized():
        if b._peer is not None:
            b._peer.all_reduce(0, (b.flat.numel() + 3) // 4 * 4)
        else:
            dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=group)
