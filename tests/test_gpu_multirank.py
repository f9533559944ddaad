"""Two NCCL ranks on two GPUs against one GPU on the whole batch (injected noise): the data-parallel step of
zhusuan.distributed -- batch columns sharded, kernels scaled by 1/B_global (`global_batch`), parameter gradients
accumulated in a `GradientBucket` and SUM-all-reduced segment by segment while backward is still running, the scalar
objective in the bucket's last slot -- reproduces the single-GPU loss, the replicated decoder's gradients and each
rank's slice of the boundary gradients.  Needs two GPUs (run with `gpurun --gpus 2`; skipped on a one-GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

K, B, Z, X = 50, 64, 40, 784


def _inputs():
    rng = np.random.RandomState(7)
    return dict(x=(rng.uniform(size=(B, X)) < 0.5).astype(np.float32),
                mean=(0.5 * rng.standard_normal((B, Z))).astype(np.float32),
                std=np.exp(0.3 * rng.standard_normal((B, Z))).astype(np.float32),
                pq=(1.0 / (1.0 + np.exp(-rng.standard_normal((B, Z))))).astype(np.float32),
                eps=rng.standard_normal((K, B, Z)).astype(np.float32),
                u=rng.uniform(size=(K, B, Z)).astype(np.float32))


def _step(dev, cols, n_global, vimco, use_bucket, backend="auto"):
    """One objective step on the batch columns `cols` of the shared inputs; returns (loss, decoder, leaves, bucket)."""
    import zhusuan.distributed as zd
    from zhusuan import _rng
    from zhusuan.distributions import Bernoulli, Normal
    from zhusuan.framework import BayesianNet
    from zhusuan.variational import ImportanceWeightedObjective
    inp = _inputs()
    t = lambda a, g=False: torch.tensor(a[cols] if a.shape[0] == B else a[:, cols], device=dev).requires_grad_(g)
    x, eps, u = t(inp["x"]), t(inp["eps"]), t(inp["u"])
    a = t(inp["pq"] if vimco else inp["mean"], True)
    b = None if vimco else t(inp["std"], True)
    torch.manual_seed(0)
    dec = torch.nn.Sequential(torch.nn.Linear(Z, 64), torch.nn.ReLU(), torch.nn.Linear(64, X)).to(dev)
    n_loc = x.shape[0]

    class Gen(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            if vimco:
                z = self.bernoulli("z", probs=torch.full((n_loc, Z), 0.5, device=dev), n_samples=K, reduce_sum_dims=[2])
            else:
                z = self.normal("z", mean=torch.zeros(n_loc, Z, device=dev), std=torch.ones(n_loc, Z, device=dev),
                                is_reparameterized=False, n_samples=K, reduce_sum_dims=[2])
            self.sn(Bernoulli(probs=torch.sigmoid(dec(z))), name="x", reduce_sum_dims=[2])
            return self

    class Var(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            if vimco:
                self.sn(Bernoulli(probs=a), name="z", n_samples=K, reduce_sum_dims=[2])
            else:
                self.sn(Normal(mean=a, std=b), name="z", n_samples=K, reduce_sum_dims=[2])
            return self

    obj = ImportanceWeightedObjective(Gen(device=dev), Var(device=dev), axis=0, estimator="vimco" if vimco else "sgvb")
    bucket = zd.GradientBucket([dec.parameters()], backend=backend) if use_bucket else None
    inj = dict(uniform=[u, u]) if vimco else dict(normal=[eps, eps])
    with zd.global_batch(n_global), _rng.inject(**inj):
        loss = obj({"x": x})
    loss.backward()
    if bucket is not None:
        bucket.finish(loss)
    return loss.detach(), dec, (a, b), bucket


def _worker(rank, world, port, vimco, backend, out):
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "zhusuan-pytorch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import zhusuan.distributed as zd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    if backend == "peer-p2p":  # the peer loads / stores kernel even where the fabric offers a multicast mapping
        os.environ["ZS_PEER_NVLS"] = "0"
        backend = "peer"
    try:
        lo, hi = zd.shard_range(B)
        try:
            loss, dec, (a, b), bucket = _step(dev, slice(lo, hi), B, vimco, True, backend)
        except Exception as e:
            if backend != "peer":
                raise
            out.put((rank, {"unavailable": repr(e)[:300]}))
            dist.barrier()
            return
        torch.cuda.synchronize()
        res = {"loss": float(bucket.loss()), "backend": bucket.backend, "local_loss": float(loss), "lo": lo, "hi": hi,
               "variant": getattr(bucket._peer, "variant", None),
               "dec": [p.grad.detach().cpu().numpy() for p in dec.parameters()],
               "a": a.grad.detach().cpu().numpy(), "b": None if b is None else b.grad.detach().cpu().numpy(),
               "views": all(p.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
                            for p in dec.parameters())}
        out.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("backend", ["peer", "peer-p2p", "nccl"])
@pytest.mark.parametrize("vimco", [False, True])
def test_two_ranks_equal_one_gpu(vimco, backend):
    """backend "peer": the exchange is this library's kernel over NVLink peer memory -- the NVSwitch multicast form
    (zs_allreduce_sum_nvls) where the fabric offers a multicast mapping, else peer loads / stores
    (zs_allreduce_sum_peer), which "peer-p2p" forces; "nccl": torch.distributed.all_reduce on the same bucket.
    The objective reaches the bucket's slot through the kernel's `extra_src` (no copy launch)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29650 + (1 if vimco else 0) + {"peer": 2, "peer-p2p": 4, "nccl": 0}[backend]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, vimco, backend, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    if "unavailable" in got[0]:
        pytest.skip("peer-mapped memory not available on this box: " + got[0]["unavailable"])
    assert got[0]["backend"] == backend.split("-")[0] and got[1]["backend"] == backend.split("-")[0]
    if backend == "peer-p2p":
        assert got[0]["variant"] == "p2p"
    print("peer variant:", got[0]["variant"])
    # the whole batch on one GPU, no process group: global_batch(B) is then the plain batch mean
    dev = torch.device("cuda", 0)
    loss, dec, (a, b), _ = _step(dev, slice(0, B), B, vimco, False)
    ref_loss = float(loss)
    close = lambda x, y, tol=2e-5: np.testing.assert_allclose(x, y, rtol=tol, atol=tol * max(np.abs(y).max(), 1e-30))
    for r in (0, 1):
        g = got[r]
        assert g["views"]  # the decoder's .grad tensors live inside the bucket's flat buffer: nothing is copied back
        assert abs(g["loss"] - ref_loss) <= 1e-5 * abs(ref_loss)          # all-reduced objective on every rank
        for have, p in zip(g["dec"], dec.parameters()):                     # all-reduced parameter gradients
            close(have, p.grad.detach().cpu().numpy(), 1e-4)
        close(g["a"], a.grad.detach().cpu().numpy()[g["lo"]:g["hi"]])       # boundary gradients stay rank-local
        if b is not None:
            close(g["b"], b.grad.detach().cpu().numpy()[g["lo"]:g["hi"]])
    assert abs(got[0]["local_loss"] + got[1]["local_loss"] - ref_loss) <= 1e-5 * abs(ref_loss)
