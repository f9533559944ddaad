"""world_size-2 `gloo` test of the sharded path on CPU: batch columns split over two ranks, the
objective and the parameter gradients all-reduced, must equal the single-process result on the whole
batch.  The kernels are replaced by the CPU oracle (tests/oracle_backend.py) — this exercises the host
logic of zhusuan.distributed and of the objective under sharding, not the GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
K, B, Z, X = 8, 10, 4, 16


def _problem():
    rng = np.random.RandomState(3)
    return dict(w=torch.tensor(0.3 * rng.standard_normal((X, Z)), dtype=torch.float32),     # "encoder" weights
                dec=torch.tensor(0.3 * rng.standard_normal((Z, X)), dtype=torch.float32),   # "decoder" weights
                x=torch.tensor((rng.uniform(size=(B, X)) < 0.5), dtype=torch.float32),
                eps=torch.tensor(rng.standard_normal((K, B, Z)), dtype=torch.float32))


def _loss_and_grads(cols, prob, estimator):
    """Local mean IW loss over `cols` through the public API, with shared parameters w / dec."""
    import zhusuan
    from zhusuan import _rng
    from zhusuan.distributions import Bernoulli, Normal
    from zhusuan.framework import BayesianNet
    from zhusuan.variational import ImportanceWeightedObjective
    w = prob["w"].clone().requires_grad_()
    dec = prob["dec"].clone().requires_grad_()
    x = prob["x"][cols]
    eps = prob["eps"][:, cols, :].contiguous()
    n = x.shape[0]

    class Gen(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            z = self.normal("z", mean=torch.zeros(n, Z), std=torch.ones(n, Z), is_reparameterized=False, n_samples=K,
                            reduce_sum_dims=[2])
            self.sn(Bernoulli(probs=torch.sigmoid(z @ dec)), name="x", reduce_sum_dims=[2])
            return self

    class Var(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            self.sn(Normal(mean=observed["x"] @ w, std=torch.ones(n, Z), is_reparameterized=estimator == "sgvb"),
                    name="z", n_samples=K, reduce_sum_dims=[2])
            return self

    obj = ImportanceWeightedObjective(Gen(), Var(), axis=0, estimator=estimator)
    with _rng.inject(normal=[eps, eps]):
        loss = obj({"x": x})
    loss.backward()
    return loss.detach(), [w, dec]


def _worker(rank, world_size, port, estimator, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend

    class MP(object):  # minimal monkeypatch stand-in
        def setattr(self, obj, name, value):
            setattr(obj, name, value)

    oracle_backend.install(MP())
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from zhusuan import distributed as zd
    prob = _problem()
    lo, hi = zd.shard_range(B)
    assert (lo, hi) == zd.shard_range(B, rank, world_size)
    loss, params = _loss_and_grads(slice(lo, hi), prob, estimator)
    gl = zd.global_mean_objective(loss, hi - lo, B)
    zd.all_reduce_gradients(params, hi - lo, B)
    zd.decorrelate_rng()
    from zhusuan import _rng
    assert _rng.rank_stride == rank * zd.RANK_OFFSET_STRIDE
    if rank == 0:
        torch.save(dict(loss=gl, grads=[p.grad.clone() for p in params]), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("estimator", ["sgvb", "vimco"])
def test_two_rank_sharding_matches_single_process(tmp_path, estimator, monkeypatch):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, estimator, out), nprocs=2, join=True)
    got = torch.load(out)
    import oracle_backend
    oracle_backend.install(monkeypatch)
    loss, params = _loss_and_grads(slice(0, B), _problem(), estimator)
    np.testing.assert_allclose(float(got["loss"]), float(loss), rtol=1e-5)
    for g, p in zip(got["grads"], params):
        np.testing.assert_allclose(g.numpy(), p.grad.numpy(), rtol=2e-4, atol=2e-5 * float(p.grad.abs().max()))


def _loss_bucket(cols, prob, estimator, n_global):
    """The round-2 protocol: kernels scaled by 1/n_global (global_batch), gradients accumulated in a two-segment
    GradientBucket (decoder first, encoder second) whose all-reduces are launched from backward hooks."""
    import zhusuan.distributed as zd
    from zhusuan import _rng
    from zhusuan.distributions import Bernoulli, Normal
    from zhusuan.framework import BayesianNet
    from zhusuan.variational import ImportanceWeightedObjective
    w = prob["w"].clone().requires_grad_()
    dec = prob["dec"].clone().requires_grad_()
    x = prob["x"][cols]
    eps = prob["eps"][:, cols, :].contiguous()
    n = x.shape[0]

    class Gen(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            z = self.normal("z", mean=torch.zeros(n, Z), std=torch.ones(n, Z), is_reparameterized=False, n_samples=K,
                            reduce_sum_dims=[2])
            self.sn(Bernoulli(probs=torch.sigmoid(z @ dec)), name="x", reduce_sum_dims=[2])
            return self

    class Var(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            self.sn(Normal(mean=observed["x"] @ w, std=torch.ones(n, Z), is_reparameterized=estimator == "sgvb"),
                    name="z", n_samples=K, reduce_sum_dims=[2])
            return self

    obj = ImportanceWeightedObjective(Gen(), Var(), axis=0, estimator=estimator)
    bucket = zd.GradientBucket([[dec], [w]])
    assert dec.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
    out = []
    for _ in range(2):  # two steps: zero_grad() re-arms the segment counters and the views survive
        bucket.zero_grad()
        with zd.global_batch(n_global), _rng.inject(normal=[eps, eps]):
            loss = obj({"x": x})
        loss.backward()
        bucket.finish(loss)
        assert dec.grad.untyped_storage().data_ptr() == bucket.flat.untyped_storage().data_ptr()
        out.append((bucket.loss().clone(), [dec.grad.clone(), w.grad.clone()]))
    assert torch.equal(out[0][0], out[1][0]) and all(torch.equal(a, b) for a, b in zip(out[0][1], out[1][1]))
    return out[-1]


def _bucket_worker(rank, world_size, port, estimator, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_backend

    class MP(object):
        def setattr(self, obj, name, value):
            setattr(obj, name, value)

    oracle_backend.install(MP())
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from zhusuan import distributed as zd
    lo, hi = zd.shard_range(B)
    loss, grads = _loss_bucket(slice(lo, hi), _problem(), estimator, B)
    if rank == 1:  # every rank holds the reduced values
        torch.save(dict(loss=loss, grads=grads), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("estimator", ["sgvb", "vimco"])
def test_gradient_bucket_and_global_batch_match_single_process(tmp_path, estimator, monkeypatch):
    out = str(tmp_path / "r1.pt")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_bucket_worker, args=(2, port, estimator, out), nprocs=2, join=True)
    got = torch.load(out)
    import oracle_backend
    oracle_backend.install(monkeypatch)
    loss, params = _loss_and_grads(slice(0, B), _problem(), estimator)  # the whole batch, plain mean objective
    np.testing.assert_allclose(float(got["loss"]), float(loss), rtol=1e-5)
    for g, p in zip(got["grads"], [params[1], params[0]]):  # bucket order: decoder, encoder
        np.testing.assert_allclose(g.numpy(), p.grad.numpy(), rtol=2e-4, atol=2e-5 * float(p.grad.abs().max()))
    # without a process group the bucket and global_batch degenerate to the local computation
    loss1, grads1 = _loss_bucket(slice(0, B), _problem(), estimator, B)
    np.testing.assert_allclose(float(loss1), float(loss), rtol=1e-6)
    for g, p in zip(grads1, [params[1], params[0]]):
        np.testing.assert_allclose(g.numpy(), p.grad.numpy(), rtol=1e-5, atol=1e-6 * float(p.grad.abs().max()))


def test_shard_range_is_a_partition():
    from zhusuan import distributed as zd
    for n in (1, 7, 8, 1024, 1025):
        for w in (1, 2, 3, 8):
            parts = [zd.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
