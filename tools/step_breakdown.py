"""In-graph cost of each launch of the config-2 step: CUDA-graph replays of growing prefixes of the step's launch
sequence (C-ABI calls through zhusuan._backend), timed with CUDA events; the difference between two prefixes is what a
launch really adds to a replayed step (its duration plus the kernel -> kernel edge), which is not what a serialised
profiler run reports for it.  Prints one JSON line.  Env: ZS_PDL (launch-site mask), ZS_LATENT_BWD_DEEP."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
sys.path.insert(0, ROOT)
import torch
from zhusuan import _backend as be
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
be.load()
K, B, Z, X = 50, int(os.environ.get("ZS_B", "1024")), 40, 784
ks = bench.KernelSequence(torch, be, False, dev, seed=1, B=B)
one = torch.ones((), device=dev)
g1 = torch.ones(1, device=dev)


def seq(n):
    def fn():
        z, logq, logpz = be.normal_latent_fwd(ks.mean, ks.std, be.KBCAST, K, B, Z, seed=1, rng_state=ks.state)
        if n < 2:
            return z
        r = be.iw_bernoulli_fused(be.SGVB, ks.probs, ks.x, logpz, logq, 1.0 / B, cost_scaled=True, want_loss=n >= 5)
        if n < 3:
            return r
        if n >= 4:
            g = torch.ones_like(one)                       # what autograd launches for loss.backward()
            be.scale_inplace(r["dprobs"], g.reshape(1), r["dlogp"], r["dlogq"])
        return be.normal_latent_bwd(r["dlogq"], r["dlogp"], ks.dz_up, z, ks.mean, ks.std, be.KBCAST, K, B, Z,
                                    reparameterized=True)
    return fn


PER_GRAPH = 8  # repetitions captured into one graph: a replay then lasts long enough that the host's graph-launch
               # rate (~8 us per replay from Python) is not what the events measure


def timed(fn, reps=100):
    run, how, _ = bench.capture(torch, lambda: [fn() for _ in range(PER_GRAPH)], dev, True)
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * PER_GRAPH) * 1e3


names = {1: "fwd", 2: "fwd+fused", 3: "fwd+fused+bwd", 4: "fwd+fused+fill+scale+bwd", 5: "same, loss written by the fused launch"}
res = {"pdl": os.environ.get("ZS_PDL", "7"), "bwd_deep": os.environ.get("ZS_LATENT_BWD_DEEP", "1"),
       "fwd_rows": os.environ.get("ZS_LATENT_FWD_ROWS", "1"), "B": B}
for n in (1, 2, 3, 4, 5):
    res[names[n]] = round(timed(seq(n)), 2)
fused_only = lambda: ks.fused_only(torch.zeros(K, B, device=dev), torch.zeros(K, B, device=dev))
other, logq = torch.randn(K, B, device=dev) - 55.0, torch.randn(K, B, device=dev) + 30.0
out = ks.fused_only(other, logq)
res["fused alone"] = round(timed(lambda: ks.fused_only(other, logq, out=out)), 2)
z, lq, lp = be.normal_latent_fwd(ks.mean, ks.std, be.KBCAST, K, B, Z, seed=1, rng_state=ks.state)
res["bwd alone"] = round(timed(lambda: be.normal_latent_bwd(out["dlogq"], out["dlogp"], ks.dz_up, z, ks.mean, ks.std, be.KBCAST,
                                                            K, B, Z, reparameterized=True)), 2)
print(json.dumps(res), flush=True)
