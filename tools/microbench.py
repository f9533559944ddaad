"""Per-kernel timings at config-2 size (K=50, B=1024, X=784) with CUDA events. GPU box only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
from zhusuan import _backend as be

K, B, X, Z = 50, int(os.environ.get("MB_B", 1024)), 784, 40
dev = "cuda"
probs = torch.sigmoid(2 * torch.randn(K, B, X, device=dev))
x = (torch.rand(B, X, device=dev) < 0.5).float()
g = torch.randn(K, B, device=dev)
other = torch.randn(K, B, device=dev) - 55
logq = torch.randn(K, B, device=dev) + 30
out = torch.empty_like(probs)


def timeit(name, fn, nbytes, reps=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    t_issue = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-34s %8.1f us  %7.1f GB/s   (host issue %.1f us/call)" % (name, ms * 1e3, nbytes / ms / 1e6, t_issue * 1e6))


nb = probs.numel() * 4
timeit("torch copy_ 160MB (R+W)", lambda: out.copy_(probs), 2 * nb)
timeit("torch mul (R+W)", lambda: torch.mul(probs, 2.0, out=out), 2 * nb)
timeit("bernoulli_logpmf_fwd (R)", lambda: be.bernoulli_logpmf_fwd(x, be.KBCAST, probs, be.FULL, K, B, X), nb)
timeit("bernoulli_logpmf_bwd (R+W)", lambda: be.bernoulli_logpmf_bwd(g, x, be.KBCAST, probs, be.FULL, K, B, X, False, True), 2 * nb)
timeit("iw_objective [K,B]", lambda: be.iw_objective(be.SGVB, other, logq, 1.0 / B), 4 * K * B * 4)
pre = dict(cost=torch.empty(B, device=dev), dprobs=out, dlogp=torch.empty(K, B, device=dev), dlogq=torch.empty(K, B, device=dev))
timeit("fused (prealloc outputs)", lambda: be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B, out=pre), 2 * nb)
timeit("fused (alloc outputs)", lambda: be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B), 2 * nb)
timeit("fused fwd only (no dprobs)", lambda: be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B, need_dprobs=False, out=pre), nb)
timeit("fused vimco", lambda: be.iw_bernoulli_fused(be.VIMCO, probs, x, other, logq, 1.0 / B, out=pre), 2 * nb)
