"""Breakdown of the host-buffer step: PCIe rates, the C-ABI host step, the public-API route. GPU box only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from zhusuan import _backend as be

K, B, X = 50, 1024, 784
dev = "cuda"
probs_h = torch.rand(K, B, X).clamp(0.01, 0.99).pin_memory()
x_h = (torch.rand(B, X) < 0.5).float().pin_memory()
other_h = (torch.randn(K, B) - 55).pin_memory(); logq_h = (torch.randn(K, B) + 30).pin_memory()
dprobs_h = torch.empty(K, B, X).pin_memory(); cost_h = torch.empty(B).pin_memory()
dlp_h = torch.empty(K, B).pin_memory(); dlq_h = torch.empty(K, B).pin_memory()
d = torch.empty(K, B, X, device=dev)

def t(name, fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("%-40s %7.2f ms" % (name, dt * 1e3)); return dt

nb = probs_h.numel() * 4
dt = t("H2D 160MB pinned", lambda: d.copy_(probs_h, non_blocking=True)); print("   %.1f GB/s" % (nb / dt / 1e9))
dt = t("D2H 160MB pinned", lambda: dprobs_h.copy_(d, non_blocking=True)); print("   %.1f GB/s" % (nb / dt / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(probs_h, non_blocking=True)
    with torch.cuda.stream(s2): dprobs_h.copy_(d, non_blocking=True)
dt = t("H2D + D2H concurrently", both); print("   %.1f GB/s aggregate" % (2 * nb / dt / 1e9))
ws = torch.empty(be.iw_step_host_workspace(K, B, X), dtype=torch.uint8, device=dev)
hs = be.HostStep(dev)
t("zs_iw_step_host (C ABI, pipelined)", lambda: be.iw_step_host(hs, be.SGVB, cost_h, dprobs_h, dlp_h, dlq_h, probs_h, x_h, other_h, logq_h, K, B, X, 1.0 / B, ws))
t("zs_iw_step_host, no dprobs", lambda: be.iw_step_host(hs, be.SGVB, cost_h, None, dlp_h, dlq_h, probs_h, x_h, other_h, logq_h, K, B, X, 1.0 / B, ws))
from zhusuan import _ops
def api():
    p = probs_h.detach().requires_grad_()
    loss = _ops.iw_bernoulli_fused_host(p, x_h, other_h, logq_h, be.SGVB)
    loss.backward()
    return p.grad
t("_ops.iw_bernoulli_fused_host + backward", api)
import bench
step = bench.e2e_setup(torch, False, B)
t("bench.e2e_setup step (full public API)", step)
