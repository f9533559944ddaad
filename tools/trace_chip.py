"""Chip-wide phase timeline of the ring kernel from %globaltimer marks (thread 0 of every CTA). GPU box only."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import numpy as np
import torch
from zhusuan import _backend as be

K, B, X = 50, 1024, 784
dev = "cuda"
probs = torch.sigmoid(2 * torch.randn(K, B, X, device=dev))
x = (torch.rand(B, X, device=dev) < 0.5).float()
other = torch.randn(K, B, device=dev) - 55
logq = torch.randn(K, B, device=dev) + 30
for _ in range(3):
    be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B)
trace = torch.zeros(400 * 40, dtype=torch.int64, device=dev)
be.load().zs_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B)
e1.record()
torch.cuda.synchronize()
be.load().zs_debug_set_trace(None)
print("kernel (traced build) %.1f us" % (e0.elapsed_time(e1) * 1e3))
t = trace.cpu().numpy().reshape(400, 8, 5).astype(np.int64)
n = int((t[:, 0, 0] != 0).sum())
t = t[:n]
t0 = t[:, 0, 0].min()
print("CTAs traced:", n, " first A(0) start spread: %.2f us" % ((t[:, 0, 0].max() - t0) / 1e3))
for c in range(7):
    ok = t[:, c, 0] != 0
    if not ok.any():
        break
    a0, a1, b0, b1 = [(t[ok, c, i] - t0) / 1e3 for i in range(4)]
    print("col %d (%3d CTAs): A start %6.2f..%6.2f  A end %6.2f..%6.2f  B start %6.2f..%6.2f  B end %6.2f..%6.2f   mean A %.2f  bar %.2f  B %.2f" % (
        c, ok.sum(), a0.min(), a0.max(), a1.min(), a1.max(), b0.min(), b0.max(), b1.min(), b1.max(),
        (a1 - a0).mean(), (b0 - a1).mean(), (b1 - b0).mean()))
last = t[:, :, 3].max()
print("last B end %.2f us after first A start" % ((last - t0) / 1e3))
# occupancy of phases over time
edges = np.arange(0, (last - t0) / 1e3 + 2, 2.0)
print("time(us)   #CTAs in A   #in B   #between")
for lo in edges[:-1]:
    mid = lo + 1.0
    tt = t0 + mid * 1e3
    inA = ((t[:, :7, 0] <= tt) & (t[:, :7, 1] > tt) & (t[:, :7, 0] != 0)).any(axis=1).sum()
    inB = ((t[:, :7, 2] <= tt) & (t[:, :7, 3] > tt) & (t[:, :7, 0] != 0)).any(axis=1).sum()
    print("%6.1f   %4d  %4d  %4d" % (mid, inA, inB, n - inA - inB))
