"""Phase timeline of the L2-resident fused kernel (clock64 trace). GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import ctypes
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
be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B)
torch.cuda.synchronize()
be.load().zs_debug_set_trace(None)
t = trace.cpu().numpy().reshape(400, 8, 5)
for cta in (0, 1, 74, 147):
    tt = t[cta]
    if tt[0, 0] == 0:
        continue
    print("CTA", cta, "(cycles @ ~1.97 GHz; thread 0 = row warp 0)")
    for c in range(8):
        if tt[c, 0] == 0:
            break
        nxt = tt[c + 1, 0] if c + 1 < 8 and tt[c + 1, 0] else None
        print("  col %d: phaseA %6d  barrier %6d  phaseB %6d  -> next col %s" % (
            c, tt[c, 1] - tt[c, 0], tt[c, 2] - tt[c, 1], tt[c, 3] - tt[c, 2], (nxt - tt[c, 3]) if nxt else "-"))
    print("  total %d cycles for %d columns" % (tt[max(i for i in range(8) if tt[i,0])][3] - tt[0, 0], sum(1 for i in range(8) if tt[i, 0])))
