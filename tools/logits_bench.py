"""Fused logits form of the likelihood + objective kernel against the probs form plus the Sigmoid round trips it
replaces (torch.sigmoid forward, its backward).  Config 2 (K=50, B=1024, X=784).  GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
from zhusuan import _backend as be

K, B, X = 50, 1024, 784
dev = "cuda"
torch.manual_seed(0)
logits = (2 * torch.randn(K, B, X, device=dev)).contiguous()
x = (torch.rand(B, X, device=dev) < 0.5).float()
other = torch.randn(K, B, device=dev) - 55
logq = torch.randn(K, B, device=dev) + 30


def t(name, fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("%-64s %8.1f us" % (name, us))
    return us


out = {}
a = t("zs_iw_bernoulli_fused_logits (logits in, dlogits out)", lambda: be.iw_bernoulli_fused(be.SGVB, logits, x, other, logq, 1.0 / B, logits=True, out=out))
probs = torch.sigmoid(logits)
b1 = t("torch.sigmoid forward", lambda: torch.sigmoid(logits))
out2 = {}
b2 = t("zs_iw_bernoulli_fused (probs in, dprobs out)", lambda: be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B, out=out2))
r = be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B)
b3 = t("sigmoid backward (dprobs * p * (1 - p))", lambda: torch.ops.aten.sigmoid_backward(r["dprobs"], probs))
print("logits form %.1f us vs probs form + Sigmoid round trips %.1f us (%.2fx); bytes %.0f MB vs %.0f MB" % (
    a, b1 + b2 + b3, (b1 + b2 + b3) / a, 2 * logits.numel() * 4 / 1e6, 7 * logits.numel() * 4 / 1e6))
rl = be.iw_bernoulli_fused(be.SGVB, logits, x, other, logq, 1.0 / B, logits=True)
ref = torch.ops.aten.sigmoid_backward(r["dprobs"], probs)
print("max rel diff dlogits vs composed: %.3g" % ((rl["dprobs"] - ref).abs().max() / ref.abs().max()).item())
