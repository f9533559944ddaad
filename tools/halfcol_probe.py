import os, sys, json
ROOT = "/root/repo" if os.path.isdir("/root/repo/zhusuan-pytorch_b200") else os.getcwd()
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200")); sys.path.insert(0, ROOT)
import torch
from zhusuan import _backend as be
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(dev); be.load()
res = {}
for K, B in ((50, 1024), (25, 2048), (50, 2048), (25, 4096), (50, 148 * 7), (25, 148 * 14), (50, 128), (25, 256)):
    X = 784
    g = torch.Generator(device=dev); g.manual_seed(1)
    probs = torch.sigmoid(2.0 * torch.randn(K, B, X, device=dev, generator=g)).contiguous()
    x = (torch.rand(B, X, device=dev, generator=g) < 0.5).float()
    other = torch.randn(K, B, device=dev) - 55.0; logq = torch.randn(K, B, device=dev) + 30.0
    out = be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B)
    run, how, _ = bench.capture(torch, lambda: [be.iw_bernoulli_fused(be.SGVB, probs, x, other, logq, 1.0 / B, out=out) for _ in range(10)], dev, True)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    bytes_ = 2 * 4 * K * B * X + 4 * B * X + 16 * K * B
    res["K%d_B%d" % (K, B)] = {"us": round(us, 2), "GBs": round(bytes_ / us / 1e3, 1)}
    del probs, out
    torch.cuda.empty_cache()
print(json.dumps(res))
