import os, sys, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200")); sys.path.insert(0, ROOT)
import torch, zhusuan, bench
K, B, X = 50, 1024, 784
vimco = len(sys.argv) > 1 and sys.argv[1] == "vimco"
host = {"probs": torch.rand(K, B, X).clamp(0.01, 0.99).pin_memory(), "x": (torch.rand(B, X) < 0.5).float().pin_memory(),
        "mean": (0.5*torch.randn(B,40)).pin_memory(), "std": torch.rand(B,40).add(0.5).pin_memory(),
        "zeros": torch.zeros(B,40).pin_memory(), "ones": torch.ones(B,40).pin_memory(),
        "pq": torch.rand(B,40).clamp(0.1,0.9).pin_memory(), "prior": torch.full((B,40),0.5).pin_memory()}
for _ in range(3): bench.api_step_host(torch, zhusuan, vimco, host)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): bench.api_step_host(torch, zhusuan, vimco, host)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
import time
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    bench.api_step_host(torch, zhusuan, vimco, host)
    torch.cuda.synchronize(); print("step %.2f ms" % ((time.perf_counter() - t0) * 1e3))
