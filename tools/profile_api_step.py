"""cProfile of the host-resident public-API step (bench.py's e2e leg).  GPU box only.  Usage: python tools/profile_api_step.py [vimco]"""
import os, sys, cProfile, pstats, io, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200")); sys.path.insert(0, ROOT)
import torch, zhusuan, bench
vimco = len(sys.argv) > 1 and sys.argv[1] == "vimco"
torch.cuda.set_device(0)
bench.pin_to_gpu_numa(0, 0, 1)
step = bench.e2e_setup(torch, vimco, 1024)
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    step()
    torch.cuda.synchronize(); print("step %.2f ms" % ((time.perf_counter() - t0) * 1e3))
