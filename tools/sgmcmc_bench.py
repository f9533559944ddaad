"""Config 5 secondary numbers: SG-MCMC update kernels across parallel chains (GPU box only).
1024 chains x 4601 weights (bnn_sgmcmc.py, layers [90,50,1]) and a 64x scaled-up chain count."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
from zhusuan import _backend as be

def timeit(fn, reps=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3

out = []
for chains in (1024, 65536):
    n = chains * 4601
    w, g, v, aux = (torch.randn(n, device="cuda") for _ in range(4))
    aux.abs_()
    wo = torch.empty_like(w)
    cases = {
        "sgld (12 B/elt)": (12, lambda: be.sgld_step(w, g, 1e-3, seed=1, offset=4, out=wo)),
        "psgld (20 B/elt)": (20, lambda: be.psgld_step(w, aux, g, 1e-3, 0.9, 1e-3, seed=1, offset=4, out=wo)),
        "sghmc first order (20 B/elt)": (20, lambda: be.sghmc_post(w, v, g, 1e-3, 0.25, 0.0, False, seed=1, offset=4, out=wo)),
        "sghmc second order pre+post (32 B/elt)": (32, lambda: (be.sghmc_pre(w, v, 1e-3, False, True, out=wo),
                                                                be.sghmc_post(wo, v, g, 1e-3, 0.25, 0.0, True, seed=1, offset=4, out=wo))),
    }
    for name, (bpe, fn) in cases.items():
        t = timeit(fn, 200 if chains == 1024 else 30)
        out.append(dict(chains=chains, elements=n, kernel=name, us=round(t * 1e6, 2), gbs=round(n * bpe / t / 1e9, 1),
                        chain_steps_per_s=round(chains / t)))
        print(out[-1])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "sgmcmc_r1.json"), "w"), indent=1)
