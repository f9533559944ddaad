"""Reference fixture at BENCHMARK row sizes: K = 50 particles, X = 784 pixels, Z = 40 latents (B = 8 columns), so that
the kernel the bench times -- k_iw_bernoulli_boxf<., 28, 7, .>, reached through the public API on CUDA -- is compared
with the REAL reference, not only with the oracle ("weak" #1).

    PYTHONPATH=/root/reference python tests/golden/make_golden_bench.py

Cases: sgvb with Normal latents, vimco with Bernoulli latents, sgvb with a likelihood given by logits.  Inputs are
regenerated from seeds by the tests (`inputs()` below is imported by them), the fixture stores the reference's
outputs: loss and the boundary gradients in float32, and in float64 the loss, the small gradients and the first two
batch columns of the big one.  The file also records how far the reference's own float32 run is from its float64 run
(max-norm and relative L2 per output) -- the budget against which tools/parity_budget.py judges the kernels.
Nothing here is imported by the product.
"""
import os
import sys
from unittest import mock

import numpy as np

OUT = os.path.dirname(os.path.abspath(__file__))
K, B, Z, X = 50, 8, 40, 784
F64_COLS = 2


def inputs():
    """Seeded float64 inputs of the three cases (shared with tests/ and tools/parity_budget.py)."""
    rng = np.random.RandomState(2026)
    d = dict(
        mean=0.5 * rng.standard_normal((B, Z)), logstd=0.3 * rng.standard_normal((B, Z)),
        probs_q=1.0 / (1.0 + np.exp(-rng.standard_normal((B, Z)))),
        logits=2.0 * rng.standard_normal((K, B, X)),
        x=(rng.uniform(size=(B, X)) < 0.5).astype(np.float64),
        eps=rng.standard_normal((K, B, Z)), u=rng.uniform(size=(K, B, Z)))
    d["probs"] = 1.0 / (1.0 + np.exp(-d["logits"]))
    return d


CASES = (("sgvb_normal", "sgvb", "normal", False), ("vimco_bernoulli", "vimco", "bernoulli", False),
         ("sgvb_normal_logits", "sgvb", "normal", True))


def main():
    import torch
    ref = os.environ.get("ZS_REFERENCE", "/root/reference")
    sys.path.insert(0, ref)
    import zhusuan
    assert os.path.realpath(zhusuan.__file__).startswith(os.path.realpath(ref)), zhusuan.__file__
    sys.path.insert(0, OUT)
    import make_golden as G  # _Gen / _Var: the path-boundary nets of the small fixtures

    from zhusuan.variational import ImportanceWeightedObjective
    inp = inputs()
    out, budget = {}, {}
    for cname, est, latent, use_logits in CASES:
        res = {}
        for dn, dt in (("f32", torch.float32), ("f64", torch.float64)):
            big = G.t(inp["logits"] if use_logits else inp["probs"], dt, True)
            if latent == "normal":
                a, b = G.t(inp["mean"], dt, True), G.t(inp["logstd"], dt, True)
            else:
                a, b = G.t(inp["probs_q"], dt, True), None
            eps, u = G.t(inp["eps"], dt), G.t(inp["u"], dt)

            def fake_normal(*args, **kw):
                if "size" in kw:
                    return eps.clone()
                return (args[0] + args[1] * eps).detach()

            def fake_bernoulli(p, *args, **kw):
                return (u < p).to(p.dtype)

            gen = G._Gen(big, K, latent, logits=use_logits)
            var = G._Var(a, b, K, latent, reparam=(est == "sgvb"))
            obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
            with mock.patch("torch.normal", fake_normal), mock.patch("torch.bernoulli", fake_bernoulli):
                loss = obj({"x": G.t(inp["x"], dt)})
            leaves = [big, a] + ([b] if b is not None else [])
            grads = torch.autograd.grad(loss, leaves)
            res[dn] = dict(loss=G.npy(loss), dbig=G.npy(grads[0]), da=G.npy(grads[1]),
                           db=G.npy(grads[2]) if b is not None else None)
        f32, f64 = res["f32"], res["f64"]
        p = cname + "_"
        out[p + "f32_loss"], out[p + "f64_loss"] = f32["loss"], f64["loss"]
        out[p + "f32_dbig"] = f32["dbig"]
        out[p + "f64_dbig_cols"] = f64["dbig"][:, :F64_COLS]
        out[p + "f32_da"], out[p + "f64_da"] = f32["da"], f64["da"]
        if f32["db"] is not None:
            out[p + "f32_db"], out[p + "f64_db"] = f32["db"], f64["db"]
        # the reference's own float32-vs-float64 distance, on the FULL arrays
        for key in ("loss", "dbig", "da", "db"):
            if f32[key] is None:
                continue
            a32, a64 = np.asarray(f32[key], np.float64), np.asarray(f64[key], np.float64)
            scale = max(np.abs(a64).max(), 1e-300)
            budget["%s_%s" % (cname, key)] = [float(np.abs(a32 - a64).max() / scale),
                                              float(np.linalg.norm(a32 - a64) / max(np.linalg.norm(a64), 1e-300))]
    for k_, v in budget.items():
        out["budget_" + k_] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, "bench_size.npz"), **out)
    print("wrote bench_size.npz", {k_: getattr(v, "shape", None) for k_, v in out.items()})
    for k_, v in sorted(budget.items()):
        print("ref f32 vs ref f64  %-28s max-norm %.3e  rel-L2 %.3e" % (k_, v[0], v[1]))


if __name__ == "__main__":
    main()
