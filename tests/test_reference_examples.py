"""Drop-in check with the reference's OWN example models: the `Generator` / `Variational` / `Net` classes of
examples/variational_autoencoder/{iwae,vae_mnist}.py and examples/bayesian_neural_nets/{bnn_vi,bnn_sgmcmc}.py are imported
UNMODIFIED from baseline/_ref/examples (tools/vendor_reference.py) while `zhusuan` resolves to THIS package, and run
one optimiser / sampler step.

On the GPU box (-m gpu) the same models, weights and injected noise are also run by the unmodified REFERENCE in a
subprocess (its package is called `zhusuan` too) on the CPU, and the two losses are compared.
The CPU variants run the package's host logic on the oracle stand-in backend (tests/oracle_backend.py)."""
import importlib
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "examples")),
                               reason="baseline/_ref not vendored (python tools/vendor_reference.py)")


def _stub_plot_modules():
    for name in ("PIL", "PIL.Image", "matplotlib", "matplotlib.pyplot"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            sys.modules[name] = m
            parent, _, child = name.rpartition(".")
            if parent:
                setattr(sys.modules[parent], child, m)


def _example(modname):
    """Import examples.<modname> from baseline/_ref with `zhusuan` bound to this package."""
    import zhusuan
    assert "zhusuan-pytorch_b200" in os.path.realpath(zhusuan.__file__), zhusuan.__file__
    _stub_plot_modules()
    if REF not in sys.path:
        sys.path.append(REF)  # AFTER the package path: `import zhusuan` keeps resolving to this package
    for k in [k for k in sys.modules if k == "examples" or k.startswith("examples.")]:
        del sys.modules[k]
    mod = importlib.import_module("examples." + modname)
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF))
    return mod


def _iwae_step(device, vimco, K=8, B=16):
    from zhusuan.variational.importance_weighted_objective import ImportanceWeightedObjective
    iwae = _example("variational_autoencoder.iwae")
    iwae.device = torch.device(device)
    iwae.reparameterization = not vimco  # module global read inside Variational.forward
    torch.manual_seed(0)
    gen, var = iwae.Generator(784, 40, K), iwae.Variational(784, 40, K)
    model = ImportanceWeightedObjective(gen, var, axis=0, estimator="vimco" if vimco else "sgvb").to(torch.device(device))
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    x = (torch.rand(B, 784) < 0.5).float().to(device)
    before = [p.detach().clone() for p in model.parameters()]
    losses = []
    for _ in range(2):
        loss = model({"x": x})
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(np.isfinite(losses))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, model.parameters()))
    # the tail of the example's main(): reconstruct and sample through the nets' caches (iwae.py:176-182)
    z = var({"x": x}).nodes["z"].tensor
    assert gen({"z": z}).cache["x_mean"].shape == (K, B, 784)
    return losses


def _vae_step(device, B=16):
    from zhusuan.variational.elbo import ELBO
    vae = _example("variational_autoencoder.vae_mnist")
    torch.manual_seed(0)
    gen, var = vae.Generator(784, 40, B), vae.Variational(784, 40, B)
    model = ELBO(gen, var).to(torch.device(device))
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    x = (torch.rand(B, 784) < 0.5).float().to(device)
    loss = model({"x": x})
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert np.isfinite(float(loss)) and all(p.grad is not None for p in model.parameters())
    return float(loss)


def _bnn_vi_step(device, K=10, B=24):
    from zhusuan.variational.elbo import ELBO
    bnn = _example("bayesian_neural_nets.bnn_vi")
    torch.manual_seed(0)
    layers = [13, 50, 1]
    net, var = bnn.Net(layers, K), bnn.Variational(layers, K)
    model = ELBO(net, var)
    model.to(torch.device(device))
    opt = torch.optim.Adam(model.parameters(), 1e-3)
    x, y = torch.randn(B, 13).to(device), torch.randn(B).to(device)
    loss = model({"x": x, "y": y})
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert np.isfinite(float(loss)) and np.isfinite(float(net.cache["rmse"]))
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    return float(loss)


def _bnn_sgmcmc_steps(device, K=6, B=24):
    from zhusuan.mcmc.SGLD import SGLD
    bnn = _example("bayesian_neural_nets.bnn_sgmcmc")
    torch.manual_seed(0)
    layers = [13, 50, 1]
    net = bnn.Net(layers, K).to(torch.device(device))
    model = SGLD(1e-3).to(torch.device(device))
    x, y = torch.randn(B, 13).to(device), torch.randn(B).to(device)
    prev = None
    for step in range(3):
        w = model.sample(net, {"x": x, "y": y}, step == 0)
        assert set(w) == {"w0", "w1"} and w["w0"].shape == (K, 50, 14) and w["w1"].shape == (K, 1, 51)
        for i, (k, v) in enumerate(w.items()):  # the example's per-step re-estimate of the prior scale (:124-127)
            net.w_logstds[i] = (0.5 * torch.log(torch.mean(v * v, [0]))).detach()
        if prev is not None and step > 1:
            assert not torch.equal(prev, w["w0"].detach())
        prev = w["w0"].detach().clone()
    net.forward({**w, "x": x, "y": y})
    assert np.isfinite(float(net.cache["rmse"]))


# ----------------------------------------------------------------------------- CPU: host logic on the oracle stand-in
@needs_ref
@pytest.mark.parametrize("vimco", [False, True])
def test_iwae_example_models_cpu(monkeypatch, vimco):
    import oracle_backend
    oracle_backend.install(monkeypatch)
    _iwae_step("cpu", vimco, K=4, B=6)


@needs_ref
def test_vae_and_bnn_example_models_cpu(monkeypatch):
    import oracle_backend
    oracle_backend.install(monkeypatch)
    _vae_step("cpu", B=6)
    _bnn_vi_step("cpu", K=4, B=8)
    _bnn_sgmcmc_steps("cpu", K=4, B=8)


# ----------------------------------------------------------------------------- GPU: the real kernels
@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("vimco", [False, True])
def test_iwae_example_models_gpu(vimco):
    _iwae_step("cuda", vimco, K=50, B=64)


@needs_ref
@pytest.mark.gpu
def test_vae_and_bnn_example_models_gpu():
    _vae_step("cuda", B=64)
    _bnn_vi_step("cuda", K=100, B=114)
    _bnn_sgmcmc_steps("cuda", K=20, B=114)


_REF_SCRIPT = r'''
import json, sys, types
from unittest import mock
import numpy as np
import torch
sys.path.insert(0, sys.argv[1])
for name in ("PIL", "PIL.Image", "matplotlib", "matplotlib.pyplot"):
    try:
        __import__(name)
    except Exception:
        m = types.ModuleType(name); sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent: setattr(sys.modules[parent], child, m)
import zhusuan
assert zhusuan.__file__.startswith(sys.argv[1]), zhusuan.__file__
from zhusuan.variational.importance_weighted_objective import ImportanceWeightedObjective
import examples.variational_autoencoder.iwae as iwae
vimco = sys.argv[3] == "1"
K, B = 50, 8
iwae.device = torch.device("cpu")
iwae.reparameterization = not vimco
d = np.load(sys.argv[2])
res = {}
for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
    torch.manual_seed(0)
    gen, var = iwae.Generator(784, 40, K), iwae.Variational(784, 40, K)
    model = ImportanceWeightedObjective(gen, var, axis=0, estimator="vimco" if vimco else "sgvb").to(dt)
    x, eps = torch.tensor(d["x"]).to(dt), torch.tensor(d["eps"]).to(dt)
    def fake_normal(*a, **k):
        if "size" in k:
            return eps.clone()
        return (a[0] + a[1] * eps).detach()
    with mock.patch("torch.normal", fake_normal):
        loss = model({"x": x})
    loss.backward()
    out = {"loss": float(loss)}
    for n, p in model.named_parameters():
        out[n] = float(p.grad.norm())
    res[tag] = out
print(json.dumps(res))
'''


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("vimco", [False, True])
def test_iwae_example_matches_the_reference_run(tmp_path, vimco):
    """Same example classes, same initial weights (seeded nn.Linear init), same injected noise: the reference on the CPU
    (subprocess, its own `zhusuan`, once in float32 and once with the same weights cast to float64) and this package on
    the GPU agree on the loss and on every parameter-gradient norm. The yardstick is the float64 run; the budget is
    north_star's 1e-5-class bound (2e-5 on the loss, 2e-4 on gradient norms that pass through exp(log w)) or 1.5 x the
    distance of the reference's OWN float32 run from its float64 run, whichever is larger: VIMCO's value holds
    sum_k log q_k * signal_k with the signal formed as a difference of two O(100) float32 numbers, which leaves the
    float32 reference 3e-5 (loss) / 6e-4 (encoder gradients) away from float64 on these inputs."""
    from zhusuan import _rng
    from zhusuan.variational.importance_weighted_objective import ImportanceWeightedObjective
    K, B = 50, 8
    rng = np.random.RandomState(5)
    x = (rng.uniform(size=(B, 784)) < 0.5).astype(np.float32)
    eps = rng.standard_normal((K, B, 40)).astype(np.float32)
    f = str(tmp_path / "in.npz")
    np.savez(f, x=x, eps=eps)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", _REF_SCRIPT, REF, f, "1" if vimco else "0"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    both = json.loads(r.stdout.strip().splitlines()[-1])
    ref, ref32 = both["f64"], both["f32"]

    iwae = _example("variational_autoencoder.iwae")
    iwae.device = torch.device("cuda")
    iwae.reparameterization = not vimco
    torch.manual_seed(0)
    gen, var = iwae.Generator(784, 40, K), iwae.Variational(784, 40, K)
    model = ImportanceWeightedObjective(gen, var, axis=0, estimator="vimco" if vimco else "sgvb").to(torch.device("cuda"))
    e = torch.tensor(eps, device="cuda")
    # reparameterised: eps is the unit noise; not reparameterised: torch.normal(mean, std) == mean + std * eps
    with _rng.inject(normal=[e, e]):
        loss = model({"x": torch.tensor(x, device="cuda")})
    loss.backward()
    gap = {n: abs(ref32[n] - ref[n]) for n in ref}
    got = float(loss.detach())
    assert abs(got - ref["loss"]) <= max(2e-5 * abs(ref["loss"]), 1.5 * gap["loss"]), (got, ref["loss"], ref32["loss"])
    for n, p in model.named_parameters():
        got = float(p.grad.norm())
        assert abs(got - ref[n]) <= max(2e-4 * max(ref[n], 1e-6), 1.5 * gap[n]), (n, got, ref[n], ref32[n])
