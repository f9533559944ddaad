"""Generate tests/golden/*.npz by running the REAL reference (thuwzy/ZhuSuan-PyTorch).

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

Each fixture holds seeded inputs plus the reference's outputs and autograd gradients, in float32
and float64.  Noise is injected by patching torch.normal / torch.bernoulli, so the fixtures pin
the oracle (tests/test_oracle_golden.py) and the CUDA kernels (tests/test_gpu_parity.py) on
identical inputs.  Nothing here is imported by the product.
"""
import math
import os
import sys
from unittest import mock

import numpy as np
import torch

REF = os.environ.get("ZS_REFERENCE", "/root/reference")
if REF not in sys.path:
    sys.path.insert(0, REF)

import zhusuan  # noqa: E402  (the reference)
from zhusuan.distributions import Bernoulli, Normal, Laplace, Logistic  # noqa: E402
from zhusuan.framework import BayesianNet  # noqa: E402
from zhusuan.variational import ELBO, ImportanceWeightedObjective  # noqa: E402
from zhusuan import mcmc  # noqa: E402

assert os.path.realpath(zhusuan.__file__).startswith(os.path.realpath(REF)), zhusuan.__file__

OUT = os.path.dirname(os.path.abspath(__file__))
DT = {"f32": torch.float32, "f64": torch.float64}


def t(a, dtype, grad=False):
    x = torch.tensor(np.asarray(a), dtype=dtype)
    x.requires_grad_(grad)
    return x


def npy(x):
    return None if x is None else x.detach().cpu().numpy()


def save(name, **arrs):
    arrs = {k: v for k, v in arrs.items() if v is not None}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


# --------------------------------------------------------------------------- distributions
def gen_normal():
    rng = np.random.RandomState(11)
    cases = {
        # q(z|x): params [B,Z] broadcast over K particles, event sum over Z (iwae.py:102-120)
        "kbcast": dict(x=(5, 6, 8), mean=(6, 8), std=(6, 8), event=1),
        # prior on observed z with same-shaped params
        "full": dict(x=(4, 3, 8), mean=(4, 3, 8), std=(4, 3, 8), event=1),
        # BNN y-likelihood: mean [K,B], scalar std, y [B]  (bnn_vi.py:55-60)
        "ylik": dict(x=(7,), mean=(5, 7), std=(1,), event=0),
        # BNN weights: group_ndims=2 over [n_out, n_in+1]  (bnn_vi.py:32-38)
        "group2": dict(x=(6, 5, 9), mean=(5, 9), std=(5, 9), event=2),
    }
    out = {}
    for cname, c in cases.items():
        x64 = rng.standard_normal(c["x"])
        mean64 = 0.5 * rng.standard_normal(c["mean"])
        std64 = np.exp(0.3 * rng.standard_normal(c["std"]))
        for dn, dt in DT.items():
            x, mean, std = t(x64, dt, True), t(mean64, dt, True), t(std64, dt, True)
            dist = Normal(mean=mean, std=std, group_ndims=c["event"] if cname == "group2" else 0)
            lp = dist.log_prob(x)
            if cname in ("kbcast", "full"):
                lp = lp.sum(-1)
            g64 = np.random.RandomState(3).standard_normal(tuple(lp.shape))
            g = t(g64, dt)
            dx, dmean, dstd = torch.autograd.grad(lp, [x, mean, std], grad_outputs=g)
            p = "%s_%s_" % (cname, dn)
            out.update({p + "x": npy(x), p + "mean": npy(mean), p + "std": npy(std), p + "g": npy(g),
                        p + "out": npy(lp), p + "dx": npy(dx), p + "dmean": npy(dmean), p + "dstd": npy(dstd)})
    save("normal_logprob", **out)


def gen_bernoulli():
    rng = np.random.RandomState(12)
    cases = {
        # likelihood: probs [K,B,X], observed x [B,X] broadcast, event sum over X (iwae.py:73-81)
        "lik": dict(x=(6, 16), probs=(4, 6, 16), binary=True),
        # real-valued observations (load_mnist_realval)
        "lik_real": dict(x=(6, 16), probs=(4, 6, 16), binary=False),
        # Bernoulli latents: probs [B,Z] broadcast over K, samples [K,B,Z]
        "latent": dict(x=(5, 6, 8), probs=(6, 8), binary=True),
    }
    out = {}
    for cname, c in cases.items():
        probs64 = 1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal(c["probs"])))
        # exercise the +1e-8 guards: exact 0 / 1 and tiny probabilities
        flat = probs64.reshape(-1)
        flat[0], flat[1], flat[2], flat[3] = 0.0, 1.0, 1e-9, 1.0 - 1e-7
        x64 = (rng.uniform(size=c["x"]) < 0.5).astype(np.float64) if c["binary"] else rng.uniform(size=c["x"])
        if c["binary"]:
            # keep log(0 + 1e-8) finite-gradient corner but make x agree there so values stay finite
            pass
        for dn, dt in DT.items():
            probs, x = t(probs64, dt, True), t(x64, dt)
            dist = Bernoulli(probs=probs)
            lp = dist.log_prob(x).sum(-1)
            g = t(np.random.RandomState(4).standard_normal(tuple(lp.shape)), dt)
            (dprobs,) = torch.autograd.grad(lp, [probs], grad_outputs=g)
            p = "%s_%s_" % (cname, dn)
            out.update({p + "x": npy(x), p + "probs": npy(probs), p + "g": npy(g), p + "out": npy(lp),
                        p + "dprobs": npy(dprobs)})
    save("bernoulli_logpmf", **out)


# --------------------------------------------------------------------------- objectives
class _Dummy(torch.nn.Module):
    pass


def gen_objectives():
    rng = np.random.RandomState(13)
    out = {}
    shapes = {"kb": (10, 7), "k50": (50, 33), "k1d": (12,), "dominant": (6, 4)}
    for sname, shp in shapes.items():
        logp64 = -90.0 + 6.0 * rng.standard_normal(shp)
        logq64 = 25.0 + 3.0 * rng.standard_normal(shp)
        if sname == "dominant":
            logp64[2] += 60.0  # one particle carries ~all the weight: LOO arg-max branch
        for dn, dt in DT.items():
            for est in ("sgvb", "vimco"):
                logp, logq = t(logp64, dt, True), t(logq64, dt, True)
                obj = ImportanceWeightedObjective(_Dummy(), _Dummy(), axis=0, estimator=est)
                loss = getattr(obj, est)(logp, logq, True)
                dlp, dlq = torch.autograd.grad(loss, [logp, logq])
                p = "%s_%s_%s_" % (sname, est, dn)
                out.update({p + "logp": npy(logp), p + "logq": npy(logq), p + "loss": npy(loss),
                            p + "dlogp": npy(dlp), p + "dlogq": npy(dlq)})
                if est == "sgvb":
                    per = obj.sgvb(logp, logq, False)
                    out[p + "cost"] = npy(per)
            # ELBO.sgvb (elbo.py:134-161) and log_mean_exp (zhusuan/utils.py:6-21)
            logp, logq = t(logp64, dt, True), t(logq64, dt, True)
            elbo = ELBO(_Dummy(), _Dummy(), estimator="sgvb")
            loss = elbo.sgvb(logp, logq, True)
            dlp, dlq = torch.autograd.grad(loss, [logp, logq])
            p = "%s_elbo_%s_" % (sname, dn)
            out.update({p + "loss": npy(loss), p + "dlogp": npy(dlp), p + "dlogq": npy(dlq)})
            lw = t(logp64 - logq64, dt, True)
            lme = zhusuan.log_mean_exp(lw, 0)
            (dlw,) = torch.autograd.grad(lme.sum(), [lw])
            p = "%s_lme_%s_" % (sname, dn)
            out.update({p + "x": npy(lw), p + "out": npy(lme), p + "dx": npy(dlw)})
    save("objectives", **out)


def gen_reinforce():
    """ELBO.reinforce (elbo.py:163-238): three consecutive calls so the moving-mean state is pinned."""
    rng = np.random.RandomState(14)
    out = {}
    for dn, dt in DT.items():
        elbo = ELBO(_Dummy(), _Dummy(), estimator="reinforce")
        for step in range(3):
            logp64 = -90.0 + 6.0 * rng.standard_normal((10, 7))
            logq64 = 25.0 + 3.0 * rng.standard_normal((10, 7))
            logp, logq = t(logp64, torch.float32, True), t(logq64, torch.float32, True)
            loss = elbo.reinforce(logp, logq, True)
            dlp, dlq = torch.autograd.grad(loss, [logp, logq])
            p = "%s_s%d_" % (dn, step)
            out.update({p + "logp": npy(logp), p + "logq": npy(logq), p + "loss": npy(loss), p + "dlogp": npy(dlp),
                        p + "dlogq": npy(dlq), p + "moving_mean": npy(elbo.moving_mean.clone()),
                        p + "local_step": npy(elbo.local_step.clone())})
        break  # buffers are float32 in the reference regardless of input dtype
    save("reinforce", **out)


# --------------------------------------------------------------------------- end-to-end IW path
class _Gen(BayesianNet):
    """Generator with the decoder output given as a leaf (the path boundary of SURVEY.md §8d)."""

    def __init__(self, probs, K, latent, logits=False):
        super().__init__()
        self.probs, self.K, self.latent, self.logits = probs, K, latent, logits

    def forward(self, observed):
        self.observe(observed)
        B, Z = self.observed["z"].shape[1:]
        if self.latent == "normal":
            self.normal("z", mean=torch.zeros([B, Z], dtype=self.probs.dtype), std=torch.ones([B, Z], dtype=self.probs.dtype),
                        is_reparameterized=False, n_samples=self.K, reduce_sum_dims=[2])
        else:
            self.bernoulli("z", probs=0.5 * torch.ones([B, Z], dtype=self.probs.dtype), n_samples=self.K,
                           reduce_sum_dims=[2])
        if self.logits:  # the decoder's pre-activations: the reference applies the sigmoid itself (bernoulli.py:47-50)
            self.sn(Bernoulli(logits=self.probs), name="x", reduce_sum_dims=[2])
        else:
            self.sn(Bernoulli(probs=self.probs), name="x", reduce_sum_dims=[2])
        return self


class _Var(BayesianNet):
    def __init__(self, a, b, K, latent, reparam):
        super().__init__()
        self.a, self.b, self.K, self.latent, self.reparam = a, b, K, latent, reparam

    def forward(self, observed):
        self.observe(observed)
        if self.latent == "normal":
            self.sn(Normal(mean=self.a, logstd=self.b, is_reparameterized=self.reparam), name="z", n_samples=self.K,
                    reduce_sum_dims=[2])
        else:
            self.sn(Bernoulli(probs=self.a), name="z", n_samples=self.K, reduce_sum_dims=[2])
        return self


def gen_iw_path():
    rng = np.random.RandomState(15)
    K, B, Z, X = 6, 5, 4, 12
    out = {"K": np.int64(K), "B": np.int64(B), "Z": np.int64(Z), "X": np.int64(X)}
    mean64 = 0.5 * rng.standard_normal((B, Z))
    logstd64 = 0.3 * rng.standard_normal((B, Z))
    pq64 = 1.0 / (1.0 + np.exp(-rng.standard_normal((B, Z))))
    probs64 = 1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal((K, B, X))))
    x64 = (rng.uniform(size=(B, X)) < 0.5).astype(np.float64)
    eps64 = rng.standard_normal((K, B, Z))
    u64 = rng.uniform(size=(K, B, Z))
    out.update(mean=mean64, logstd=logstd64, probs_q=pq64, probs=probs64, x=x64, eps=eps64, u=u64)

    for dn, dt in DT.items():
        for est, latent in (("sgvb", "normal"), ("vimco", "normal"), ("vimco", "bernoulli")):
            probs = t(probs64, dt, True)
            if latent == "normal":
                a, b = t(mean64, dt, True), t(logstd64, dt, True)
            else:
                a, b = t(pq64, dt, True), None
            eps, u = t(eps64, dt), t(u64, dt)

            def fake_normal(*args, **kw):
                if "size" in kw:  # reparameterised draw: torch.normal(0., 1., size=shape)  (normal.py:104)
                    return eps.clone()
                m, s = args[0], args[1]  # non-reparameterised: torch.normal(mean, std)  (normal.py:102)
                return (m + s * eps).detach()

            def fake_bernoulli(p, *args, **kw):
                return (u < p).to(p.dtype)

            gen = _Gen(probs, K, latent)
            var = _Var(a, b, K, latent, reparam=(est == "sgvb"))
            obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
            with mock.patch("torch.normal", fake_normal), mock.patch("torch.bernoulli", fake_bernoulli):
                loss = obj({"x": t(x64, dt)})
            leaves = [probs, a] + ([b] if b is not None else [])
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            p = "%s_%s_%s_" % (est, latent, dn)
            out[p + "loss"] = npy(loss)
            out[p + "dprobs"] = npy(grads[0])
            out[p + "da"] = npy(grads[1])
            if b is not None:
                out[p + "db"] = npy(grads[2])
            out[p + "z"] = npy(var.nodes["z"].dist.sample_cache)
            out[p + "logq"] = npy(var.nodes["z"].log_prob())
            out[p + "logpz"] = npy(gen.nodes["z"].log_prob())
            out[p + "logpx"] = npy(gen.nodes["x"].log_prob())
    save("iw_path", **out)


# --------------------------------------------------------------------------- Bernoulli(logits=...)
def gen_logits_path():
    """Bernoulli given by logits: the node's log_prob + gradient, and the whole IW path with a logits likelihood
    at a row length the fused logits kernel is instantiated for (X = 128)."""
    rng = np.random.RandomState(23)
    out = {}
    # (a) node level, with saturated logits and a non-binary x case
    K, M, E = 5, 7, 20
    l64 = 3.0 * rng.standard_normal((K, M, E))
    l64.reshape(-1)[:6] = [0.0, 20.0, -20.0, 40.0, -40.0, 1e-4]
    xb64 = (rng.uniform(size=(M, E)) < 0.5).astype(np.float64)
    xr64 = rng.uniform(size=(M, E))
    g64 = rng.standard_normal((K, M))
    out.update(node_logits=l64, node_x_binary=xb64, node_x_real=xr64, node_g=g64)
    for dn, dt in DT.items():
        for xn, x64 in (("binary", xb64), ("real", xr64)):
            l = t(l64, dt, True)
            d = Bernoulli(logits=l, group_ndims=1)
            lp = d.log_prob(t(x64, dt))
            (dl,) = torch.autograd.grad(lp, [l], grad_outputs=t(g64, dt))
            out["node_%s_%s_lp" % (xn, dn)] = npy(lp)
            out["node_%s_%s_dlogits" % (xn, dn)] = npy(dl)
    # (b) IW path
    K, B, Z, X = 6, 5, 4, 128
    out.update(K=np.int64(K), B=np.int64(B), Z=np.int64(Z), X=np.int64(X))
    mean64 = 0.5 * rng.standard_normal((B, Z))
    logstd64 = 0.3 * rng.standard_normal((B, Z))
    pq64 = 1.0 / (1.0 + np.exp(-rng.standard_normal((B, Z))))
    logits64 = 2.0 * rng.standard_normal((K, B, X))
    x64 = (rng.uniform(size=(B, X)) < 0.5).astype(np.float64)
    eps64 = rng.standard_normal((K, B, Z))
    u64 = rng.uniform(size=(K, B, Z))
    out.update(mean=mean64, logstd=logstd64, probs_q=pq64, logits=logits64, x=x64, eps=eps64, u=u64)
    for dn, dt in DT.items():
        for est, latent in (("sgvb", "normal"), ("vimco", "bernoulli")):
            logits = t(logits64, dt, True)
            if latent == "normal":
                a, b = t(mean64, dt, True), t(logstd64, dt, True)
            else:
                a, b = t(pq64, dt, True), None
            eps, u = t(eps64, dt), t(u64, dt)

            def fake_normal(*args, **kw):
                if "size" in kw:
                    return eps.clone()
                m, s_ = args[0], args[1]
                return (m + s_ * eps).detach()

            def fake_bernoulli(p, *args, **kw):
                return (u < p).to(p.dtype)

            gen = _Gen(logits, K, latent, logits=True)
            var = _Var(a, b, K, latent, reparam=(est == "sgvb"))
            obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
            with mock.patch("torch.normal", fake_normal), mock.patch("torch.bernoulli", fake_bernoulli):
                loss = obj({"x": t(x64, dt)})
            leaves = [logits, a] + ([b] if b is not None else [])
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            p_ = "%s_%s_%s_" % (est, latent, dn)
            out[p_ + "loss"] = npy(loss)
            out[p_ + "dlogits"] = npy(grads[0])
            out[p_ + "da"] = npy(grads[1])
            if b is not None:
                out[p_ + "db"] = npy(grads[2])
            out[p_ + "logq"] = npy(var.nodes["z"].log_prob())
            out[p_ + "logpz"] = npy(gen.nodes["z"].log_prob())
            out[p_ + "logpx"] = npy(gen.nodes["x"].log_prob())
    save("logits_path", **out)


# --------------------------------------------------------------------------- Logistic / Laplace
def gen_locscale():
    """Logistic and Laplace nodes: log_prob + gradients with parameters broadcast over particles, and the Logistic
    reparameterised sample with injected uniforms (torch.nn.init.uniform_ patched)."""
    rng = np.random.RandomState(29)
    K, M, E = 5, 6, 8
    x64 = 2.0 * rng.standard_normal((K, M, E))
    x64.reshape(-1)[:3] = [0.0, 60.0, -60.0]
    loc64 = rng.standard_normal((M, E))
    loc64.reshape(-1)[0] = 0.0  # x == loc: the sign(0) corner of the Laplace gradient
    scale64 = np.exp(0.4 * rng.standard_normal((M, E)))
    g64 = rng.standard_normal((K, M))
    u64 = rng.uniform(0.02, 0.98, size=(K, M, E))
    dz64 = rng.standard_normal((K, M, E))
    out = dict(x=x64, loc=loc64, scale=scale64, g=g64, u=u64, dz=dz64)
    for dn, dt in DT.items():
        for name, cls in (("logistic", Logistic), ("laplace", Laplace)):
            x, loc, scale = t(x64, dt, True), t(loc64, dt, True), t(scale64, dt, True)
            d = cls(loc=loc, scale=scale, group_ndims=1)
            lp = d.log_prob(x)
            gr = torch.autograd.grad(lp, [x, loc, scale], grad_outputs=t(g64, dt))
            p = "%s_%s_" % (name, dn)
            out.update({p + "lp": npy(lp), p + "dx": npy(gr[0]), p + "dloc": npy(gr[1]), p + "dscale": npy(gr[2])})
        loc, scale = t(loc64, dt, True), t(scale64, dt, True)
        u = t(u64, dt)
        with mock.patch("torch.nn.init.uniform_", lambda tensor, a=0., b=1.: u.clone()):
            z = Logistic(loc=loc, scale=scale).sample(K)
        gr = torch.autograd.grad(z, [loc, scale], grad_outputs=t(dz64, dt))
        out.update({"logistic_%s_z" % dn: npy(z), "logistic_%s_sdloc" % dn: npy(gr[0]), "logistic_%s_sdscale" % dn: npy(gr[1])})
    save("locscale", **out)


# --------------------------------------------------------------------------- Uniform
def gen_uniform():
    """Uniform node: log_prob + gradients with parameters broadcast over particles (values on the boundaries included),
    and both sampling branches with injected unit uniforms (torch.rand patched), INCLUDING what the reference leaves in
    `sample_cache` and what `log_prob(None)` then evaluates -- the quirks of uniform.py:51-68."""
    from zhusuan.distributions.uniform import Uniform
    rng = np.random.RandomState(31)
    K, M, E = 5, 6, 8
    low64 = rng.standard_normal((M, E)) - 1.0
    span64 = np.exp(0.5 * rng.standard_normal((M, E))) + 1.5      # > 1: the unscaled u of a draw stays inside [low, high]
    low64 = np.minimum(low64, -0.05)                               # low < 0 < 1 < high  (u in [0,1) is in the support)
    high64 = np.maximum(low64 + span64, 1.05)
    t01 = rng.uniform(size=(K, M, E))
    x64 = low64 + t01 * (high64 - low64)
    x64.reshape(K, -1)[0, 0] = low64.reshape(-1)[0]                # x == low : inside  (-log(high - low))
    x64.reshape(K, -1)[1, 1] = high64.reshape(-1)[1]               # x == high: -inf   (log(0))
    g64 = rng.standard_normal((K, M))
    u64 = rng.uniform(size=(K, M, E))
    dz64 = rng.standard_normal((K, M, E))
    out = dict(x=x64, low=low64, high=high64, g=g64, u=u64, dz=dz64)
    for dn, dt in DT.items():
        low, high = t(low64, dt, True), t(high64, dt, True)
        d = Uniform(low=low, high=high, group_ndims=1)
        lp = d.log_prob(t(x64, dt))
        finite = torch.isfinite(lp)
        gr = torch.autograd.grad(lp, [low, high], grad_outputs=t(g64, dt))
        out.update({"%s_lp" % dn: npy(lp), "%s_dlow" % dn: npy(gr[0]), "%s_dhigh" % dn: npy(gr[1])})
        assert not bool(finite.all())  # the x == high corner really is -inf in the reference
        u = t(u64, dt)
        for name, reparam in (("rep", True), ("norep", False)):
            low, high = t(low64, dt, True), t(high64, dt, True)
            d = Uniform(low=low, high=high, is_reparameterized=reparam)
            with mock.patch("torch.rand", lambda *a, **k: u.clone()):
                z = d.sample(K)
            gr = torch.autograd.grad(z, [low, high], grad_outputs=t(dz64, dt))
            out.update({"%s_%s_z" % (dn, name): npy(z), "%s_%s_cache" % (dn, name): npy(d.sample_cache),
                        "%s_%s_dlow" % (dn, name): npy(gr[0]), "%s_%s_dhigh" % (dn, name): npy(gr[1])})
            if reparam:
                out["%s_rep_lp_cache" % dn] = npy(d.log_prob(None))  # evaluated at the UNSCALED draw
    save("uniform", **out)


# --------------------------------------------------------------------------- VAE ELBO (cfg 1 shapes, small)
def gen_elbo_path():
    rng = np.random.RandomState(16)
    B, Z, X = 6, 4, 12
    mean64 = 0.5 * rng.standard_normal((B, Z))
    std64 = np.exp(0.3 * rng.standard_normal((B, Z)))
    probs64 = 1.0 / (1.0 + np.exp(-2.0 * rng.standard_normal((B, X))))
    x64 = (rng.uniform(size=(B, X)) < 0.5).astype(np.float64)
    eps64 = rng.standard_normal((B, Z))
    out = dict(mean=mean64, std=std64, probs=probs64, x=x64, eps=eps64)

    class G(BayesianNet):
        def __init__(self, probs):
            super().__init__()
            self.probs = probs

        def forward(self, observed):
            self.observe(observed)
            dt = self.probs.dtype
            self.normal("z", mean=torch.zeros([B, Z], dtype=dt), std=torch.ones([B, Z], dtype=dt),
                        reduce_mean_dims=[0], reduce_sum_dims=[1])
            self.bernoulli("x", probs=self.probs, reduce_mean_dims=[0], reduce_sum_dims=[1])
            return self

    class V(BayesianNet):
        def __init__(self, m, s):
            super().__init__()
            self.m, self.s = m, s

        def forward(self, observed):
            self.observe(observed)
            self.normal("z", mean=self.m, std=self.s, reduce_mean_dims=[0], reduce_sum_dims=[1])
            return self

    for dn, dt in DT.items():
        m, s, probs = t(mean64, dt, True), t(std64, dt, True), t(probs64, dt, True)
        eps = t(eps64, dt)
        with mock.patch("torch.normal", lambda *a, **k: eps.clone()):
            loss = ELBO(G(probs), V(m, s))({"x": t(x64, dt)})
        dm, ds, dp = torch.autograd.grad(loss, [m, s, probs])
        out.update({dn + "_loss": npy(loss), dn + "_dmean": npy(dm), dn + "_dstd": npy(ds), dn + "_dprobs": npy(dp)})
    save("elbo_path", **out)


# --------------------------------------------------------------------------- ELBO with a flow (elbo.py:90-119,155-161)
def gen_elbo_flow():
    """ELBO(transform=...) with K particles: the variational sample goes through an elementwise affine flow
    z' = z * exp(s) + t whose log-determinant [K, B] enters the objective as  + sum(log_det)  (elbo.py:159-160)."""
    rng = np.random.RandomState(17)
    K, B, Z, X = 4, 5, 6, 8
    mean64 = 0.5 * rng.standard_normal((B, Z))
    std64 = np.exp(0.3 * rng.standard_normal((B, Z)))
    s64 = 0.2 * rng.standard_normal((Z,))
    t64 = 0.3 * rng.standard_normal((Z,))
    w64 = 0.4 * rng.standard_normal((Z, X))
    x64 = (rng.uniform(size=(B, X)) < 0.5).astype(np.float64)
    eps64 = rng.standard_normal((K, B, Z))
    out = dict(mean=mean64, std=std64, s=s64, t=t64, w=w64, x=x64, eps=eps64, K=np.int64(K))

    class G(BayesianNet):
        def __init__(self, w):
            super().__init__()
            self.w = w

        def forward(self, observed):
            self.observe(observed)
            dt = self.w.dtype
            z = self.normal("z", mean=torch.zeros([B, Z], dtype=dt), std=torch.ones([B, Z], dtype=dt), n_samples=K,
                            reduce_sum_dims=[2])
            self.bernoulli("x", probs=torch.sigmoid(torch.matmul(z, self.w)), reduce_sum_dims=[2])
            return self

    class V(BayesianNet):
        def __init__(self, m, s):
            super().__init__()
            self.m, self.s = m, s

        def forward(self, observed):
            self.observe(observed)
            self.normal("z", mean=self.m, std=self.s, n_samples=K, reduce_sum_dims=[2])
            return self

    for dn, dt in DT.items():
        m, sd, sc, sh, w = (t(a, dt, True) for a in (mean64, std64, s64, t64, w64))
        eps = t(eps64, dt)

        def flow(inputs):
            (z,) = inputs
            return {"z": z * torch.exp(sc) + sh}, sc.sum().expand(z.shape[0], z.shape[1])

        with mock.patch("torch.normal", lambda *a, **k: eps.clone()):
            loss = ELBO(G(w), V(m, sd), transform=flow, transform_var=["z"])({"x": t(x64, dt)})
        grads = torch.autograd.grad(loss, [m, sd, sc, sh, w])
        out[dn + "_loss"] = npy(loss)
        for name, gr in zip(("dmean", "dstd", "ds", "dt", "dw"), grads):
            out[dn + "_" + name] = npy(gr)
    save("elbo_flow", **out)


# --------------------------------------------------------------------------- SG-MCMC
class _Well(BayesianNet):
    """Quadratic-quartic log joint over one latent 'x' (test/mcmc/test_mcmc.py:24-37 without its noise)."""

    def __init__(self, x0):
        super().__init__()
        self.nodes["x"] = type("N", (), {"tensor": x0})()

    def forward(self, observed):
        self.observe(observed)
        return self

    def _log_joint(self, use_cache=False):
        x = self.observed["x"]
        return (2 * torch.pow(x, 2) - torch.pow(x, 4)).sum()


def gen_sgmcmc():
    rng = np.random.RandomState(17)
    n, steps = 64, 4
    out = {"n": np.int64(n), "steps": np.int64(steps)}
    x0_64 = 0.7 * rng.standard_normal(n)
    noise64 = rng.standard_normal((steps + 1, 2, n))  # per step: [velocity-resample, gaussian] unit draws
    out.update(x0=x0_64, unit_noise=noise64)
    samplers = {
        "sgld": lambda: mcmc.SGLD(learning_rate=0.01),
        "psgld": lambda: mcmc.PSGLD(learning_rate=0.01),
        "sghmc1": lambda: mcmc.SGHMC(learning_rate=0.01, n_iter_resample_v=2, friction=0.3, variance_estimate=0.02,
                                     second_order=False),
        "sghmc2": lambda: mcmc.SGHMC(learning_rate=0.01, n_iter_resample_v=2, friction=0.3, variance_estimate=0.02,
                                     second_order=True),
    }
    for sname, mk in samplers.items():
        for dn, dt in (("f32", torch.float32),):
            sampler = mk()
            x0 = t(x0_64, dt, True)
            model = _Well(x0)
            calls = []
            state = {"step": 0, "j": 0}

            def fake_normal(*args, **kw):
                # every torch.normal call of one update consumes the next unit draw of that step
                mean = kw.get("mean", args[0] if len(args) > 0 else 0.0)
                std = kw.get("std", args[1] if len(args) > 1 else 1.0)
                unit = torch.tensor(noise64[state["step"], state["j"] % 2], dtype=torch.float32)
                state["j"] += 1
                if torch.is_tensor(std):
                    r = mean + std * unit.to(std.dtype)
                else:
                    r = torch.tensor(mean, dtype=torch.float32) + torch.tensor(std, dtype=torch.float32) * unit
                calls.append((state["step"], float(std) if not torch.is_tensor(std) else -1.0))
                return r

            traj = []
            with mock.patch("torch.normal", fake_normal):
                sampler.sample(model, {}, True)  # resample=True: no update (SGMCMC.py:39-52)
                for s in range(steps):
                    state["step"], state["j"] = s, 0
                    w = sampler.sample(model, {}, False)["x"]
                    traj.append(npy(w).copy())
            out[sname + "_" + dn + "_traj"] = np.stack(traj)
            out[sname + "_" + dn + "_calls"] = np.array(calls, dtype=np.float64)
    save("sgmcmc", **out)


# --------------------------------------------------------------------------- BNN (configs 4 and 5, small)
class _BnnNet(BayesianNet):
    """The reference's BNN model body (examples/bayesian_neural_nets/bnn_vi.py:16-60 / bnn_sgmcmc.py:16-70),
    written against the public API: per-layer weight nodes with group_ndims=2 over [n_out, n_in+1],
    K particles, then a Normal likelihood on y with a scalar logstd, mean over particles and batch."""

    def __init__(self, layer_sizes, K, y_logstd, w_logstds=None):
        super().__init__()
        self.layer_sizes, self.K, self.y_logstd, self.w_logstds = layer_sizes, K, y_logstd, w_logstds

    def forward(self, observed):
        self.observe(observed)
        x = self.observed["x"]
        h = x.repeat([self.K, 1, 1])
        B = x.shape[0]
        for i, (n_in, n_out) in enumerate(zip(self.layer_sizes[:-1], self.layer_sizes[1:])):
            kw = dict(std=torch.ones([n_out, n_in + 1], dtype=x.dtype)) if self.w_logstds is None else \
                dict(logstd=self.w_logstds[i])
            w = self.normal("w" + str(i), mean=torch.zeros([n_out, n_in + 1], dtype=x.dtype), group_ndims=2,
                            n_samples=self.K, reduce_mean_dims=[0], **kw)
            h = torch.cat([h, torch.ones([*h.shape[:-1], 1], dtype=x.dtype)], -1)
            h = torch.einsum("kof,kbf->kbo", w, h) / math.sqrt(n_in + 1)
            if i < len(self.layer_sizes) - 2:
                h = torch.relu(h)
        y_mean = h.squeeze(2)
        self.normal("y", mean=y_mean, logstd=self.y_logstd, reduce_mean_dims=[0, 1], multiplier=456)
        return self


class _BnnVar(BayesianNet):
    def __init__(self, layer_sizes, K, means, logstds):
        super().__init__()
        self.layer_sizes, self.K, self.means, self.logstds = layer_sizes, K, means, logstds

    def forward(self, observed):
        self.observe(observed)
        for i in range(len(self.layer_sizes) - 1):
            self.normal("w" + str(i), mean=self.means[i], logstd=self.logstds[i], group_ndims=2, n_samples=self.K,
                        reduce_mean_dims=[0])
        return self


def gen_bnn():
    rng = np.random.RandomState(18)
    K, B, D, H = 5, 7, 6, 4
    sizes = [D, H, 1]
    shapes = [(H, D + 1), (1, H + 1)]
    x64 = rng.standard_normal((B, D))
    y64 = rng.standard_normal(B)
    means64 = [0.3 * rng.standard_normal(s) for s in shapes]
    logstds64 = [0.2 * rng.standard_normal(s) - 1.0 for s in shapes]
    eps64 = [rng.standard_normal((K,) + s) for s in shapes]
    out = dict(K=np.int64(K), x=x64, y=y64, w0_mean=means64[0], w1_mean=means64[1], w0_logstd=logstds64[0],
               w1_logstd=logstds64[1], eps0=eps64[0], eps1=eps64[1])
    for dn, dt in DT.items():
        # --- config 4: ELBO / SGVB with weight particles
        means = [t(m, dt, True) for m in means64]
        logstds = [t(s, dt, True) for s in logstds64]
        y_logstd = t(np.array([-0.3]), dt, True)
        queue = [t(e, dt) for e in eps64] * 2  # .tensor is read twice per node (stochastic_node + ELBO.forward)
        order = [0, 1, 0, 1]
        it = iter(order)

        def fake_normal(*args, **kw):
            return t(eps64[next(it)], dt)

        elbo = ELBO(_BnnNet(sizes, K, y_logstd), _BnnVar(sizes, K, means, logstds))
        with mock.patch("torch.normal", fake_normal):
            loss = elbo({"x": t(x64, dt), "y": t(y64, dt)})
        grads = torch.autograd.grad(loss, means + logstds + [y_logstd])
        p = "vi_%s_" % dn
        out[p + "loss"] = npy(loss)
        for name, g in zip(["dm0", "dm1", "ds0", "ds1", "dylogstd"], grads):
            out[p + name] = npy(g)
        # --- config 5: SGLD over the weight chains (prior logstd fixed), 3 updates
        if dn == "f32":
            net = _BnnNet(sizes, K, t(np.array([-0.3]), dt), w_logstds=[t(np.zeros(s), dt) for s in shapes])
            sampler = mcmc.SGLD(learning_rate=1e-3)
            noise = rng.standard_normal((4, 2) + (K,)).astype(np.float64)  # placeholder keeps the rng stream fixed
            draws = {"n": 0}
            sg_eps = [rng.standard_normal((K,) + s) for s in shapes] + [rng.standard_normal((K,) + s) for s in shapes]
            unit = [[rng.standard_normal((K,) + s) for s in shapes] for _ in range(3)]
            state = {"calls": []}

            def fake_normal2(*args, **kw):
                if "size" in kw and len(args) >= 2 and args[1] == 1.0:  # initial reparameterised draws
                    e = sg_eps[draws["n"]]
                    draws["n"] += 1
                    return t(e, dt)
                std = args[1]
                size = tuple(kw["size"])
                i = 0 if size == (K,) + shapes[0] else 1
                return t(unit[state["step"]][i], dt) * std

            obs = {"x": t(x64, dt), "y": t(y64, dt)}
            with mock.patch("torch.normal", fake_normal2):
                w = sampler.sample(net, obs, True)
                out["sgld_w0_init"], out["sgld_w1_init"] = npy(w["w0"]), npy(w["w1"])
                traj0, traj1 = [], []
                for s in range(3):
                    state["step"] = s
                    w = sampler.sample(net, obs, False)
                    traj0.append(npy(w["w0"]).copy())
                    traj1.append(npy(w["w1"]).copy())
            out["sgld_w0_traj"], out["sgld_w1_traj"] = np.stack(traj0), np.stack(traj1)
            out["sgld_unit0"] = np.stack([u[0] for u in unit])
            out["sgld_unit1"] = np.stack([u[1] for u in unit])
            out["sgld_init_eps0"], out["sgld_init_eps1"] = sg_eps[2], sg_eps[3]
    save("bnn", **out)


if __name__ == "__main__":
    torch.manual_seed(0)
    gen_normal()
    gen_bernoulli()
    gen_objectives()
    gen_reinforce()
    gen_iw_path()
    gen_logits_path()
    gen_locscale()
    gen_uniform()
    gen_elbo_path()
    gen_elbo_flow()
    gen_sgmcmc()
    gen_bnn()
