"""API-level tests of the drop-in `zhusuan` package: the reference's public surface
(zhusuan.distributions / framework / variational / mcmc) with the reference's behaviour.

Every test runs in two modes:
  * `oracle` — CPU container: the ctypes backend is replaced by the CPU oracle (tests/oracle_backend.py)
    so the package's host logic is exercised without a GPU;
  * `cuda`   — B200 (`-m gpu`): the real kernels, CUDA tensors;
  * `cuda-host` — B200: CPU tensors in, results back on the CPU (the reference's tests feed CPU tensors).
Scenarios mirror the reference's test strategy: constructor errors, shapes, dtypes,
known answers from scipy, reparameterisation gradients, analytic-KL checks of the objectives,
golden fixtures produced by the real reference with injected noise, and the double-well sampler test.
"""
import math

import numpy as np
import pytest
import torch
from scipy import stats

import zhusuan
import zhusuan.mcmc
import zhusuan.variational
from zhusuan import _rng
from zhusuan.distributions import Bernoulli, Categorical, Normal
from zhusuan.framework import BayesianNet, StochasticTensor
from zhusuan.variational import ELBO, ImportanceWeightedObjective

MODES = [pytest.param("oracle"), pytest.param("cuda", marks=pytest.mark.gpu),
         pytest.param("cuda-host", marks=pytest.mark.gpu)]


@pytest.fixture(params=MODES)
def dev(request, monkeypatch):
    mode = request.param
    if mode == "oracle":
        import oracle_backend
        oracle_backend.install(monkeypatch)
        return torch.device("cpu")
    return torch.device("cuda") if mode == "cuda" else torch.device("cpu")


def T(a, dev, dtype=torch.float32, grad=False):
    t = torch.tensor(np.asarray(a), dtype=dtype, device=dev)
    t.requires_grad_(grad)
    return t


def close(a, ref, rtol=1e-5):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    ref = ref.detach().cpu().numpy() if torch.is_tensor(ref) else np.asarray(ref)
    a, ref = a.astype(np.float64), ref.astype(np.float64)
    np.testing.assert_allclose(a, ref, rtol=rtol, atol=rtol * max(np.abs(ref).max(), 1e-30))


# ============================================================================ distributions
def test_constructor_errors():
    with pytest.raises(ValueError, match="Either.*should be passed"):
        Normal(mean=0.)
    with pytest.raises(ValueError, match="Either.*should be passed"):
        Normal(mean=0., std=1., logstd=0.)
    with pytest.raises(RuntimeError):
        Normal(mean=torch.zeros(2, 3), std=torch.ones(4))
    with pytest.raises(ValueError, match="Either.*should be passed"):
        Bernoulli()
    with pytest.raises(TypeError, match="must have a dtype in"):
        Bernoulli(logits=torch.zeros(3, dtype=torch.int32))
    with pytest.raises(TypeError, match="must have a dtype in"):
        Normal(mean=torch.zeros(3, dtype=torch.float16), std=torch.ones(3, dtype=torch.float16))
    with pytest.raises(ValueError, match="non-negative"):
        Normal(mean=0., std=1., group_ndims=-1)
    # unknown kwargs are swallowed (reference base.py:77); check_numerics is accepted and ignored
    Normal(mean=0., std=1., check_numerics=True, reparameterize=True, reduce_mean_dims=[0])


def test_bernoulli_parameterisations():
    b = Bernoulli(0.)
    assert b.dtype == torch.float32 and float(b.probs) == 0.5
    p = torch.tensor([0.2, 0.7])
    close(Bernoulli(probs=p).logits, torch.log(p / (1 - p)))
    lg = torch.tensor([-1.0, 2.0], dtype=torch.float64)
    assert Bernoulli(logits=lg).dtype == torch.float64
    close(Bernoulli(logits=lg).probs, torch.sigmoid(lg))
    assert not Bernoulli(probs=p).is_reparameterized and Normal(0., 1.).is_reparameterized


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_normal_shapes_and_values(dev, dtype):
    # sample shapes, incl. the reference's broadcasting case mean [1,3] + std [2,1], n=2 -> [2,2,3]
    n = Normal(mean=torch.zeros(1, 3, dtype=dtype, device=dev), std=torch.ones(2, 1, dtype=dtype, device=dev))
    assert tuple(n.batch_shape) == (2, 3)
    assert tuple(n.sample(2).shape) == (2, 2, 3) and tuple(n.sample().shape) == (2, 3)
    n = Normal(mean=torch.zeros(4, 5, dtype=dtype, device=dev), logstd=torch.zeros(4, 5, dtype=dtype, device=dev))
    s = n.sample(7)
    assert tuple(s.shape) == (7, 4, 5) and s.dtype == dtype and s.device.type == dev.type
    assert tuple(n.sample(1).shape) == (4, 5) and tuple(n.sample(None).shape) == (4, 5)  # no particle axis
    assert n.sample_cache is not None
    # log_prob broadcasting and known answers (scipy, as test/distributions/test_normal.py:92-126)
    rng = np.random.RandomState(0)
    mean, logstd, x = rng.standard_normal((2, 3)), 0.5 * rng.standard_normal((2, 3)), rng.standard_normal((5, 2, 3))
    n = Normal(mean=T(mean, dev, dtype), logstd=T(logstd, dev, dtype))
    lp = n.log_prob(T(x, dev, dtype))
    assert tuple(lp.shape) == (5, 2, 3) and lp.dtype == dtype
    close(lp, stats.norm.logpdf(x, mean, np.exp(logstd)), 1e-5 if dtype == torch.float32 else 1e-7)
    close(n.prob(T(x, dev, dtype)), stats.norm.pdf(x, mean, np.exp(logstd)), 1e-5)
    g1 = Normal(mean=T(mean, dev, dtype), logstd=T(logstd, dev, dtype), group_ndims=1)
    close(g1.log_prob(T(x, dev, dtype)), stats.norm.logpdf(x, mean, np.exp(logstd)).sum(-1), 1e-5)
    g2 = Normal(mean=T(mean, dev, dtype), logstd=T(logstd, dev, dtype), group_ndims=2)
    assert tuple(g2.log_prob(T(x, dev, dtype)).shape) == (5,)
    close(n.logstd, logstd, 1e-5)


def test_normal_reparameterisation_gradients(dev):
    """test/distributions/test_normal.py:65-84: gradients reach the parameters through the sample
    iff the distribution is reparameterised."""
    mean = torch.zeros(3, 4, device=dev, requires_grad=True)
    std = torch.ones(3, 4, device=dev, requires_grad=True)
    z = Normal(mean=mean, std=std).sample(6)
    gm, gs = torch.autograd.grad(z.sum() + (z * z).sum(), [mean, std])
    assert gm.abs().sum() > 0 and gs.abs().sum() > 0
    close(gm, (1 + 2 * z).sum(0))
    z2 = Normal(mean=mean, std=std, is_reparameterized=False).sample(6)
    assert not z2.requires_grad
    # the score-function path: log_prob of a detached sample still differentiates wrt the parameters
    lp = Normal(mean=mean, std=std, is_reparameterized=False).log_prob(z2)
    gm2, = torch.autograd.grad(lp.sum(), [mean])
    close(gm2, ((z2 - mean) / std ** 2).sum(0), 1e-4)


def test_normal_sampler_statistics(dev):
    mean = torch.tensor([[-2.0, 0.0, 3.0]], device=dev)
    std = torch.tensor([[0.5, 1.0, 2.0]], device=dev)
    s = Normal(mean=mean, std=std).sample(40000).cpu().numpy()[:, 0, :]
    for j in range(3):
        m, sd = float(mean[0, j]), float(std[0, j])
        assert abs(s[:, j].mean() - m) < 4 * sd / math.sqrt(40000) + 1e-3
        assert abs(s[:, j].std() / sd - 1) < 0.02
        assert stats.kstest((s[:, j] - m) / sd, "norm").pvalue > 1e-3
    assert abs(np.corrcoef(s[:, 0], s[:, 1])[0, 1]) < 0.03  # coordinates are independent


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_bernoulli_shapes_and_values(dev, dtype):
    rng = np.random.RandomState(1)
    logits = rng.standard_normal((2, 3))
    probs = 1 / (1 + np.exp(-logits))
    b = Bernoulli(logits=T(logits, dev, dtype))
    s = b.sample(5)
    assert tuple(s.shape) == (5, 2, 3) and s.dtype == dtype and tuple(b.sample().shape) == (2, 3)
    assert set(np.unique(s.cpu().numpy())) <= {0.0, 1.0}
    x = (rng.uniform(size=(4, 2, 3)) < 0.5).astype(np.float64)
    lp = b.log_prob(T(x, dev, dtype))
    assert tuple(lp.shape) == (4, 2, 3)
    close(lp, stats.bernoulli.logpmf(x, probs), 1e-5)  # test/distributions/test_bernoulli.py:56-72
    b1 = Bernoulli(probs=T(probs, dev, dtype), group_ndims=1)
    close(b1.log_prob(T(x, dev, dtype)), stats.bernoulli.logpmf(x, probs).sum(-1), 1e-5)
    # log_prob(None) evaluates at the cached sample (reference Q2)
    s = b.sample(3)
    close(b.log_prob(None), b.log_prob(s))
    # gradient wrt probs
    p = T(probs, dev, dtype, grad=True)
    (g,) = torch.autograd.grad(Bernoulli(probs=p).log_prob(T(x, dev, dtype)).sum(), [p])
    close(g, (x / (probs + 1e-8) - (1 - x) / (1 - probs + 1e-8)).sum(0), 1e-5)


def test_bernoulli_sampler_statistics(dev):
    p = torch.tensor([0.0, 0.1, 0.5, 0.9, 1.0], device=dev)
    s = Bernoulli(probs=p).sample(40000).cpu().numpy()
    f = s.mean(0)
    assert f[0] == 0 and f[4] == 1
    assert np.abs(f[1:4] - np.array([0.1, 0.5, 0.9])).max() < 0.01


def test_categorical(dev):
    rng = np.random.RandomState(2)
    logits = T(rng.standard_normal((3, 5)), dev, grad=True)
    c = Categorical(logits=logits)
    assert tuple(c.batch_shape) == (3,) and c.n_categories == 5 and not c.is_reparameterized
    s = c.sample(7)
    assert tuple(s.shape) == (7, 3) and s.dtype == torch.float32
    lp = c.log_prob(s)
    ref = torch.log_softmax(logits, -1).unsqueeze(0).expand(7, 3, 5).gather(-1, s.long().unsqueeze(-1)).squeeze(-1)
    close(lp, ref, 1e-5)
    (g,) = torch.autograd.grad(lp.sum(), [logits])
    (gr,) = torch.autograd.grad(ref.sum(), [logits])
    close(g, gr, 1e-5)
    freq = np.bincount(Categorical(logits=logits[0].detach()).sample(30000).cpu().numpy().astype(int), minlength=5)
    close(freq / 30000.0, torch.softmax(logits[0], -1), 0.03)
    with pytest.raises(ValueError, match="Either"):
        Categorical()


# ============================================================================ framework
class _Net(BayesianNet):
    def __init__(self, K, dev, dtype=torch.float32):
        super().__init__(device=dev)
        self.K, self.dt = K, dtype

    def forward(self, observed):
        self.observe(observed)
        d = self.device
        z = self.normal("z", mean=torch.zeros(4, 3, dtype=self.dt, device=d),
                        std=torch.ones(4, 3, dtype=self.dt, device=d), n_samples=self.K, reduce_sum_dims=[2])
        self.cache["z_seen"] = z
        self.bernoulli("x", probs=torch.sigmoid(z), reduce_sum_dims=[-1])
        return self


def test_bayesian_net_protocol(dev):
    net = _Net(5, dev)
    assert net(dict()) is net and set(net.nodes) == {"z", "x"} and isinstance(net.nodes["z"], StochasticTensor)
    z1, z2 = net.nodes["z"].tensor, net.nodes["z"].tensor
    assert tuple(z1.shape) == (5, 4, 3) and not torch.equal(z1, z2)  # unobserved: a NEW draw per access (Q1)
    assert tuple(net.nodes["z"].shape) == (5, 4, 3)
    zobs = torch.zeros(5, 4, 3, device=dev)
    net({"z": zobs})
    assert net.nodes["z"].tensor is zobs and net.observed["z"] is zobs and net.cache["z_seen"] is zobs
    assert net.nodes["z"].dist.sample_cache is zobs
    lp = net.nodes["z"].log_prob()
    assert tuple(lp.shape) == (5, 4)
    close(lp, torch.full((5, 4), 3 * -0.9189385332))
    xobs = torch.ones(4, 3, device=dev)
    net({"z": zobs, "x": xobs})
    lj = net.log_joint()
    assert tuple(lj.shape) == (5, 4)
    close(lj, torch.full((5, 4), 3 * -0.9189385332 + 3 * math.log(0.5)), 1e-5)
    assert net.log_joint(use_cache=True) is net._log_joint_cache
    # registered by name or by instance; unknown names fail loudly
    net.stochastic_node("Normal", "a", mean=0., std=1.)
    net.sn(Bernoulli(probs=torch.tensor(0.3)), "b", n_samples=4)
    assert tuple(net.nodes["b"].tensor.shape) == (4,)
    with pytest.raises(NotImplementedError):
        net.stochastic_node("Gamma", "g", alpha=1., beta=1.)
    with pytest.raises(ValueError):
        net.stochastic_node(3, "c")
    with pytest.raises(ValueError):
        net.normal(3, mean=0., std=1.)


@pytest.mark.parametrize("plan", [
    dict(group_ndims=0, reduce_mean_dims=None, reduce_sum_dims=[2]),
    dict(group_ndims=1, reduce_mean_dims=None, reduce_sum_dims=None),
    dict(group_ndims=0, reduce_mean_dims=[0], reduce_sum_dims=[2]),
    dict(group_ndims=0, reduce_mean_dims=[0, 1], reduce_sum_dims=None, multiplier=456),
    dict(group_ndims=2, reduce_mean_dims=[0], reduce_sum_dims=None),
    dict(group_ndims=0, reduce_mean_dims=[1], reduce_sum_dims=[0]),
    dict(group_ndims=0, reduce_mean_dims=None, reduce_sum_dims=[1, 2]),
    dict(group_ndims=0, reduce_mean_dims=None, reduce_sum_dims=None),
])
def test_stochastic_tensor_reductions(dev, plan):
    """StochasticTensor.log_prob = group_ndims sum -> mean dims -> sum dims -> squeeze -> multiplier
    (reference stochastic_tensor.py:160-181), whichever part of it the kernel absorbs."""
    rng = np.random.RandomState(3)
    mean, std, x = rng.standard_normal((4, 6)), np.exp(0.2 * rng.standard_normal((4, 6))), rng.standard_normal((5, 4, 6))
    net = BayesianNet(device=dev)
    net.observe({"w": T(x, dev)})
    kw = {k: v for k, v in plan.items() if k != "group_ndims"}
    net.normal("w", mean=T(mean, dev), std=T(std, dev), group_ndims=plan["group_ndims"], n_samples=5, **kw)
    got = net.nodes["w"].log_prob()
    ref = torch.tensor(stats.norm.logpdf(x, mean, std))
    g = plan["group_ndims"]
    if g:
        ref = ref.sum(list(range(-g, 0)))
    dims = []
    if plan.get("reduce_mean_dims"):
        ref = ref.mean(plan["reduce_mean_dims"], keepdim=True)
        dims += plan["reduce_mean_dims"]
    if plan.get("reduce_sum_dims"):
        ref = ref.sum(plan["reduce_sum_dims"], keepdim=True)
        dims += plan["reduce_sum_dims"]
    for d in sorted(dims, reverse=True):
        ref = ref.squeeze(d)
    if plan.get("multiplier"):
        ref = ref * plan["multiplier"]
    assert tuple(got.shape) == tuple(ref.shape)
    close(got, ref, 1e-5)


# ============================================================================ objectives
def _kl(m1, s1, m2, s2):
    return torch.log(s2 / s1) + (s1 ** 2 + (m1 - m2) ** 2) / (2 * s2 ** 2) - 0.5


class _GenNode(object):
    """Duck-typed node, as the reference's tests plant them (test_elbo.py:23-66, test_iw.py:23-66)."""

    def __init__(self, mean, std, dev):
        self.mean, self.std, self.dev, self.observed = mean, std, dev, {}

    def log_prob(self):
        return Normal(mean=self.mean, std=self.std).log_prob(self.observed["x"])


class _GenNet(BayesianNet):
    def __init__(self, mean, std, dev):
        super().__init__()
        self._nodes["test"] = _GenNode(mean, std, dev)

    def forward(self, observed):
        self._nodes["test"].observed = dict(observed)
        return self


class _VarNode(object):
    def __init__(self, samples, log_q):
        self.tensor, self._lq = samples, log_q

    def log_prob(self):
        return self._lq


class _VarNet(BayesianNet):
    def __init__(self, samples, log_q):
        super().__init__()
        self._nodes["x"] = _VarNode(samples, log_q)

    def forward(self, observed):
        return self


def test_elbo_against_analytic_kl(dev):
    """-ELBO of q=N(0,1) against p=N(m,s) equals KL(q||p); test/variational/test_elbo.py:78-121."""
    rng = np.random.RandomState(1)
    eps = T(rng.standard_normal(100000), dev)
    logq = T(stats.norm.logpdf(eps.cpu().numpy()), dev)
    for m, s in ((0., 1.), (2., 3.)):
        model = ELBO(_GenNet(torch.tensor(m, device=dev), torch.tensor(s, device=dev), dev), _VarNet(eps, logq))
        kl = float(_kl(torch.tensor(0.), torch.tensor(1.), torch.tensor(m), torch.tensor(s)))
        assert abs(float(model({})) - kl) < 1e-2
    # sgvb gradients wrt the variational parameters vs the analytic KL gradients
    mu = torch.tensor(2., device=dev, requires_grad=True)
    sigma = torch.tensor(3., device=dev, requires_grad=True)
    x = eps * sigma + mu
    logq = Normal(mean=mu, std=sigma).log_prob(x)
    for m, s in ((0., 1.), (2., 3.)):
        model = ELBO(_GenNet(torch.tensor(m, device=dev), torch.tensor(s, device=dev), dev), _VarNet(x, logq))
        grads = torch.autograd.grad(model({}), [mu, sigma], retain_graph=True)
        true = torch.autograd.grad(_kl(mu, sigma, torch.tensor(m, device=dev), torch.tensor(s, device=dev)), [mu, sigma])
        np.testing.assert_allclose([float(g) for g in grads], [float(g) for g in true], rtol=1e-2, atol=1e-2)
    with pytest.raises(NotImplementedError):
        ELBO(None, None, estimator="nope")


def test_iw_objective_properties(dev):
    """K=1 reduces to the ELBO; the bound tightens with K; VIMCO and SGVB gradients agree in
    expectation (test/variational/test_iw.py:78-175)."""
    rng = np.random.RandomState(1)
    n1 = T(rng.standard_normal((1, 1000)), dev)
    n3 = T(rng.standard_normal(10000), dev)
    for m, s in ((0., 1.), (2., 3.)):
        gm, gs = torch.tensor(m, device=dev), torch.tensor(s, device=dev)
        kl = float(_kl(torch.tensor(0.), torch.tensor(1.), torch.tensor(m), torch.tensor(s)))
        lb1 = -float(ImportanceWeightedObjective(_GenNet(gm, gs, dev), _VarNet(n1, T(stats.norm.logpdf(n1.cpu().numpy()), dev)),
                                                 axis=0)({}))
        assert abs(lb1 + kl) < 6e-2
        lb3 = -float(ImportanceWeightedObjective(_GenNet(gm, gs, dev), _VarNet(n3, T(stats.norm.logpdf(n3.cpu().numpy()), dev)),
                                                 axis=0)({}))
        assert lb3 > -kl - 1e-6
    mu = torch.tensor(2., device=dev, requires_grad=True)
    sigma = torch.tensor(3., device=dev, requires_grad=True)
    x = n3 * sigma + mu
    norm = Normal(mean=mu, std=sigma)
    logq = norm.log_prob(x)
    xv = (n3 * sigma + mu).detach()
    logqv = norm.log_prob(xv)
    for m, s, thr in ((0., 1., 1e-2), (2., 3., 1e-6)):
        gm, gs = torch.tensor(m, device=dev), torch.tensor(s, device=dev)
        sg = ImportanceWeightedObjective(_GenNet(gm, gs, dev), _VarNet(x, logq), axis=0, estimator="sgvb")
        vi = ImportanceWeightedObjective(_GenNet(gm, gs, dev), _VarNet(xv, logqv), axis=0, estimator="vimco")
        g_vi = [float(g) for g in torch.autograd.grad(vi({}), [mu, sigma], retain_graph=True)]
        g_sg = [float(g) for g in torch.autograd.grad(sg({}), [mu, sigma], retain_graph=True)]
        np.testing.assert_allclose(g_vi, g_sg, rtol=thr, atol=thr)
    with pytest.raises(ValueError, match="axis"):
        ImportanceWeightedObjective(None, None)
    with pytest.raises(NotImplementedError):
        ImportanceWeightedObjective(None, None, axis=0, estimator="nope")
    obj = ImportanceWeightedObjective(None, None, axis=0, estimator="vimco")
    with pytest.raises(ValueError, match="multi-sample"):
        obj.vimco(torch.zeros(1, 4, device=dev), torch.zeros(1, 4, device=dev))


@pytest.mark.parametrize("dn", ["f32", "f64"])
@pytest.mark.parametrize("shape", ["kb", "k50", "k1d", "dominant"])
def test_objective_methods_against_reference_goldens(dev, golden, shape, dn):
    """sgvb / vimco / ELBO.sgvb / log_mean_exp called on raw tensors, as recorded from the reference."""
    g = golden("objectives")
    dt = torch.float32 if dn == "f32" else torch.float64
    rt = 1e-5 if dn == "f32" else 1e-10
    for est in ("sgvb", "vimco"):
        p = "%s_%s_%s_" % (shape, est, dn)
        lp, lq = T(g[p + "logp"], dev, dt, True), T(g[p + "logq"], dev, dt, True)
        obj = ImportanceWeightedObjective(None, None, axis=0, estimator=est)
        loss = getattr(obj, est)(lp, lq, True)
        assert loss.dim() == 0 and loss.dtype == dt and loss.device.type == dev.type
        close(loss, g[p + "loss"], rt)
        dlp, dlq = torch.autograd.grad(loss, [lp, lq])
        ref = g if est == "sgvb" else {k.replace("f32", "f64") if False else k: v for k, v in g.items()}
        if est == "sgvb":
            close(dlp, g[p + "dlogp"], rt)
            close(dlq, g[p + "dlogq"], rt)
            close(obj.sgvb(lp, lq, False), g[p + "cost"], rt)
        else:  # compare with the float64 reference run (see tests/test_gpu_kernels.py on fp32 VIMCO noise)
            p64 = "%s_%s_f64_" % (shape, est)
            close(dlp, g[p64 + "dlogp"], rt)
            close(dlq, g[p64 + "dlogq"], rt)
    p = "%s_elbo_%s_" % (shape, dn)
    ps = "%s_sgvb_%s_" % (shape, dn)
    lp, lq = T(g[ps + "logp"], dev, dt, True), T(g[ps + "logq"], dev, dt, True)
    loss = ELBO(None, None).sgvb(lp, lq, True)
    close(loss, g[p + "loss"], rt)
    dlp, dlq = torch.autograd.grad(loss, [lp, lq])
    close(dlp, g[p + "dlogp"], rt)
    close(dlq, g[p + "dlogq"], rt)
    close(ELBO(None, None).sgvb(lp, lq, False), -(g[ps + "logp"] - g[ps + "logq"]), rt)
    p = "%s_lme_%s_" % (shape, dn)
    xx = T(g[p + "x"], dev, dt, True)
    out = zhusuan.log_mean_exp(xx, 0)
    close(out, g[p + "out"], rt)
    (dx,) = torch.autograd.grad(out.sum(), [xx])
    close(dx, g[p + "dx"], rt)
    assert tuple(zhusuan.log_mean_exp(xx, 0, keepdims=True).shape) == (1,) + tuple(xx.shape[1:])


def test_reinforce_against_reference_golden(dev, golden):
    g = golden("reinforce")
    elbo = ELBO(None, None, estimator="reinforce").to(dev)
    assert set(elbo.state_dict()) == {"moving_mean", "local_step"}
    for step in range(3):
        p = "f32_s%d_" % step
        lp, lq = T(g[p + "logp"], dev, grad=True), T(g[p + "logq"], dev, grad=True)
        loss = elbo.reinforce(lp, lq, True)
        close(loss, g[p + "loss"], 1e-5)
        dlp, dlq = torch.autograd.grad(loss, [lp, lq])
        close(dlp, g[p + "dlogp"], 1e-5)
        close(dlq, g[p + "dlogq"], 1e-4)
        close(elbo.moving_mean, g[p + "moving_mean"], 1e-5)
        assert int(elbo.local_step) == int(g[p + "local_step"])


class _PathGen(BayesianNet):
    def __init__(self, probs, K, latent, logits=False):
        super().__init__(device=probs.device)
        self.probs, self.K, self.latent, self.logits = probs, K, latent, logits

    def forward(self, observed):
        self.observe(observed)
        B, Z = self.observed["z"].shape[1:]
        dt, d = self.probs.dtype, self.probs.device
        if self.latent == "normal":
            self.normal("z", mean=torch.zeros([B, Z], dtype=dt, device=d), std=torch.ones([B, Z], dtype=dt, device=d),
                        is_reparameterized=False, n_samples=self.K, reduce_sum_dims=[2])
        else:
            self.bernoulli("z", probs=0.5 * torch.ones([B, Z], dtype=dt, device=d), n_samples=self.K,
                           reduce_sum_dims=[2])
        if self.logits:
            self.sn(Bernoulli(logits=self.probs), name="x", reduce_sum_dims=[2])
        else:
            self.sn(Bernoulli(probs=self.probs), name="x", reduce_sum_dims=[2])
        return self


class _PathVar(BayesianNet):
    def __init__(self, a, b, K, latent, reparam):
        super().__init__(device=a.device)
        self.a, self.b, self.K, self.latent, self.reparam = a, b, K, latent, reparam

    def forward(self, observed):
        self.observe(observed)
        if self.latent == "normal":
            self.sn(Normal(mean=self.a, logstd=self.b, is_reparameterized=self.reparam), name="z", n_samples=self.K,
                    reduce_sum_dims=[2])
        else:
            self.sn(Bernoulli(probs=self.a), name="z", n_samples=self.K, reduce_sum_dims=[2])
        return self


@pytest.mark.parametrize("dn", ["f32", "f64"])
@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "normal"), ("vimco", "bernoulli")])
def test_iw_path_against_reference_golden(dev, golden, est, latent, dn):
    """The whole hot path through the public API — sample, log q, log p(z), log p(x|z), objective,
    backward — against the reference run with the SAME injected noise (make_golden.py:gen_iw_path).
    `.tensor` is read twice per step in the reference (Q1), so the noise is injected twice."""
    g = golden("iw_path")
    dt = torch.float32 if dn == "f32" else torch.float64
    K = int(g["K"])
    p = "%s_%s_%s_" % (est, latent, dn)
    probs = T(g["probs"], dev, dt, True)
    if latent == "normal":
        a, b = T(g["mean"], dev, dt, True), T(g["logstd"], dev, dt, True)
        inj = dict(normal=[T(g["eps"], dev, dt)] * 2)
    else:
        a, b = T(g["probs_q"], dev, dt, True), None
        inj = dict(uniform=[T(g["u"], dev, dt)] * 2)
    gen, var = _PathGen(probs, K, latent), _PathVar(a, b, K, latent, est == "sgvb")
    obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
    with _rng.inject(**inj):
        loss = obj({"x": T(g["x"], dev, dt)})
    rt = 1e-5 if dn == "f32" else 1e-10
    close(loss, g[p + "loss"], rt)
    close(var.nodes["z"].dist.sample_cache, g[p + "z"], rt)
    close(var.nodes["z"].log_prob(), g[p + "logq"], rt)
    close(gen.nodes["z"].log_prob(), g[p + "logpz"], rt)
    close(gen.nodes["x"].log_prob(), g[p + "logpx"], rt)
    leaves = [probs, a] + ([b] if b is not None else [])
    grads = torch.autograd.grad(loss, leaves)
    # fp32: gradients go through exp(log w) with |log w| ~ 30: 4e-5 covers one ulp of the log-weights
    gt = 4e-5 if dn == "f32" else 1e-10
    ref = {k: g[(p if est == "sgvb" or dn == "f64" else p.replace("f32", "f64")) + k] for k in ("dprobs", "da")}
    close(grads[0], ref["dprobs"], gt)
    close(grads[1], ref["da"], gt)
    if b is not None:
        close(grads[2], g[(p if est == "sgvb" or dn == "f64" else p.replace("f32", "f64")) + "db"], gt)


@pytest.mark.parametrize("dn", ["f32", "f64"])
@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "bernoulli")])
def test_iw_logits_path_against_reference_golden(dev, golden, est, latent, dn):
    """The hot path with the likelihood given by LOGITS (Bernoulli(logits=...)): the sigmoid is applied inside the
    kernels and the gradient reaches the logits directly; against the reference run (make_golden.py:gen_logits_path)."""
    g = golden("logits_path")
    dt = torch.float32 if dn == "f32" else torch.float64
    K = int(g["K"])
    p = "%s_%s_%s_" % (est, latent, dn)
    logits = T(g["logits"], dev, dt, True)
    if latent == "normal":
        a, b = T(g["mean"], dev, dt, True), T(g["logstd"], dev, dt, True)
        inj = dict(normal=[T(g["eps"], dev, dt)] * 2)
    else:
        a, b = T(g["probs_q"], dev, dt, True), None
        inj = dict(uniform=[T(g["u"], dev, dt)] * 2)
    gen, var = _PathGen(logits, K, latent, logits=True), _PathVar(a, b, K, latent, est == "sgvb")
    obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
    with _rng.inject(**inj):
        loss = obj({"x": T(g["x"], dev, dt)})
    rt = 1e-5 if dn == "f32" else 1e-10
    close(loss, g[p + "loss"], rt)
    close(gen.nodes["x"].log_prob(), g[p + "logpx"], rt)
    leaves = [logits, a] + ([b] if b is not None else [])
    grads = torch.autograd.grad(loss, leaves)
    gt = 4e-5 if dn == "f32" else 1e-10
    pr = p if est == "sgvb" or dn == "f64" else p.replace("f32", "f64")
    close(grads[0], g[pr + "dlogits"], gt)
    close(grads[1], g[pr + "da"], gt)
    if b is not None:
        close(grads[2], g[pr + "db"], gt)
    # the node API: probs is still available (computed on demand), logits is what was passed
    d = Bernoulli(logits=logits)
    close(d.probs, 1.0 / (1.0 + np.exp(-g["logits"])), rt)
    assert d.logits is not None and tuple(d.batch_shape) == tuple(logits.shape)


@pytest.mark.parametrize("name", ["logistic", "laplace"])
def test_locscale_distributions_against_reference_golden(dev, golden, name):
    """zhusuan.distributions.{Logistic, Laplace}: log_prob with parameters broadcast over particles, gradients,
    the Logistic reparameterised sample with injected uniforms, properties and errors (reference logistic.py / laplace.py)."""
    from zhusuan.distributions import Laplace, Logistic
    g = golden("locscale")
    cls = Logistic if name == "logistic" else Laplace
    for dn, dt, rt in (("f32", torch.float32, 1e-5), ("f64", torch.float64, 1e-10)):
        x, loc, scale = T(g["x"], dev, dt, True), T(g["loc"], dev, dt, True), T(g["scale"], dev, dt, True)
        d = cls(loc=loc, scale=scale, group_ndims=1)
        assert d.is_reparameterized == (name == "logistic") and tuple(d.batch_shape) == tuple(loc.shape)
        lp = d.log_prob(x)
        p = "%s_%s_" % (name, dn)
        close(lp, g[p + "lp"], rt)
        gr = torch.autograd.grad(lp, [x, loc, scale], grad_outputs=T(g["g"], dev, dt))
        close(gr[0], g[p + "dx"], rt)
        close(gr[1], g[p + "dloc"], rt)
        close(gr[2], g[p + "dscale"], rt)
        K = g["x"].shape[0]
        if name == "logistic":
            with _rng.inject(uniform=[T(g["u"], dev, dt)]):
                z = cls(loc=loc, scale=scale).sample(K)
            close(z, g[p + "z"], rt)
            sg = torch.autograd.grad(z, [loc, scale], grad_outputs=T(g["dz"], dev, dt))
            close(sg[0], g[p + "sdloc"], rt)
            close(sg[1], g[p + "sdscale"], rt)
        else:
            z = cls(loc=loc, scale=scale).sample(K)
            assert tuple(z.shape) == (K,) + tuple(loc.shape) and not z.requires_grad
    with pytest.raises(RuntimeError):
        cls(loc=torch.zeros(2, 3), scale=torch.ones(4))
    if name == "logistic":
        with pytest.raises(ValueError, match="scale less than zero"):
            Logistic(loc=torch.zeros(3), scale=torch.tensor([1.0, 0.0, 2.0]))


def test_uniform_distribution_against_reference_golden(dev, golden):
    """zhusuan.distributions.Uniform incl. the reference's quirks (uniform.py:51-83): the reparameterised draw caches the
    UNSCALED unit draw (so log_prob(None) is evaluated there), the non-reparameterised draw is scaled twice, x == high
    is -inf, low >= high and out-of-support values raise ValueError (torch.distributions' argument validation)."""
    from zhusuan.distributions import Uniform
    g = golden("uniform")
    K = g["x"].shape[0]
    for dn, dt, rt in (("f32", torch.float32, 1e-5), ("f64", torch.float64, 1e-10)):
        low, high = T(g["low"], dev, dt, True), T(g["high"], dev, dt, True)
        d = Uniform(low=low, high=high, group_ndims=1)
        assert d.is_reparameterized and tuple(d.batch_shape) == tuple(low.shape)
        lp = d.log_prob(T(g["x"], dev, dt))
        ref = g[dn + "_lp"]
        got = lp.detach().cpu().numpy()
        assert np.array_equal(np.isinf(got), np.isinf(ref))
        close(got[np.isfinite(got)], ref[np.isfinite(ref)], rt)
        gr = torch.autograd.grad(lp, [low, high], grad_outputs=T(g["g"], dev, dt))
        close(gr[0], g[dn + "_dlow"], rt)
        close(gr[1], g[dn + "_dhigh"], rt)
        for name, reparam in (("rep", True), ("norep", False)):
            low, high = T(g["low"], dev, dt, True), T(g["high"], dev, dt, True)
            d = Uniform(low=low, high=high, is_reparameterized=reparam)
            with _rng.inject(uniform=[T(g["u"], dev, dt)]):
                z = d.sample(K)
            p = "%s_%s_" % (dn, name)
            close(z, g[p + "z"], rt)
            close(d.sample_cache, g[p + "cache"], rt)
            sg = torch.autograd.grad(z, [low, high], grad_outputs=T(g["dz"], dev, dt))
            close(sg[0], g[p + "dlow"], rt)
            close(sg[1], g[p + "dhigh"], rt)
            if reparam:
                close(d.log_prob(None), g[dn + "_rep_lp_cache"], rt)
    z = Uniform(low=torch.zeros(4, 8, device=dev), high=torch.ones(4, 8, device=dev)).sample(3)
    assert tuple(z.shape) == (3, 4, 8) and float(z.min()) >= 0.0 and float(z.max()) < 1.0
    with pytest.raises(ValueError):  # reference test_uniform.py:74-76
        Uniform(low=torch.tensor([10.0]), high=torch.tensor([2.0])).log_prob(torch.tensor([3.0]))
    with pytest.raises(ValueError):
        Uniform(low=torch.tensor([0.0]), high=torch.tensor([1.0])).log_prob(torch.tensor([1.5]))
    with pytest.raises(TypeError, match="must have a dtype in"):
        Uniform(2, 2, dtype=torch.int64)
    with pytest.raises(RuntimeError):
        Uniform(torch.zeros([2, 1]), torch.zeros([2, 4, 3]))

    class Net(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            self.uniform("u", low=torch.zeros(3, device=dev), high=torch.ones(3, device=dev), n_samples=4)
            self.stochastic_node("Uniform", "v", low=torch.zeros(3, device=dev), high=torch.ones(3, device=dev))
            return self

    net = Net(device=dev)({})
    assert type(net.nodes["u"].dist).__name__ == "Uniform" and tuple(net.nodes["u"].tensor.shape) == (4, 3)


def test_elbo_with_flow_against_reference_golden(dev, golden):
    """ELBO(transform=...) (elbo.py:90-119): the flow's outputs replace the latents, its log-determinants enter the
    objective as + sum(log_det) (:159-160) -- here folded into the objective's reduction (zs_combine_sums)."""
    g = golden("elbo_flow")
    K = int(g["K"])
    B, Z = g["mean"].shape
    for dn, dt, rt in (("f32", torch.float32, 2e-5), ("f64", torch.float64, 1e-10)):
        m, sd, sc, sh, w = (T(g[k], dev, dt, True) for k in ("mean", "std", "s", "t", "w"))

        class G(BayesianNet):
            def forward(self, observed):
                self.observe(observed)
                z = self.normal("z", mean=torch.zeros(B, Z, dtype=dt, device=dev), std=torch.ones(B, Z, dtype=dt, device=dev),
                                n_samples=K, reduce_sum_dims=[2])
                self.bernoulli("x", probs=torch.sigmoid(torch.matmul(z, w)), reduce_sum_dims=[2])
                return self

        class V(BayesianNet):
            def forward(self, observed):
                self.observe(observed)
                self.normal("z", mean=m, std=sd, n_samples=K, reduce_sum_dims=[2])
                return self

        def flow(inputs):
            (z,) = inputs
            return {"z": z * torch.exp(sc) + sh}, sc.sum().expand(z.shape[0], z.shape[1])

        eps = T(g["eps"], dev, dt)
        with _rng.inject(normal=[eps, eps]):
            loss = ELBO(G(device=dev), V(device=dev), transform=flow, transform_var=["z"])({"x": T(g["x"], dev, dt)})
        close(loss, g[dn + "_loss"], rt)
        grads = torch.autograd.grad(loss, [m, sd, sc, sh, w])
        for name, gr in zip(("dmean", "dstd", "ds", "dt", "dw"), grads):
            close(gr, g[dn + "_" + name], rt * 5)


def test_bn_logistic_builds_laplace_like_the_reference(dev):
    """framework/bn.py:336-352 of the reference: `bn.logistic` constructs a Laplace."""
    from zhusuan.distributions import Laplace

    class Net(BayesianNet):
        def forward(self, observed):
            self.observe(observed)
            self.logistic("a", loc=torch.zeros(3), scale=torch.ones(3))
            self.laplace("b", loc=torch.zeros(3), scale=torch.ones(3))
            self.stochastic_node("Logistic", "c", loc=torch.zeros(3), scale=torch.ones(3))
            return self

    net = Net()({})
    assert isinstance(net.nodes["a"].dist, Laplace) and isinstance(net.nodes["b"].dist, Laplace)
    assert type(net.nodes["c"].dist).__name__ == "Logistic"


@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "bernoulli")])
def test_fused_latent_route_equals_separate_kernels(dev, golden, est, latent):
    """A latent node whose reductions end in an event sum draws its sample and log q in ONE launch
    (distribution.sample_for_node); loss and gradients equal the separate sample / log-prob kernels' on the same
    injected noise, and the node protocol (sample_cache, log_prob at another value) is unchanged."""
    import zhusuan.distributions as D
    g = golden("iw_path")
    K = int(g["K"])
    res = {}
    for fused in (True, False):
        D.FUSED_LATENT = fused
        try:
            probs = T(g["probs"], dev, torch.float32, True)
            if latent == "normal":
                a, b = T(g["mean"], dev, torch.float32, True), T(g["logstd"], dev, torch.float32, True)
                inj = dict(normal=[T(g["eps"], dev)] * 2)
            else:
                a, b = T(g["probs_q"], dev, torch.float32, True), None
                inj = dict(uniform=[T(g["u"], dev)] * 2)
            gen, var = _PathGen(probs, K, latent), _PathVar(a, b, K, latent, est == "sgvb")
            obj = ImportanceWeightedObjective(gen, var, axis=0, estimator=est)
            with _rng.inject(**inj):
                loss = obj({"x": T(g["x"], dev)})
            node = var.nodes["z"]
            used = node.dist._logq_cache is not None
            leaves = [probs, a] + ([b] if b is not None else [])
            grads = torch.autograd.grad(loss, leaves)
            other = node.log_prob(torch.zeros_like(node.dist.sample_cache))  # not the cached sample: separate kernel
            res[fused] = (loss, grads, node.log_prob(), other, used)
        finally:
            D.FUSED_LATENT = True
    assert res[True][4] and not res[False][4]
    close(res[True][0], res[False][0].detach().cpu().numpy(), 2e-6)
    for ga, gb in zip(res[True][1], res[False][1]):
        close(ga, gb.detach().cpu().numpy(), 2e-5)
    close(res[True][2], res[False][2].detach().cpu().numpy(), 2e-6)
    close(res[True][3], res[False][3].detach().cpu().numpy(), 2e-6)


def test_plain_python_broadcast_shapes_matches_torch():
    """zhusuan._shapes.broadcast_shapes (used on the host path instead of torch.broadcast_shapes, which costs ~30 us
    per call) gives torch's result and torch's error type."""
    import itertools
    from zhusuan._shapes import broadcast_shapes
    dims = [(), (1,), (3,), (0,), (2, 1), (1, 3), (2, 3), (4, 1, 3), (1, 2, 1), (5, 2, 3)]
    for a, b in itertools.product(dims, dims):
        try:
            ref = torch.broadcast_shapes(a, b)
        except RuntimeError:
            ref = None
        if ref is None:
            with pytest.raises(RuntimeError):
                broadcast_shapes(a, b)
        else:
            out = broadcast_shapes(a, b)
            assert isinstance(out, torch.Size) and out == ref
    assert broadcast_shapes((2, 1), (1, 3), (4, 1, 1)) == torch.Size([4, 2, 3])


def test_upload_memo_shares_one_device_copy(dev):
    """Inside ops.upload_memo() a host tensor is moved to the compute device once and its gradient comes back through
    one edge; outside it every call uploads again.  (With the oracle backend `to_compute` is the identity.)"""
    from zhusuan import _ops
    t = torch.randn(4, 3, requires_grad=True)
    with _ops.upload_memo():
        a, b = _ops.to_compute(t), _ops.to_compute(t)
        assert a is b
        with _ops.upload_memo():  # nested contexts share the outer memo
            assert _ops.to_compute(t) is a
        t2 = t.detach().clone()
        t2.add_(1.0)  # a different tensor / version is a different entry
        assert _ops.to_compute(t2) is not a
    (a.sum() * 2 + b.sum()).backward()
    torch.testing.assert_close(t.grad, torch.full_like(t, 3.0))


def test_particle_linear_matches_reference_layer():
    """zhusuan.particle_linear == the repeat + matmul layer of the reference's BNN examples (bnn_vi.py:39-45), values
    and gradients, without materialising [K, B, n_out, n_in + 1]."""
    import zhusuan
    torch.manual_seed(3)
    K, B, n_in, n_out = 5, 7, 6, 4
    w = torch.randn(K, n_out, n_in + 1, dtype=torch.float64, requires_grad=True)
    h = torch.randn(K, B, n_in, dtype=torch.float64, requires_grad=True)

    def reference_layer(w, h):
        ww = torch.unsqueeze(w, 1).repeat([1, B, 1, 1])
        hh = torch.cat((h, torch.ones([*h.shape[:-1], 1], dtype=h.dtype)), -1)
        hh = torch.unsqueeze(hh, -1)
        p = torch.sqrt(torch.as_tensor(hh.shape[2], dtype=torch.float32))
        return torch.squeeze(torch.matmul(ww, hh) / p, -1)

    ref = reference_layer(w, h)
    out = zhusuan.particle_linear(w, h)
    torch.testing.assert_close(out, ref, rtol=1e-6, atol=1e-9)
    g = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, [w, h], g)
    go = torch.autograd.grad(out, [w, h], g)
    torch.testing.assert_close(go[0], gr[0], rtol=1e-6, atol=1e-9)
    torch.testing.assert_close(go[1], gr[1], rtol=1e-6, atol=1e-9)
    # 2-D activations are shared by all particles (the first layer of the examples)
    x = torch.randn(B, n_in, dtype=torch.float64)
    torch.testing.assert_close(zhusuan.particle_linear(w, x), reference_layer(w, x.unsqueeze(0).repeat(K, 1, 1)),
                               rtol=1e-6, atol=1e-9)
    with pytest.raises(RuntimeError):
        zhusuan.particle_linear(w, torch.randn(K, B, n_in + 2, dtype=torch.float64))


def test_elbo_path_against_reference_golden(dev, golden):
    """VAE ELBO (config 1 shapes, reduce_mean_dims=[0], reduce_sum_dims=[1]) vs the reference."""
    g = golden("elbo_path")
    B, Z = g["mean"].shape

    class G(BayesianNet):
        def __init__(self, probs):
            super().__init__(device=probs.device)
            self.probs = probs

        def forward(self, observed):
            self.observe(observed)
            d = self.probs.device
            self.normal("z", mean=torch.zeros([B, Z], device=d), std=torch.ones([B, Z], device=d),
                        reduce_mean_dims=[0], reduce_sum_dims=[1])
            self.bernoulli("x", probs=self.probs, reduce_mean_dims=[0], reduce_sum_dims=[1])
            return self

    class V(BayesianNet):
        def __init__(self, m, s):
            super().__init__(device=m.device)
            self.m, self.s = m, s

        def forward(self, observed):
            self.observe(observed)
            self.normal("z", mean=self.m, std=self.s, reduce_mean_dims=[0], reduce_sum_dims=[1])
            return self

    m, s, probs = T(g["mean"], dev, grad=True), T(g["std"], dev, grad=True), T(g["probs"], dev, grad=True)
    with _rng.inject(normal=[T(g["eps"], dev)] * 2):
        loss = ELBO(G(probs), V(m, s))({"x": T(g["x"], dev)})
    assert loss.dim() == 0
    close(loss, g["f32_loss"], 1e-5)
    dm, ds, dp = torch.autograd.grad(loss, [m, s, probs])
    close(dm, g["f32_dmean"], 1e-5)
    close(ds, g["f32_dstd"], 1e-5)
    close(dp, g["f32_dprobs"], 1e-5)


@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "bernoulli")])
def test_fused_route_matches_two_pass_route(dev, est, latent):
    """K >= 8 with a [K,B,X] Bernoulli likelihood takes the fused kernel; its loss and gradients must
    equal those of the general route (zhusuan.variational.FUSED = False) on the same noise."""
    if dev.type == "cpu" and torch.cuda.is_available():
        pytest.skip("the fused route needs device-resident probabilities")
    rng = np.random.RandomState(5)
    K, B, Z, X = 10, 7, 4, 16
    x = T((rng.uniform(size=(B, X)) < 0.5).astype(np.float32), dev)
    results = []
    for fused in (True, False):
        zhusuan.variational.FUSED = fused
        try:
            probs = T(1 / (1 + np.exp(-rng.RandomState(6).standard_normal((K, B, X)))) if False else
                      1 / (1 + np.exp(-np.random.RandomState(6).standard_normal((K, B, X)))), dev, grad=True)
            if latent == "normal":
                a = T(0.3 * np.random.RandomState(7).standard_normal((B, Z)), dev, grad=True)
                b = T(0.1 * np.random.RandomState(8).standard_normal((B, Z)), dev, grad=True)
                inj = dict(normal=[T(np.random.RandomState(9).standard_normal((K, B, Z)), dev)] * 2)
            else:
                a, b = T(np.random.RandomState(7).uniform(0.2, 0.8, size=(B, Z)), dev, grad=True), None
                inj = dict(uniform=[T(np.random.RandomState(9).uniform(size=(K, B, Z)), dev)] * 2)
            obj = ImportanceWeightedObjective(_PathGen(probs, K, latent), _PathVar(a, b, K, latent, est == "sgvb"),
                                              axis=0, estimator=est)
            with _rng.inject(**inj):
                loss = obj({"x": x})
            scale = 1.0 if fused else 1.0
            (loss * 3.0).backward()  # a non-unit upstream gradient exercises zs_scale_inplace
            results.append((loss.detach(), probs.grad.clone(), a.grad.clone(), None if b is None else b.grad.clone()))
        finally:
            zhusuan.variational.FUSED = True
    for u, v in zip(*results):
        if u is not None:
            close(u, v, 2e-5)


def test_fused_backward_runs_once(dev):
    if dev.type == "cpu" and torch.cuda.is_available():
        pytest.skip("the fused route needs device-resident probabilities")
    K, B, Z, X = 8, 3, 2, 8
    probs = torch.rand(K, B, X, device=dev).clamp(0.05, 0.95).requires_grad_()
    a = torch.zeros(B, Z, device=dev, requires_grad=True)
    b = torch.zeros(B, Z, device=dev, requires_grad=True)
    obj = ImportanceWeightedObjective(_PathGen(probs, K, "normal"), _PathVar(a, b, K, "normal", True), axis=0)
    loss = obj({"x": torch.ones(B, X, device=dev)})
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="once"):
        loss.backward()


def test_vimco_rejects_reparameterised_latents(dev):
    K, B, Z, X = 4, 3, 2, 8
    probs = torch.rand(K, B, X, device=dev)
    a, b = torch.zeros(B, Z, device=dev), torch.zeros(B, Z, device=dev)
    obj = ImportanceWeightedObjective(_PathGen(probs, K, "normal"), _PathVar(a, b, K, "normal", True), axis=0,
                                      estimator="vimco")
    with pytest.raises(ValueError, match="is_reparameterized must be false"):
        obj({"x": torch.ones(B, X, device=dev)})


# ============================================================================ SG-MCMC
class _Well(BayesianNet):
    """2x^2 - x^4 double well over one latent (test/mcmc/test_mcmc.py:24-37)."""

    def __init__(self, x0, noise_std=0.0):
        super().__init__()
        self.nodes["x"] = type("N", (), {"tensor": x0})()
        self.noise_std = noise_std

    def forward(self, observed):
        self.observe(observed)
        return self

    def _log_joint(self, use_cache=False):
        x = self.observed["x"]
        res = 2 * torch.pow(x, 2) - torch.pow(x, 4)
        if self.noise_std:
            res = res + self.noise_std * torch.randn_like(x)
        return res.sum()


def _sampler(name):
    m = zhusuan.mcmc
    return {"sgld": lambda: m.SGLD(learning_rate=0.01), "psgld": lambda: m.PSGLD(learning_rate=0.01),
            "sghmc1": lambda: m.SGHMC(learning_rate=0.01, n_iter_resample_v=2, friction=0.3, variance_estimate=0.02,
                                      second_order=False),
            "sghmc2": lambda: m.SGHMC(learning_rate=0.01, n_iter_resample_v=2, friction=0.3, variance_estimate=0.02,
                                      second_order=True)}[name]()


@pytest.mark.parametrize("name", ["sgld", "psgld", "sghmc1", "sghmc2"])
def test_sgmcmc_against_reference_trajectories(dev, golden, name):
    """Four updates of each sampler with the noise the reference consumed (make_golden.py:gen_sgmcmc)."""
    g = golden("sgmcmc")
    unit, calls = g["unit_noise"].astype(np.float32), g[name + "_f32_calls"]
    sampler = _sampler(name)
    x0 = T(g["x0"], dev, grad=True)
    model = _Well(x0)
    out = sampler.sample(model, {}, True)
    assert out["x"] is x0 and sampler.t == 1  # resample=True: no update, pre-detach draws (Q13)
    for s in range(int(g["steps"])):
        stds = calls[calls[:, 0] == s][:, 1]
        # the reference's torch.normal(mean=0, std=s, size) calls of this update, in order; PSGLD's
        # tensor-std call (recorded as -1) takes unit normals
        inj = [T(unit[s, j % 2] * (np.float32(sd) if sd >= 0 else np.float32(1.0)), dev) for j, sd in enumerate(stds)]
        with _rng.inject(normal=inj):
            w = sampler.sample(model, {}, False)["x"]
        assert w.requires_grad and w.is_leaf and w.device.type == dev.type
        np.testing.assert_allclose(w.detach().cpu().numpy(), g[name + "_f32_traj"][s], rtol=2e-5, atol=2e-6)
    assert sampler.t == 1 + int(g["steps"])


@pytest.mark.parametrize("name,bound", [("sgld", 0.023), ("psgld", 0.088), ("sghmc1", 0.016), ("sghmc2", 0.016)])
def test_sgmcmc_double_well(dev, name, bound):
    """100 chains on the double well with N(0,2) gradient noise; KDE of the samples against the true
    density, with the reference's own error bounds (test/mcmc/test_mcmc.py:74-107)."""
    if dev.type == "cpu" and torch.cuda.is_available():
        pytest.skip("host-shuttle mode is covered by the trajectory test")
    n_iters = 8000 if dev.type == "cuda" else 3000
    m = zhusuan.mcmc
    sampler = {"sgld": lambda: m.SGLD(learning_rate=0.01), "psgld": lambda: m.PSGLD(learning_rate=0.01),
               "sghmc1": lambda: m.SGHMC(learning_rate=0.01, n_iter_resample_v=50, friction=0.3,
                                         variance_estimate=0.02, second_order=False),
               "sghmc2": lambda: m.SGHMC(learning_rate=0.01, n_iter_resample_v=50, friction=0.3,
                                         variance_estimate=0.02, second_order=True)}[name]()
    torch.manual_seed(0)
    x = torch.zeros([100], device=dev, requires_grad=True)
    model = _Well(x, noise_std=2.0)
    samples = []
    burn = n_iters * 2 // 3
    for t in range(n_iters):
        xs = sampler.sample(model, {}, t == 0)["x"]
        if t >= burn and t % 50 == 0:
            samples.append(xs.detach().cpu().numpy().copy())
    samples = np.array(samples).reshape(-1)
    assert not np.isnan(samples).any()
    xs = np.linspace(-3, 3, 1000)
    pdf = np.exp(2 * xs ** 2 - xs ** 4)
    pdf = pdf / pdf.mean() / 6
    err = np.abs(stats.gaussian_kde(samples)(xs) - pdf).mean()
    assert err < (bound if dev.type == "cuda" else 2.5 * bound), err


def test_sghmc_multiple_latents(dev):
    """The reference raises for differently-shaped latents (shared gaussian_term, SGHMC.py:34); the
    B200 build draws one term per variable."""

    class Two(BayesianNet):
        def __init__(self, d):
            super().__init__(device=d)

        def forward(self, observed):
            self.observe(observed)
            d = self.device
            self.normal("a", mean=torch.zeros(3, 4, device=d), std=torch.ones(3, 4, device=d), n_samples=5,
                        group_ndims=2, reduce_mean_dims=[0])
            self.normal("b", mean=torch.zeros(2, device=d), std=torch.ones(2, device=d), n_samples=5, group_ndims=1,
                        reduce_mean_dims=[0])
            return self

    net = Two(dev)
    for second in (False, True):
        s = zhusuan.mcmc.SGHMC(learning_rate=1e-2, second_order=second)
        w = s.sample(net, {}, True)
        assert set(w) == {"a", "b"}
        for _ in range(3):
            w = s.sample(net, {}, False)
        assert tuple(w["a"].shape) == (5, 3, 4) and tuple(w["b"].shape) == (5, 2)
        assert torch.isfinite(w["a"]).all() and w["a"].requires_grad


# ============================================================================ BNN (configs 4 and 5)
class _BnnNet(BayesianNet):
    """The reference's BNN model body (examples/bayesian_neural_nets/bnn_vi.py:16-60, bnn_sgmcmc.py:16-70)
    against the public API: weight nodes with group_ndims=2, K particles, scalar-logstd likelihood on y,
    mean over particles and batch, multiplier."""

    def __init__(self, layer_sizes, K, y_logstd, dev, w_logstds=None):
        super().__init__(device=dev)
        self.layer_sizes, self.K, self.y_logstd, self.w_logstds, self.dev = layer_sizes, K, y_logstd, w_logstds, dev

    def forward(self, observed):
        self.observe(observed)
        x = self.observed["x"]
        d, dt = self.dev, x.dtype
        h = x.repeat([self.K, 1, 1])
        for i, (n_in, n_out) in enumerate(zip(self.layer_sizes[:-1], self.layer_sizes[1:])):
            kw = dict(std=torch.ones([n_out, n_in + 1], dtype=dt, device=d)) if self.w_logstds is None else \
                dict(logstd=self.w_logstds[i])
            w = self.normal("w" + str(i), mean=torch.zeros([n_out, n_in + 1], dtype=dt, device=d), group_ndims=2,
                            n_samples=self.K, reduce_mean_dims=[0], **kw)
            h = torch.cat([h, torch.ones([*h.shape[:-1], 1], dtype=dt, device=d)], -1)
            h = torch.einsum("kof,kbf->kbo", w, h) / math.sqrt(n_in + 1)
            if i < len(self.layer_sizes) - 2:
                h = torch.relu(h)
        self.normal("y", mean=h.squeeze(2), logstd=self.y_logstd, reduce_mean_dims=[0, 1], multiplier=456)
        return self


class _BnnVar(BayesianNet):
    def __init__(self, layer_sizes, K, means, logstds, dev):
        super().__init__(device=dev)
        self.layer_sizes, self.K, self.means, self.logstds = layer_sizes, K, means, logstds

    def forward(self, observed):
        self.observe(observed)
        for i in range(len(self.layer_sizes) - 1):
            self.normal("w" + str(i), mean=self.means[i], logstd=self.logstds[i], group_ndims=2, n_samples=self.K,
                        reduce_mean_dims=[0])
        return self


@pytest.mark.parametrize("dn", ["f32", "f64"])
def test_bnn_vi_against_reference_golden(dev, golden, dn):
    """Config 4 (bnn_vi.py) at a small size: ELBO with 5 weight particles, loss and every gradient."""
    g = golden("bnn")
    dt = torch.float32 if dn == "f32" else torch.float64
    K = int(g["K"])
    sizes = [g["x"].shape[1], g["w0_mean"].shape[0], 1]
    means = [T(g["w0_mean"], dev, dt, True), T(g["w1_mean"], dev, dt, True)]
    logstds = [T(g["w0_logstd"], dev, dt, True), T(g["w1_logstd"], dev, dt, True)]
    y_logstd = T(np.array([-0.3]), dev, dt, True)
    eps = [T(g["eps0"], dev, dt), T(g["eps1"], dev, dt)]
    elbo = ELBO(_BnnNet(sizes, K, y_logstd, dev), _BnnVar(sizes, K, means, logstds, dev))
    with _rng.inject(normal=[eps[0], eps[1], eps[0], eps[1]]):
        loss = elbo({"x": T(g["x"], dev, dt), "y": T(g["y"], dev, dt)})
    rt = 2e-5 if dn == "f32" else 1e-10
    p = "vi_%s_" % dn
    assert loss.dim() == 0
    close(loss, g[p + "loss"], rt)
    grads = torch.autograd.grad(loss, means + logstds + [y_logstd])
    for name, gr in zip(["dm0", "dm1", "ds0", "ds1", "dylogstd"], grads):
        close(gr, g[p + name], rt * 5)


def test_bnn_sgld_against_reference_golden(dev, golden):
    """Config 5 (bnn_sgmcmc.py) at a small size: SGLD over two differently shaped weight latents,
    5 chains, three updates with the reference's noise."""
    g = golden("bnn")
    K = int(g["K"])
    sizes = [g["x"].shape[1], g["w0_mean"].shape[0], 1]
    shapes = [tuple(g["w0_mean"].shape), tuple(g["w1_mean"].shape)]
    net = _BnnNet(sizes, K, T(np.array([-0.3]), dev), dev, w_logstds=[torch.zeros(s, device=dev) for s in shapes])
    sampler = zhusuan.mcmc.SGLD(learning_rate=1e-3)
    obs = {"x": T(g["x"], dev), "y": T(g["y"], dev)}
    init = [T(g["sgld_init_eps0"], dev), T(g["sgld_init_eps1"], dev)]
    # resample=True: forward draws once per node, then the sampler reads .tensor again (those draws are kept)
    with _rng.inject(normal=[init[0], init[1], init[0], init[1]]):
        w = sampler.sample(net, obs, True)
    close(w["w0"], g["sgld_w0_init"], 1e-6)
    close(w["w1"], g["sgld_w1_init"], 1e-6)
    std = np.float32(math.sqrt(float(torch.as_tensor(1e-3))))
    for s in range(3):
        inj = [T(g["sgld_unit0"][s].astype(np.float32) * std, dev), T(g["sgld_unit1"][s].astype(np.float32) * std, dev)]
        with _rng.inject(normal=inj):
            w = sampler.sample(net, obs, False)
        close(w["w0"], g["sgld_w0_traj"][s], 2e-5)
        close(w["w1"], g["sgld_w1_traj"][s], 2e-5)
        assert w["w0"].requires_grad and w["w0"].is_leaf
