"""Pins the CPU oracle (oracle/) against fixtures produced by the REAL reference
(tests/golden/make_golden.py) and against published known-answer vectors.  CPU only."""
import numpy as np
import pytest
from scipy import stats

F32, F64 = "f32", "f64"
# oracle and reference share the formula and op order; they differ by libm vs SLEEF ulps and by the
# order of the fp32 event sum, hence a few-ulp relative tolerance (scaled by the largest magnitude).
TOL = {F32: 3e-6, F64: 1e-12}


def close(a, b, dn, scale=1.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    tol = TOL[dn] * scale
    ref = max(np.abs(b).max(), 1e-30)
    np.testing.assert_allclose(a, b, rtol=tol, atol=tol * ref)


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert oracle.philox_kat([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox_kat([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox_kat([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_philox_streams(oracle):
    raw = oracle.philox_raw(16, seed=7, offset=3)
    assert raw.dtype == np.uint32 and len(set(raw.tolist())) == 16
    # element i uses counter i/4, word i%4
    assert oracle.philox_kat([2, 0, 3, 0], [7, 0]) == raw[8:12].tolist()
    u = oracle.philox_uniform(100000, seed=1, offset=0)
    assert 0.0 <= u.min() and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    n = oracle.philox_normal(100000, seed=1, offset=1)
    assert abs(n.mean()) < 0.02 and abs(n.std() - 1) < 0.02
    assert stats.kstest(n, "norm").pvalue > 1e-3


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("case,event", [("kbcast", 1), ("full", 1), ("ylik", 0), ("group2", 2)])
def test_normal_logprob(oracle, golden, case, event, dn):
    g = golden("normal_logprob")
    p = "%s_%s_" % (case, dn)
    x, mean, std, up = g[p + "x"], g[p + "mean"], g[p + "std"], g[p + "g"]
    if case == "ylik":
        K, M, E = mean.shape[0], mean.shape[1], 1
    elif case == "group2":
        K, M, E = x.shape[0], 1, x.shape[1] * x.shape[2]
    else:
        K, M, E = x.shape
    out = oracle.normal_logprob_fwd(x, mean, std, K, M, E)
    close(out.reshape(g[p + "out"].shape), g[p + "out"], dn)
    dx, dmean, dstd = oracle.normal_logprob_bwd(up, x, mean, std, K, M, E)
    close(dx, g[p + "dx"], dn, 4)
    close(dmean, g[p + "dmean"], dn, 4)
    close(dstd, g[p + "dstd"], dn, 4)


def test_normal_logprob_scipy(oracle):
    # the reference's own known-answer source: test/distributions/test_normal.py:92-126
    rng = np.random.RandomState(0)
    for shape in [(3,), (2, 3), (1, 5)]:
        mean, logstd, x = rng.standard_normal(shape), rng.standard_normal(shape), rng.standard_normal(shape)
        std = np.exp(logstd)
        n = int(np.prod(shape))
        out = oracle.normal_logprob_fwd(x, mean, std, 1, n, 1).reshape(shape)
        np.testing.assert_allclose(out, stats.norm.logpdf(x, mean, std), rtol=1e-6)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("case", ["lik", "lik_real", "latent"])
def test_bernoulli_logpmf(oracle, golden, case, dn):
    g = golden("bernoulli_logpmf")
    p = "%s_%s_" % (case, dn)
    x, probs, up = g[p + "x"], g[p + "probs"], g[p + "g"]
    K, M, E = (probs.shape if case != "latent" else x.shape)
    out = oracle.bernoulli_logpmf_fwd(x, probs, K, M, E)
    close(out, g[p + "out"], dn)
    dprobs = oracle.bernoulli_logpmf_bwd(up, x, probs, K, M, E)
    close(dprobs, g[p + "dprobs"], dn, 4)


def test_bernoulli_logpmf_scipy(oracle):
    # test/distributions/test_bernoulli.py:56-72 (rtol 1e-3 there because of the +1e-8 guard)
    rng = np.random.RandomState(0)
    logits = rng.standard_normal((2, 3))
    probs = 1 / (1 + np.exp(-logits))
    x = (rng.uniform(size=(2, 3)) < 0.5).astype(np.float64)
    out = oracle.bernoulli_logpmf_fwd(x, probs, 1, 6, 1).reshape(2, 3)
    np.testing.assert_allclose(out, stats.bernoulli.logpmf(x, probs), rtol=1e-6)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("shape", ["kb", "k50", "k1d", "dominant"])
@pytest.mark.parametrize("est", ["sgvb", "vimco"])
def test_iw_objectives(oracle, golden, shape, est, dn):
    g = golden("objectives")
    p = "%s_%s_%s_" % (shape, est, dn)
    logp, logq = g[p + "logp"], g[p + "logq"]
    code = oracle.SGVB if est == "sgvb" else oracle.VIMCO
    cost, dlp, dlq = oracle.iw_objective(code, logp, logq)
    close(cost.mean(), g[p + "loss"], dn)
    if est == "sgvb":
        close(cost.reshape(g[p + "cost"].shape), g[p + "cost"], dn)
    # The VIMCO signal is a difference of two O(|log w|) numbers (L - control variate,
    # importance_weighted_objective.py:187): in fp32 both the reference and the oracle carry
    # ~ulp(|log w|)/|signal| relative noise there, so f32 gradients are compared more loosely.
    scale = 2000 if (est == "vimco" and dn == F32) else 8
    close(dlp.reshape(logp.shape), g[p + "dlogp"], dn, scale)
    close(dlq.reshape(logp.shape), g[p + "dlogq"], dn, scale)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("shape", ["kb", "k50", "dominant"])
def test_log_mean_exp(oracle, golden, shape, dn):
    g = golden("objectives")
    p = "%s_lme_%s_" % (shape, dn)
    close(oracle.log_mean_exp(g[p + "x"]), g[p + "out"], dn)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "normal"), ("vimco", "bernoulli")])
def test_iw_path_composite(oracle, golden, est, latent, dn):
    """sample -> log q, log p(z), log p(x|z) -> objective -> gradients, against the reference run
    with the same injected noise (tests/golden/make_golden.py:gen_iw_path)."""
    g = golden("iw_path")
    dt = np.float32 if dn == F32 else np.float64
    K, B, Z, X = int(g["K"]), int(g["B"]), int(g["Z"]), int(g["X"])
    p = "%s_%s_%s_" % (est, latent, dn)
    probs, x = g["probs"].astype(dt), g["x"].astype(dt)
    if latent == "normal":
        mean, logstd, eps = g["mean"].astype(dt), g["logstd"].astype(dt), g["eps"].astype(dt)
        std = np.exp(logstd)
        z = oracle.normal_sample(mean, std, eps, K, B * Z).reshape(K, B, Z)
        logq = oracle.normal_logprob_fwd(z, mean, std, K, B, Z)
        logpz = oracle.normal_logprob_fwd(z, np.zeros((B, Z), dt), np.ones((B, Z), dt), K, B, Z)
    else:
        pq, u = g["probs_q"].astype(dt), g["u"].astype(dt)
        z = oracle.bernoulli_sample(pq, u, K, B * Z).reshape(K, B, Z)
        logq = oracle.bernoulli_logpmf_fwd(z, pq, K, B, Z)
        logpz = oracle.bernoulli_logpmf_fwd(z, np.full((B, Z), 0.5, dt), K, B, Z)
    close(z, g[p + "z"], dn)
    close(logq, g[p + "logq"], dn)
    close(logpz, g[p + "logpz"], dn)
    code = oracle.SGVB if est == "sgvb" else oracle.VIMCO
    r = oracle.iw_bernoulli_step(code, probs, x, logpz, logq)
    close(r["logpx"], g[p + "logpx"], dn)
    close(r["cost"].mean(), g[p + "loss"], dn)
    close(r["dprobs"], g[p + "dprobs"], dn, 30)
    # gradients of the variational parameters
    dlogp, dlogq = r["dlogp"], r["dlogq"]
    loose = 3000 if (est == "vimco" and dn == F32) else 30
    if latent == "normal":
        # d/dz through log p(z) (+ log q for the reparameterised estimator), then the pathwise sample grad
        dz_p, _, _ = oracle.normal_logprob_bwd(dlogp, z, np.zeros((B, Z), dt), np.ones((B, Z), dt), K, B, Z)
        dz_q, dmean_q, dstd_q = oracle.normal_logprob_bwd(dlogq, z, mean, std, K, B, Z)
        if est == "sgvb":
            dmean_s, dstd_s = oracle.normal_sample_bwd((dz_p + dz_q).reshape(K, B * Z), eps.reshape(K, B * Z),
                                                       mean, std, K, B * Z)
            dmean, dstd = dmean_q + dmean_s, dstd_q + dstd_s
        else:  # score-function estimator: the sample is detached
            dmean, dstd = dmean_q, dstd_q
        close(dmean, g[p + "da"], dn, loose)
        close(dstd * std, g[p + "db"], dn, loose)  # chain through std = exp(logstd)
    else:
        dpq = oracle.bernoulli_logpmf_bwd(dlogq, z, pq, K, B, Z)
        close(dpq, g[p + "da"], dn, loose)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("xn", ["binary", "real"])
def test_bernoulli_logits_node(oracle, golden, xn, dn):
    """Bernoulli(logits=...).log_prob and its gradient w.r.t. the logits, incl. saturated logits
    (tests/golden/make_golden.py:gen_logits_path, reference bernoulli.py:47-50,84-95)."""
    g = golden("logits_path")
    dt = np.float32 if dn == F32 else np.float64
    l, x, up = g["node_logits"].astype(dt), g["node_x_" + xn].astype(dt), g["node_g"].astype(dt)
    K, M, E = l.shape
    close(oracle.bernoulli_logits_logpmf_fwd(x, l, K, M, E), g["node_%s_%s_lp" % (xn, dn)], dn)
    close(oracle.bernoulli_logits_logpmf_bwd(up, x, l, K, M, E), g["node_%s_%s_dlogits" % (xn, dn)], dn, 30)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("est,latent", [("sgvb", "normal"), ("vimco", "bernoulli")])
def test_iw_logits_path_composite(oracle, golden, est, latent, dn):
    """The IW path with the likelihood given by logits, against the reference run with the same injected noise."""
    g = golden("logits_path")
    dt = np.float32 if dn == F32 else np.float64
    K, B, Z, X = int(g["K"]), int(g["B"]), int(g["Z"]), int(g["X"])
    p = "%s_%s_%s_" % (est, latent, dn)
    logits, x = g["logits"].astype(dt), g["x"].astype(dt)
    if latent == "normal":
        mean, logstd, eps = g["mean"].astype(dt), g["logstd"].astype(dt), g["eps"].astype(dt)
        std = np.exp(logstd)
        z = oracle.normal_sample(mean, std, eps, K, B * Z).reshape(K, B, Z)
        logq = oracle.normal_logprob_fwd(z, mean, std, K, B, Z)
        logpz = oracle.normal_logprob_fwd(z, np.zeros((B, Z), dt), np.ones((B, Z), dt), K, B, Z)
    else:
        pq, u = g["probs_q"].astype(dt), g["u"].astype(dt)
        z = oracle.bernoulli_sample(pq, u, K, B * Z).reshape(K, B, Z)
        logq = oracle.bernoulli_logpmf_fwd(z, pq, K, B, Z)
        logpz = oracle.bernoulli_logpmf_fwd(z, np.full((B, Z), 0.5, dt), K, B, Z)
    close(logq, g[p + "logq"], dn)
    close(logpz, g[p + "logpz"], dn)
    code = oracle.SGVB if est == "sgvb" else oracle.VIMCO
    r = oracle.iw_bernoulli_logits_step(code, logits, x, logpz, logq)
    close(r["logpx"], g[p + "logpx"], dn)
    close(r["cost"].mean(), g[p + "loss"], dn)
    close(r["dprobs"], g[p + "dlogits"], dn, 30)


@pytest.mark.parametrize("dn", [F32, F64])
@pytest.mark.parametrize("name", ["logistic", "laplace"])
def test_locscale_nodes(oracle, golden, name, dn):
    """Logistic / Laplace log_prob + gradients (parameters broadcast over particles, saturated z, x == loc) and the
    Logistic reparameterised sample with injected uniforms (tests/golden/make_golden.py:gen_locscale)."""
    g = golden("locscale")
    dt = np.float32 if dn == F32 else np.float64
    fam = oracle.LOGISTIC if name == "logistic" else oracle.LAPLACE
    x, loc, scale, up = (g[k].astype(dt) for k in ("x", "loc", "scale", "g"))
    K, M, E = x.shape
    p = "%s_%s_" % (name, dn)
    close(oracle.locscale_logprob_fwd(fam, x, loc, scale, K, M, E), g[p + "lp"], dn)
    dx, dloc, dscale = oracle.locscale_logprob_bwd(fam, up, x, loc, scale, K, M, E)
    close(dx.reshape(K, M, E), g[p + "dx"], dn, 30)
    close(dloc.reshape(M, E), g[p + "dloc"], dn, 30)
    close(dscale.reshape(M, E), g[p + "dscale"], dn, 30)
    if name == "logistic":
        u, dz = g["u"].astype(dt), g["dz"].astype(dt)
        close(oracle.locscale_sample(fam, loc, scale, u, K, M * E).reshape(K, M, E), g[p + "z"], dn)
        sl, ss = oracle.locscale_sample_bwd(fam, dz, u, K, M * E)
        close(sl.reshape(M, E), g[p + "sdloc"], dn, 30)
        close(ss.reshape(M, E), g[p + "sdscale"], dn, 30)


def _close_inf(a, ref, dn):
    """Compare arrays that may hold -inf (the x == high corner of the Uniform log-density)."""
    a, ref = np.asarray(a, np.float64), np.asarray(ref, np.float64)
    assert np.array_equal(np.isinf(a), np.isinf(ref)) and np.array_equal(a[np.isinf(a)], ref[np.isinf(ref)])
    close(a[np.isfinite(a)], ref[np.isfinite(ref)], dn)


@pytest.mark.parametrize("dn", [F32, F64])
def test_uniform_node(oracle, golden, dn):
    """Uniform log_prob (values on both boundaries: x == low inside, x == high -inf) + gradients, and the unit draw
    scaled by (high - low) (tests/golden/make_golden.py:gen_uniform; reference uniform.py:51-83)."""
    g = golden("uniform")
    dt = np.float32 if dn == F32 else np.float64
    x, low, high, up = (g[k].astype(dt) for k in ("x", "low", "high", "g"))
    K, M, E = x.shape
    _close_inf(oracle.locscale_logprob_fwd(oracle.UNIFORM, x, low, high, K, M, E), g[dn + "_lp"], dn)
    dx, dlow, dhigh = oracle.locscale_logprob_bwd(oracle.UNIFORM, up, x, low, high, K, M, E)
    assert not dx.any()
    close(dlow.reshape(M, E), g[dn + "_dlow"], dn, 30)
    close(dhigh.reshape(M, E), g[dn + "_dhigh"], dn, 30)
    u = g["u"].astype(dt)
    z = oracle.locscale_sample(oracle.UNIFORM, low, high - low, u, K, M * E).reshape(K, M, E)
    close(z, g[dn + "_rep_z"], dn)


def test_reinforce_steps(oracle, golden):
    """ELBO.reinforce over three consecutive calls: cost, gradients and the in-place moving-mean state
    (tests/golden/make_golden.py:gen_reinforce, reference elbo.py:200-238)."""
    g = golden("reinforce")
    mm, ls = 0.0, 0
    for step in range(3):
        p = "f32_s%d_" % step
        cost, dlp, dlq, mm, ls = oracle.reinforce_step(g[p + "logp"], g[p + "logq"], mm, ls, 0.8)
        close(cost, g[p + "loss"], F32)
        close(dlp, g[p + "dlogp"], F32)
        close(dlq, g[p + "dlogq"], F32, 30)
        close(mm, g[p + "moving_mean"], F32)
        assert ls == int(np.asarray(g[p + "local_step"]).reshape(-1)[0])


def _replay_sgmcmc(oracle, g, name):
    """Re-run the reference trajectory with the oracle's single-step updates and the recorded noise."""
    n, steps = int(g["n"]), int(g["steps"])
    unit = g["unit_noise"].astype(np.float32)
    calls = g[name + "_f32_calls"]
    w = g["x0"].astype(np.float32)
    grad = lambda w: (4 * w - 4 * w ** 3).astype(np.float32)  # d/dx (2x^2 - x^4)
    lr, alpha, beta = 0.01, 0.3, 0.02
    traj = []
    aux = np.zeros(n, np.float32)
    v = None
    for s in range(steps):
        ncalls = int((calls[:, 0] == s).sum())
        if name == "sgld":
            noise = (np.float32(np.sqrt(np.float64(np.float32(lr)))) * unit[s, 0]).astype(np.float32)
            w = oracle.sgld_step(w, grad(w), noise, lr)
        elif name == "psgld":
            w, aux = oracle.psgld_step(w, aux, grad(w), unit[s, 0], lr)
        else:
            second = name == "sghmc2"
            std_v = np.float32(np.sqrt(lr))
            std_n = np.float32(np.sqrt(2 * (alpha - beta) * lr))
            j = 0
            resample = False
            v_noise = None
            if v is None:  # SGHMC.py:26-27 lazily creates vs with the first draw of the first update
                v = (std_v * unit[s, j]).astype(np.float32)
                j += 1
            elif ncalls == 2:  # SGHMC.py:31-33 velocity resample
                resample, v_noise = True, (std_v * unit[s, j]).astype(np.float32)
                j += 1
            noise = (std_n * unit[s, j % 2]).astype(np.float32)
            w, v = oracle.sghmc_pre(w, v, v_noise, resample, second)
            w, v = oracle.sghmc_post(w, v, grad(w), noise, lr, alpha, second)
        traj.append(w.copy())
    return np.stack(traj)


@pytest.mark.parametrize("name", ["sgld", "psgld", "sghmc1", "sghmc2"])
def test_sgmcmc_trajectories(oracle, golden, name):
    g = golden("sgmcmc")
    traj = _replay_sgmcmc(oracle, g, name)
    # 4 chained fp32 updates; the quartic gradient amplifies 1-ulp differences slightly
    np.testing.assert_allclose(traj, g[name + "_f32_traj"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("dn", [F32, F64])
def test_categorical_pinned_to_torch_distributions(oracle, dn):
    """The reference has no Categorical (bn.py:8-19), so the pin is the implementation a reference user would reach for:
    torch.distributions.Categorical -- log_prob values and the autograd gradient w.r.t. the logits, parameters broadcast
    over particles and per particle."""
    import torch
    rng = np.random.RandomState(6)
    dt, tdt = (np.float32, torch.float32) if dn == F32 else (np.float64, torch.float64)
    for K, M, C, full in ((5, 33, 10, False), (4, 20, 100, True), (1, 9, 3, False)):
        logits = (2 * rng.standard_normal((K, M, C) if full else (M, C))).astype(dt)
        x = rng.randint(0, C, size=(K, M))
        up = rng.standard_normal((K, M)).astype(dt)
        lt = torch.tensor(logits, requires_grad=True)
        lp = torch.distributions.Categorical(logits=lt).log_prob(torch.tensor(x))
        (gl,) = torch.autograd.grad(lp, [lt], grad_outputs=torch.tensor(up))
        close(oracle.categorical_logpmf_fwd(x.astype(dt), logits, K, M, C), lp.detach().numpy(), dn, 10)
        close(oracle.categorical_logpmf_bwd(up, x.astype(dt), logits, K, M, C).reshape(logits.shape), gl.numpy(), dn, 30)


def test_categorical_closed_form(oracle):
    # parity UNPINNED (no Categorical in the reference): closed form vs scipy log_softmax
    from scipy.special import log_softmax
    rng = np.random.RandomState(5)
    K, M, C = 4, 6, 7
    logits = rng.standard_normal((M, C))
    x = rng.randint(0, C, size=(K, M)).astype(np.float64)
    out = oracle.categorical_logpmf_fwd(x, logits, K, M, C)
    ref = log_softmax(logits, -1)[np.arange(M)[None, :], x.astype(int)]
    np.testing.assert_allclose(out, ref, rtol=1e-12)
