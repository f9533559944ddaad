"""Small-shape calls of every kernel added or reworked in the second session, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
GPU box only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "zhusuan-pytorch_b200"))
import torch
from zhusuan import _backend as be

dev = "cuda"
torch.manual_seed(0)
FULL, KBCAST, SCALAR = be.FULL, be.KBCAST, be.SCALAR
for K, B, X in ((50, 9, 784), (7, 5, 128), (25, 160, 256), (33, 20, 512), (20, 6, 1024), (12, 4, 100)):
    probs = torch.sigmoid(2 * torch.randn(K, B, X, device=dev))
    x = (torch.rand(B, X, device=dev) < 0.5).float()
    other, logq = torch.randn(K, B, device=dev) - 55, torch.randn(K, B, device=dev) + 30
    for est in (be.SGVB, be.VIMCO):
        for impl in ("box", "boxg", "ring"):
            be.set_fused_impl({"box": be.IMPL_BOX, "boxg": be.IMPL_BOXG, "ring": be.IMPL_RING}[impl])
            be.iw_bernoulli_fused(est, probs, x, other, logq, 1.0 / B, want_logpx=True)
            be.iw_bernoulli_fused(est, probs, x, other, logq, 1.0 / B, need_dprobs=False)
        be.set_fused_impl(be.IMPL_DEFAULT)
        acc = torch.zeros(B, device=dev)
        be.iw_bernoulli_fused(est, probs, x, other, logq, 1.0 / B, out={"cost": acc}, accumulate_cost=True)
        for impl in (be.IMPL_BOX, be.IMPL_BOXG, be.IMPL_RING):  # the objective written by the launch itself
            be.set_fused_impl(impl)
            for _ in range(2):
                be.iw_bernoulli_fused(est, probs, x, other, logq, 1.0 / B, cost_scaled=True, want_loss=True)
        be.set_fused_impl(be.IMPL_DEFAULT)
        if be.fused_logits_supported(K, X, torch.float32):
            be.iw_bernoulli_fused(est, torch.randn(K, B, X, device=dev), x, other, logq, 1.0 / B, logits=True)
    l = torch.randn(K, B, X, device=dev)
    be.bernoulli_logpmf_fwd(x, KBCAST, l, FULL, K, B, X, logits=True)
    be.bernoulli_logpmf_bwd(torch.randn(K, B, device=dev), x, KBCAST, l, FULL, K, B, X, False, True, logits=True)
for n in (1, 7, 4097, 51200):
    mm, ls = torch.zeros(1, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    be.reinforce_step(torch.randn(n, device=dev), torch.randn(n, device=dev), mm, ls, 0.8)
state = torch.zeros(2, dtype=torch.int64, device=dev)  # device-side Philox position
for K, M, E in ((50, 37, 40), (3, 5, 8), (1, 9, 4), (6, 2, 132), (4, 3, 4600), (9, 77, 64)):
    mean, std = torch.randn(M, E, device=dev), torch.rand(M, E, device=dev) + 0.5
    for impl in (0, 1):  # lane-per-unit / row-per-thread forward
        be.set_latent_fwd_impl(impl)
        be.normal_latent_fwd(mean, std, KBCAST if K > 1 else FULL, K, M, E, seed=1, offset=4)
        be.normal_latent_fwd(mean, std, KBCAST if K > 1 else FULL, K, M, E, seed=1, rng_state=state, prior_mean=mean, prior_std=std)
        be.bernoulli_latent_fwd(torch.rand(M, E, device=dev).clamp(0.05, 0.95), KBCAST if K > 1 else FULL, K, M, E, seed=1, rng_state=state)
    be.set_latent_fwd_impl(-1)
    z, lq, lp = be.normal_latent_fwd(mean, std, KBCAST if K > 1 else FULL, K, M, E, seed=1, offset=4)
    be.normal_latent_bwd(lq, lp, torch.randn_like(z), z, mean, std, KBCAST if K > 1 else FULL, K, M, E, reparameterized=True)
    p = torch.rand(M, E, device=dev).clamp(0.05, 0.95)
    zb, lqb, lpb, bits = be.bernoulli_latent_fwd(p, KBCAST if K > 1 else FULL, K, M, E, seed=1, offset=8, want_bits=True)
    be.bernoulli_latent_bwd(lqb, zb, p, KBCAST if K > 1 else FULL, K, M, E)
    be.bernoulli_latent_bwd(lqb, zb, p, KBCAST if K > 1 else FULL, K, M, E, zbits=bits)
for K, E in ((5, 2048), (3, 4601), (2, 100003)):
    mean, y = torch.randn(K, 1, E, device=dev), torch.randn(1, E, device=dev)
    std = torch.full((1,), 0.3, device=dev)
    be.normal_logprob_fwd(y, KBCAST, mean, FULL, std, SCALAR, K, 1, E)
    be.normal_logprob_bwd(torch.randn(K, 1, device=dev), y, KBCAST, mean, FULL, std, SCALAR, K, 1, E, False, True, False)
for fam in (be.FAM_LOGISTIC, be.FAM_LAPLACE):
    K, M, E = 5, 6, 9
    loc, scale = torch.randn(M * E, device=dev), torch.rand(M * E, device=dev) + 0.5
    z = be.locscale_sample(fam, loc, KBCAST, scale, KBCAST, K, M * E, seed=3, offset=4)
    be.locscale_sample_bwd(fam, torch.randn_like(z), loc, KBCAST, scale, KBCAST, K, M * E, seed=3, offset=4)
    be.locscale_logprob_fwd(fam, z.reshape(K, M, E), FULL, loc.reshape(M, E), KBCAST, scale.reshape(M, E), KBCAST, K, M, E)
    be.locscale_logprob_bwd(fam, torch.randn(K, M, device=dev), z.reshape(K, M, E), FULL, loc.reshape(M, E), KBCAST,
                            scale.reshape(M, E), KBCAST, K, M, E, True, True, True)
# host step
K, B, X = 8, 300, 64
pin = lambda t: t.pin_memory()
probs = pin(torch.rand(K, B, X).clamp(0.01, 0.99)); x = pin((torch.rand(B, X) < 0.5).float())
cost, dprobs = pin(torch.empty(B)), pin(torch.empty(K, B, X))
ws = torch.empty(be.iw_step_host_workspace(K, B, X), dtype=torch.uint8, device=dev)
hs = be.HostStep(dev)
do, dq = torch.randn(K, B, device=dev) - 55, torch.randn(K, B, device=dev) + 30
dlp, dlq = torch.empty(K, B, device=dev), torch.empty(K, B, device=dev)
be.iw_step_host_begin(hs, be.SGVB, cost, dprobs, dlp, dlq, probs, x, do, dq, K, B, X, 1.0 / B, ws, True)
be.iw_step_host_wait(hs, 1)
torch.cuda.synchronize()
print("sanitize_small: all launches completed,", be.launch_count, "launches")
