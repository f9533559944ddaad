"""ctypes binding of the sm_100a C-ABI library (include/zs_b200.h).

This is the only place the Python package touches native code.  There is NO CPU fallback: if
`lib/libzs_b200.so` is missing, or no CUDA device is present, every op raises.  Tensors are passed
as raw device pointers; all launches go to torch's current CUDA stream so stream / CUDA-graph
semantics of the caller hold.
"""
import ctypes
import os

import torch

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_ROOT, "lib", "libzs_b200.so")

F32, F64 = 0, 1
FULL, KBCAST, SCALAR = 0, 1, 2
SGVB, VIMCO, ELBO = 0, 1, 2
ERR_UNSUPPORTED, ERR_ALIGN = -6, -7

_lib = None
launch_count = 0  # kernels launched through this binding (bench.py's gpu_launches)


class BackendError(RuntimeError):
    pass


def _declare(lib):
    c = ctypes
    vp, i64, u64, i32, dbl = c.c_void_p, c.c_int64, c.c_uint64, c.c_int, c.c_double
    sig = {
        "zs_abi_version": (i32, []),
        "zs_strerror": (c.c_char_p, [i32]),
        "zs_last_error": (c.c_char_p, []),
        "zs_device_info": (i32, [c.POINTER(i32)] * 3),
        "zs_philox_uniform": (i32, [i32, vp, i64, u64, u64, vp]),
        "zs_philox_normal": (i32, [i32, vp, i64, dbl, dbl, u64, u64, vp]),
        "zs_philox_raw": (i32, [vp, i64, u64, u64, vp]),
        "zs_normal_sample": (i32, [i32, vp, vp, i32, vp, i32, vp, vp, i64, i64, u64, u64, vp]),
        "zs_normal_sample_bwd": (i32, [i32, vp, i32, vp, i32, vp, vp, i64, i64, u64, u64, vp]),
        "zs_normal_logprob_fwd": (i32, [i32, vp, vp, i32, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_normal_logprob_bwd": (i32, [i32, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_normal_latent_fwd": (i32, [i32, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, i64, i64, i64, u64, u64, vp]),
        "zs_bernoulli_latent_fwd": (i32, [i32, vp, vp, vp, vp, i32, vp, vp, i64, i64, i64, u64, u64, vp]),
        "zs_normal_latent_bwd": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, i32, vp, i32, vp, vp, i32, i64, i64, i64, vp]),
        "zs_bernoulli_latent_bwd": (i32, [i32, vp, vp, vp, vp, i32, i64, i64, i64, vp]),
        "zs_bernoulli_sample": (i32, [i32, vp, vp, i32, vp, i64, i64, u64, u64, vp]),
        "zs_bernoulli_logpmf_fwd": (i32, [i32, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_bernoulli_logpmf_bwd": (i32, [i32, vp, vp, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_locscale_sample": (i32, [i32, i32, vp, vp, i32, vp, i32, vp, i64, i64, u64, u64, vp]),
        "zs_locscale_sample_bwd": (i32, [i32, i32, vp, i32, vp, i32, vp, vp, i64, i64, u64, u64, vp]),
        "zs_locscale_logprob_fwd": (i32, [i32, i32, vp, vp, i32, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_locscale_logprob_bwd": (i32, [i32, i32, vp, vp, vp, vp, vp, i32, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_bernoulli_logits_logpmf_fwd": (i32, [i32, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_bernoulli_logits_logpmf_bwd": (i32, [i32, vp, vp, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_categorical_sample": (i32, [i32, vp, vp, i32, vp, i64, i64, i64, u64, u64, vp]),
        "zs_categorical_logpmf_fwd": (i32, [i32, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_categorical_logpmf_bwd": (i32, [i32, vp, vp, vp, i32, vp, i32, i64, i64, i64, vp]),
        "zs_iw_objective": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, i64, i64, dbl, vp]),
        "zs_log_mean_exp": (i32, [i32, vp, vp, i64, i64, vp]),
        "zs_log_mean_exp_bwd": (i32, [i32, vp, vp, vp, i64, i64, vp]),
        "zs_iw_bernoulli_fused_smem_bytes": (i64, [i64, i64]),
        "zs_iw_bernoulli_fused": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, dbl, vp]),
        "zs_iw_bernoulli_fused_logits": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, dbl, vp]),
        "zs_iw_bernoulli_fused_accumulate": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, dbl, vp]),
        "zs_scale_inplace": (i32, [i32, vp, i64, vp, vp]),
        "zs_debug_set_trace": (i32, [vp]),
        "zs_sgld_step": (i32, [i32, vp, vp, vp, vp, i64, dbl, u64, u64, vp]),
        "zs_psgld_step": (i32, [i32, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, u64, u64, vp]),
        "zs_sghmc_pre": (i32, [i32, vp, vp, vp, vp, i64, dbl, i32, i32, u64, u64, vp]),
        "zs_sghmc_post": (i32, [i32, vp, vp, vp, vp, vp, i64, dbl, dbl, dbl, i32, u64, u64, vp]),
        "zs_reinforce_step": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, i64, dbl, dbl, vp]),
        "zs_iw_step_host_workspace": (i64, [i64, i64, i64]),
        "zs_iw_step_host": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, dbl, vp, i64, vp]),
        "zs_iw_step_host_begin": (i32, [i32, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, dbl, vp, i64, i32, vp]),
        "zs_iw_step_host_wait": (i32, [i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTS = None


def load():
    """Load the library (once).  Raises BackendError if it has not been built."""
    global _lib, EXPORTS
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendError(
                "zhusuan (B200): %s not found. Build it with `python __graft_entry__.py build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        EXPORTS = _declare(lib)
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        lib = load()
        raise BackendError("%s failed: %s (%d) %s" % (what, lib.zs_strerror(rc).decode(), rc,
                                                      lib.zs_last_error().decode()))


def require_cuda():
    if not torch.cuda.is_available():
        raise BackendError("zhusuan (B200): no CUDA device available; the hot path has no CPU fallback")


def on_compute_device(t):
    """True when `t` already lives where the kernels run (no host<->device shuttle needed)."""
    return t.is_cuda


def dtype_code(dt):
    if dt == torch.float32:
        return F32
    if dt == torch.float64:
        return F64
    raise TypeError("zhusuan (B200) kernels support float32 / float64, got %s" % dt)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk_tensor(t, name, dtype=None):
    if t is None:
        return
    if not t.is_cuda:
        raise BackendError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise BackendError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s has dtype %s, expected %s" % (name, t.dtype, dtype))


def _count(n=1):
    global launch_count
    launch_count += n


# ----------------------------------------------------------------------------- RNG
def philox_raw(n, seed, offset, device):
    out = torch.empty(n, dtype=torch.int32, device=device)
    check(load().zs_philox_raw(_ptr(out), n, seed, offset, _stream()), "zs_philox_raw")
    _count()
    return out


def philox_uniform(n, dtype, seed, offset, device):
    out = torch.empty(n, dtype=dtype, device=device)
    check(load().zs_philox_uniform(dtype_code(dtype), _ptr(out), n, seed, offset, _stream()), "zs_philox_uniform")
    _count()
    return out


def philox_normal(n, dtype, mean, std, seed, offset, device):
    out = torch.empty(n, dtype=dtype, device=device)
    check(load().zs_philox_normal(dtype_code(dtype), _ptr(out), n, float(mean), float(std), seed, offset, _stream()),
          "zs_philox_normal")
    _count()
    return out


# ----------------------------------------------------------------------------- Normal
def normal_sample(mean, mean_mode, std, std_mode, K, N, eps_in=None, eps_out=None, seed=0, offset=0):
    dt = mean.dtype
    for t, n in ((mean, "mean"), (std, "std"), (eps_in, "eps_in"), (eps_out, "eps_out")):
        _chk_tensor(t, n, dt)
    z = torch.empty((K, N), dtype=dt, device=mean.device)
    check(load().zs_normal_sample(dtype_code(dt), _ptr(z), _ptr(mean), mean_mode, _ptr(std), std_mode, _ptr(eps_in),
                                  _ptr(eps_out), K, N, seed, offset, _stream()), "zs_normal_sample")
    _count()
    return z


def normal_sample_bwd(dz, mean_like, mean_mode, std_like, std_mode, K, N, eps=None, seed=0, offset=0,
                      need_mean=True, need_std=True):
    dt = dz.dtype
    _chk_tensor(dz, "dz", dt)
    _chk_tensor(eps, "eps", dt)
    dmean = torch.empty_like(mean_like) if need_mean else None
    dstd = torch.empty_like(std_like) if need_std else None
    check(load().zs_normal_sample_bwd(dtype_code(dt), _ptr(dmean), mean_mode, _ptr(dstd), std_mode, _ptr(dz), _ptr(eps),
                                      K, N, seed, offset, _stream()), "zs_normal_sample_bwd")
    _count()
    return dmean, dstd


def normal_logprob_fwd(x, xm, mean, mm, std, sm, K, M, E):
    dt = x.dtype
    for t, n in ((x, "x"), (mean, "mean"), (std, "std")):
        _chk_tensor(t, n, dt)
    out = torch.empty((K, M), dtype=dt, device=x.device)
    check(load().zs_normal_logprob_fwd(dtype_code(dt), _ptr(out), _ptr(x), xm, _ptr(mean), mm, _ptr(std), sm, K, M, E,
                                       _stream()), "zs_normal_logprob_fwd")
    _count()
    return out


def normal_logprob_bwd(g, x, xm, mean, mm, std, sm, K, M, E, need_x, need_mean, need_std):
    dt = x.dtype
    _chk_tensor(g, "g", dt)
    dx = torch.empty_like(x) if need_x else None
    dmean = torch.empty_like(mean) if need_mean else None
    dstd = torch.empty_like(std) if need_std else None
    check(load().zs_normal_logprob_bwd(dtype_code(dt), _ptr(dx), _ptr(dmean), _ptr(dstd), _ptr(g), _ptr(x), xm,
                                       _ptr(mean), mm, _ptr(std), sm, K, M, E, _stream()), "zs_normal_logprob_bwd")
    _count()
    return dx, dmean, dstd


# ----------------------------------------------------------------------------- fused latent nodes
def normal_latent_fwd(mean, std, mode, K, M, E, prior_mean=None, prior_std=None, eps_in=None, want_logq=True,
                      want_logp=True, seed=0, offset=0):
    """-> (z [K,M,E], logq [K,M] | None, logp [K,M] | None), or None when the shape is not supported."""
    dt = mean.dtype
    for t, n in ((mean, "mean"), (std, "std"), (prior_mean, "prior_mean"), (prior_std, "prior_std"), (eps_in, "eps")):
        _chk_tensor(t, n, dt)
    z = torch.empty((K, M, E), dtype=dt, device=mean.device)
    logq = torch.empty((K, M), dtype=dt, device=mean.device) if want_logq else None
    logp = torch.empty((K, M), dtype=dt, device=mean.device) if want_logp else None
    rc = load().zs_normal_latent_fwd(dtype_code(dt), _ptr(z), _ptr(logq), _ptr(logp), _ptr(mean), mode, _ptr(std), mode,
                                     _ptr(prior_mean), _ptr(prior_std), _ptr(eps_in), K, M, E, seed, offset, _stream())
    if rc in (ERR_UNSUPPORTED, ERR_ALIGN):
        return None
    check(rc, "zs_normal_latent_fwd")
    _count()
    return z, logq, logp


def normal_latent_bwd(dlogq, dlogp, dz_up, z, mean, std, mode, K, M, E, prior_mean=None, prior_std=None,
                      reparameterized=True):
    dt = z.dtype
    for t, n in ((dlogq, "dlogq"), (dlogp, "dlogp"), (dz_up, "dz_up"), (z, "z"), (mean, "mean"), (std, "std")):
        _chk_tensor(t, n, dt)
    dmean, dstd = torch.empty_like(mean), torch.empty_like(std)
    check(load().zs_normal_latent_bwd(dtype_code(dt), _ptr(dmean), _ptr(dstd), _ptr(dlogq), _ptr(dlogp), _ptr(dz_up),
                                      _ptr(z), _ptr(mean), mode, _ptr(std), mode, _ptr(prior_mean), _ptr(prior_std),
                                      int(bool(reparameterized)), K, M, E, _stream()), "zs_normal_latent_bwd")
    _count()
    return dmean, dstd


def bernoulli_latent_fwd(probs, mode, K, M, E, prior_probs=None, u_in=None, want_logq=True, want_logp=True, seed=0,
                         offset=0):
    dt = probs.dtype
    for t, n in ((probs, "probs"), (prior_probs, "prior_probs"), (u_in, "u_in")):
        _chk_tensor(t, n, dt)
    z = torch.empty((K, M, E), dtype=dt, device=probs.device)
    logq = torch.empty((K, M), dtype=dt, device=probs.device) if want_logq else None
    logp = torch.empty((K, M), dtype=dt, device=probs.device) if want_logp else None
    rc = load().zs_bernoulli_latent_fwd(dtype_code(dt), _ptr(z), _ptr(logq), _ptr(logp), _ptr(probs), mode,
                                        _ptr(prior_probs), _ptr(u_in), K, M, E, seed, offset, _stream())
    if rc in (ERR_UNSUPPORTED, ERR_ALIGN):
        return None
    check(rc, "zs_bernoulli_latent_fwd")
    _count()
    return z, logq, logp


def bernoulli_latent_bwd(dlogq, z, probs, mode, K, M, E):
    dt = z.dtype
    for t, n in ((dlogq, "dlogq"), (z, "z"), (probs, "probs")):
        _chk_tensor(t, n, dt)
    dprobs = torch.empty_like(probs)
    check(load().zs_bernoulli_latent_bwd(dtype_code(dt), _ptr(dprobs), _ptr(dlogq), _ptr(z), _ptr(probs), mode, K, M, E,
                                         _stream()), "zs_bernoulli_latent_bwd")
    _count()
    return dprobs


# ----------------------------------------------------------------------------- Bernoulli
def bernoulli_sample(probs, pm, K, N, u_in=None, seed=0, offset=0):
    dt = probs.dtype
    _chk_tensor(probs, "probs", dt)
    _chk_tensor(u_in, "u_in", dt)
    out = torch.empty((K, N), dtype=dt, device=probs.device)
    check(load().zs_bernoulli_sample(dtype_code(dt), _ptr(out), _ptr(probs), pm, _ptr(u_in), K, N, seed, offset,
                                     _stream()), "zs_bernoulli_sample")
    _count()
    return out


def bernoulli_logpmf_fwd(x, xm, probs, pm, K, M, E, logits=False):
    """`logits=True`: `probs` holds logits and the sigmoid is applied in registers (zs_bernoulli_logits_logpmf_fwd)."""
    dt = probs.dtype
    _chk_tensor(x, "x", dt)
    _chk_tensor(probs, "probs", dt)
    out = torch.empty((K, M), dtype=dt, device=probs.device)
    fn = load().zs_bernoulli_logits_logpmf_fwd if logits else load().zs_bernoulli_logpmf_fwd
    check(fn(dtype_code(dt), _ptr(out), _ptr(x), xm, _ptr(probs), pm, K, M, E, _stream()),
          "zs_bernoulli_logits_logpmf_fwd" if logits else "zs_bernoulli_logpmf_fwd")
    _count()
    return out


def bernoulli_logpmf_bwd(g, x, xm, probs, pm, K, M, E, need_x, need_probs, logits=False):
    dt = probs.dtype
    _chk_tensor(g, "g", dt)
    dx = torch.empty_like(x) if need_x else None
    dprobs = torch.empty_like(probs) if need_probs else None
    fn = load().zs_bernoulli_logits_logpmf_bwd if logits else load().zs_bernoulli_logpmf_bwd
    check(fn(dtype_code(dt), _ptr(dx), _ptr(dprobs), _ptr(g), _ptr(x), xm, _ptr(probs), pm, K, M, E, _stream()),
          "zs_bernoulli_logits_logpmf_bwd" if logits else "zs_bernoulli_logpmf_bwd")
    _count()
    return dx, dprobs


# ----------------------------------------------------------------------------- Logistic / Laplace
FAM_LOGISTIC, FAM_LAPLACE = 1, 2


def locscale_sample(family, loc, loc_mode, scale, scale_mode, K, N, u_in=None, seed=0, offset=0):
    dt = loc.dtype
    for t, n in ((loc, "loc"), (scale, "scale"), (u_in, "u_in")):
        _chk_tensor(t, n, dt)
    z = torch.empty((K, N), dtype=dt, device=loc.device)
    check(load().zs_locscale_sample(dtype_code(dt), family, _ptr(z), _ptr(loc), loc_mode, _ptr(scale), scale_mode,
                                    _ptr(u_in), K, N, seed, offset, _stream()), "zs_locscale_sample")
    _count()
    return z


def locscale_sample_bwd(family, dz, loc_like, loc_mode, scale_like, scale_mode, K, N, u=None, seed=0, offset=0,
                        need_loc=True, need_scale=True):
    dt = dz.dtype
    _chk_tensor(dz, "dz", dt)
    _chk_tensor(u, "u", dt)
    dloc = torch.empty_like(loc_like) if need_loc else None
    dscale = torch.empty_like(scale_like) if need_scale else None
    check(load().zs_locscale_sample_bwd(dtype_code(dt), family, _ptr(dloc), loc_mode, _ptr(dscale), scale_mode, _ptr(dz),
                                        _ptr(u), K, N, seed, offset, _stream()), "zs_locscale_sample_bwd")
    _count()
    return dloc, dscale


def locscale_logprob_fwd(family, x, xm, loc, lm, scale, sm, K, M, E):
    dt = x.dtype
    for t, n in ((x, "x"), (loc, "loc"), (scale, "scale")):
        _chk_tensor(t, n, dt)
    out = torch.empty((K, M), dtype=dt, device=x.device)
    check(load().zs_locscale_logprob_fwd(dtype_code(dt), family, _ptr(out), _ptr(x), xm, _ptr(loc), lm, _ptr(scale), sm,
                                         K, M, E, _stream()), "zs_locscale_logprob_fwd")
    _count()
    return out


def locscale_logprob_bwd(family, g, x, xm, loc, lm, scale, sm, K, M, E, need_x, need_loc, need_scale):
    dt = x.dtype
    _chk_tensor(g, "g", dt)
    dx = torch.empty_like(x) if need_x else None
    dloc = torch.empty_like(loc) if need_loc else None
    dscale = torch.empty_like(scale) if need_scale else None
    check(load().zs_locscale_logprob_bwd(dtype_code(dt), family, _ptr(dx), _ptr(dloc), _ptr(dscale), _ptr(g), _ptr(x), xm,
                                         _ptr(loc), lm, _ptr(scale), sm, K, M, E, _stream()), "zs_locscale_logprob_bwd")
    _count()
    return dx, dloc, dscale


# ----------------------------------------------------------------------------- Categorical
def categorical_sample(logits, lm, K, M, C, u_in=None, seed=0, offset=0):
    dt = logits.dtype
    _chk_tensor(logits, "logits", dt)
    _chk_tensor(u_in, "u_in", dt)
    out = torch.empty((K, M), dtype=dt, device=logits.device)
    check(load().zs_categorical_sample(dtype_code(dt), _ptr(out), _ptr(logits), lm, _ptr(u_in), K, M, C, seed, offset,
                                       _stream()), "zs_categorical_sample")
    _count()
    return out


def categorical_logpmf_fwd(x, xm, logits, lm, K, M, C):
    dt = logits.dtype
    _chk_tensor(x, "x", dt)
    _chk_tensor(logits, "logits", dt)
    out = torch.empty((K, M), dtype=dt, device=logits.device)
    check(load().zs_categorical_logpmf_fwd(dtype_code(dt), _ptr(out), _ptr(x), xm, _ptr(logits), lm, K, M, C, _stream()),
          "zs_categorical_logpmf_fwd")
    _count()
    return out


def categorical_logpmf_bwd(g, x, xm, logits, lm, K, M, C):
    dt = logits.dtype
    _chk_tensor(g, "g", dt)
    d = torch.empty_like(logits)
    check(load().zs_categorical_logpmf_bwd(dtype_code(dt), _ptr(d), _ptr(g), _ptr(x), xm, _ptr(logits), lm, K, M, C,
                                           _stream()), "zs_categorical_logpmf_bwd")
    _count()
    return d


# ----------------------------------------------------------------------------- objectives
def iw_objective(estimator, logp, logq, grad_scale, extra=None, need_grads=True):
    """logp/logq [K,B] -> (cost[B], dlogp[K,B], dlogq[K,B])."""
    dt = logp.dtype
    _chk_tensor(logp, "logp", dt)
    _chk_tensor(logq, "logq", dt)
    _chk_tensor(extra, "extra", dt)
    K, B = logp.shape
    cost = torch.empty(B, dtype=dt, device=logp.device)
    dlp = torch.empty_like(logp) if need_grads else None
    dlq = torch.empty_like(logp) if need_grads else None
    check(load().zs_iw_objective(dtype_code(dt), estimator, _ptr(cost), _ptr(dlp), _ptr(dlq), _ptr(logp), _ptr(logq),
                                 _ptr(extra), K, B, float(grad_scale), _stream()), "zs_iw_objective")
    _count()
    return cost, dlp, dlq


def log_mean_exp(x):
    dt = x.dtype
    _chk_tensor(x, "x", dt)
    K, B = x.shape
    out = torch.empty(B, dtype=dt, device=x.device)
    check(load().zs_log_mean_exp(dtype_code(dt), _ptr(out), _ptr(x), K, B, _stream()), "zs_log_mean_exp")
    _count()
    return out


def log_mean_exp_bwd(g, x):
    dt = x.dtype
    _chk_tensor(g, "g", dt)
    K, B = x.shape
    dx = torch.empty_like(x)
    check(load().zs_log_mean_exp_bwd(dtype_code(dt), _ptr(dx), _ptr(g), _ptr(x), K, B, _stream()),
          "zs_log_mean_exp_bwd")
    _count()
    return dx


def fused_supported(K, X, dtype):
    """Shapes zs_iw_bernoulli_fused takes (the ring streams rows, so K*X need not fit anywhere)."""
    return dtype == torch.float32 and X % 4 == 0 and 1 <= K <= 4096 and X <= (1 << 20)


def fused_logits_supported(K, X, dtype):
    """Shapes zs_iw_bernoulli_fused_logits takes: a fixed-geometry box kernel must be instantiated for the row
    length and a column plus one box must fit in shared memory (mirrors launch_fused_box in zs_fused_iw.cu)."""
    geo = {784: (112, 7), 128: (128, 1), 256: (256, 1), 512: (256, 2), 1024: (256, 4)}
    if dtype != torch.float32 or X not in geo or not (1 <= K <= 50):
        return False
    inner, nbox = geo[X]
    kpad = (K + 3) & ~3
    slot = (K * inner * 4 + 127) // 128 * 128
    fixed = 6 * X * 4 + 16 + 6 * kpad * 8 + 12 * kpad * 4 + 64
    return (nbox + 1) * (slot + 16) + fixed <= 227 * 1024


def iw_bernoulli_fused(estimator, probs, x, logp_other, logq, grad_scale, need_dprobs=True, want_logpx=False,
                       out=None, logits=False, accumulate_cost=False):
    """probs [K,B,X], x [B,X], logp_other/logq [K,B] or None.  `logits=True`: `probs` holds logits and "dprobs" is
    the gradient w.r.t. them (zs_iw_bernoulli_fused_logits).  `accumulate_cost=True` (needs out["cost"]): the
    per-column objectives are ADDED to out["cost"] (zs_iw_bernoulli_fused_accumulate).
    Returns dict(cost[B], dprobs, dlogp, dlogq, logpx) or None when the shape is not supported."""
    for t, n in ((probs, "probs"), (x, "x"), (logp_other, "logp_other"), (logq, "logq")):
        _chk_tensor(t, n, torch.float32)
    K, B, X = probs.shape
    dev = probs.device
    o = out if out is not None else {}
    cost = o.get("cost") if "cost" in o else torch.empty(B, dtype=torch.float32, device=dev)
    dprobs = (o.get("dprobs") if "dprobs" in o else torch.empty_like(probs)) if need_dprobs else None
    dlp = o.get("dlogp") if "dlogp" in o else torch.empty((K, B), dtype=torch.float32, device=dev)
    dlq = o.get("dlogq") if "dlogq" in o else torch.empty((K, B), dtype=torch.float32, device=dev)
    lpx = (o.get("logpx") if "logpx" in o else torch.empty((K, B), dtype=torch.float32, device=dev)) if want_logpx else None
    fn = load().zs_iw_bernoulli_fused_logits if logits else (
        load().zs_iw_bernoulli_fused_accumulate if accumulate_cost else load().zs_iw_bernoulli_fused)
    rc = fn(estimator, _ptr(cost), _ptr(dprobs), _ptr(dlp), _ptr(dlq), _ptr(lpx), _ptr(probs), _ptr(x), _ptr(logp_other),
            _ptr(logq), K, B, X, float(grad_scale), _stream())
    if rc in (ERR_UNSUPPORTED, ERR_ALIGN):
        return None
    check(rc, "zs_iw_bernoulli_fused")
    _count()
    return dict(cost=cost, dprobs=dprobs, dlogp=dlp, dlogq=dlq, logpx=lpx)


def reinforce_step(logp, logq, moving_mean, local_step, decay, need_grads=True):
    """ELBO.reinforce with the moving-mean baseline in one launch (zs_reinforce_step).  logp / logq: same-shaped
    contiguous CUDA tensors; moving_mean [1] float32 and local_step [1] int32 CUDA buffers, updated in place.
    Returns (cost [1], dlogp, dlogq) with the gradients of the mean surrogate."""
    dt = logq.dtype
    _chk_tensor(logp, "logp", dt)
    _chk_tensor(logq, "logq", dt)
    _chk_tensor(moving_mean, "moving_mean", torch.float32)
    _chk_tensor(local_step, "local_step", torch.int32)
    n = logq.numel()
    cost = torch.empty(1, dtype=dt, device=logq.device)
    dlp = torch.empty_like(logq) if need_grads else None
    dlq = torch.empty_like(logq) if need_grads else None
    check(load().zs_reinforce_step(dtype_code(dt), _ptr(cost), _ptr(dlp), _ptr(dlq), _ptr(moving_mean), _ptr(local_step),
                                   _ptr(logp), _ptr(logq), n, float(decay), 1.0 / n, _stream()), "zs_reinforce_step")
    _count()
    return cost, dlp, dlq


# ----------------------------------------------------------------------------- SG-MCMC
def scale_inplace(buf, scale_dev):
    """buf *= scale_dev[0] on the device; the launch exits immediately when the scalar is 1."""
    dt = buf.dtype
    _chk_tensor(buf, "buf", dt)
    _chk_tensor(scale_dev, "scale", dt)
    check(load().zs_scale_inplace(dtype_code(dt), _ptr(buf), buf.numel(), _ptr(scale_dev), _stream()),
          "zs_scale_inplace")
    _count()


def sgld_step(w, g, lr, noise=None, seed=0, offset=0, out=None):
    """Returns the updated chain state (a new tensor unless `out` is given; out may be w)."""
    dt = w.dtype
    for t, n in ((w, "w"), (g, "g"), (noise, "noise"), (out, "out")):
        _chk_tensor(t, n, dt)
    out = torch.empty_like(w) if out is None else out
    check(load().zs_sgld_step(dtype_code(dt), _ptr(out), _ptr(w), _ptr(g), _ptr(noise), w.numel(), float(lr), seed,
                              offset, _stream()), "zs_sgld_step")
    _count()
    return out


def psgld_step(w, aux, g, lr, decay, epsilon, noise_unit=None, seed=0, offset=0, out=None):
    dt = w.dtype
    for t, n in ((w, "w"), (aux, "aux"), (g, "g"), (noise_unit, "noise_unit"), (out, "out")):
        _chk_tensor(t, n, dt)
    out = torch.empty_like(w) if out is None else out
    check(load().zs_psgld_step(dtype_code(dt), _ptr(out), _ptr(w), _ptr(aux), _ptr(g), _ptr(noise_unit), w.numel(),
                               float(lr), float(decay), float(epsilon), seed, offset, _stream()), "zs_psgld_step")
    _count()
    return out


def sghmc_pre(w, v, lr, resample, second_order, v_noise=None, seed=0, offset=0, out=None):
    """Velocity resample (in place on v) and, for second order, the half step; returns the new w
    (w itself when no half step is taken)."""
    dt = w.dtype
    for t, n in ((w, "w"), (v, "v"), (v_noise, "v_noise"), (out, "out")):
        _chk_tensor(t, n, dt)
    if not resample and not second_order:
        return w
    if second_order:
        out = torch.empty_like(w) if out is None else out
    else:
        out = w
    check(load().zs_sghmc_pre(dtype_code(dt), _ptr(out), _ptr(w), _ptr(v), _ptr(v_noise), w.numel(), float(lr),
                              int(bool(resample)), int(bool(second_order)), seed, offset, _stream()), "zs_sghmc_pre")
    _count()
    return out


def sghmc_post(w, v, g, lr, alpha, beta, second_order, noise=None, seed=0, offset=0, out=None):
    dt = w.dtype
    for t, n in ((w, "w"), (v, "v"), (g, "g"), (noise, "noise"), (out, "out")):
        _chk_tensor(t, n, dt)
    out = torch.empty_like(w) if out is None else out
    check(load().zs_sghmc_post(dtype_code(dt), _ptr(out), _ptr(w), _ptr(v), _ptr(g), _ptr(noise), w.numel(), float(lr),
                               float(alpha), float(beta), int(bool(second_order)), seed, offset, _stream()),
          "zs_sghmc_post")
    _count()
    return out


# ----------------------------------------------------------------------------- host-buffer step (e2e)
def iw_step_host_workspace(K, B, X):
    return int(load().zs_iw_step_host_workspace(K, B, X))


def iw_step_host(estimator, cost_h, dprobs_h, dlogp_h, dlogq_h, probs_h, x_h, other_h, logq_h, K, B, X, grad_scale, ws):
    """All *_h are pinned CPU tensors (or None); ws is a device uint8 workspace.  Synchronises."""
    hp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    check(load().zs_iw_step_host(estimator, hp(cost_h), hp(dprobs_h), hp(dlogp_h), hp(dlogq_h), hp(probs_h), hp(x_h),
                                 hp(other_h), hp(logq_h), K, B, X, float(grad_scale), _ptr(ws), ws.numel(), _stream()),
          "zs_iw_step_host")
    _count()


def iw_step_host_begin(estimator, cost_h, dprobs_h, dlogp, dlogq, probs_h, x_h, other, logq, K, B, X, grad_scale, ws,
                       scalars_on_device):
    """Enqueue the host-buffer step and return.  cost_h / dprobs_h / probs_h / x_h are pinned CPU tensors;
    other / logq / dlogp / dlogq are CUDA tensors [K,B] when `scalars_on_device` else pinned CPU tensors.
    Follow with iw_step_host_wait(0) (cost and every kernel done) and iw_step_host_wait(1) (dprobs landed)."""
    hp = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    check(load().zs_iw_step_host_begin(estimator, hp(cost_h), hp(dprobs_h), hp(dlogp), hp(dlogq), hp(probs_h), hp(x_h),
                                       hp(other), hp(logq), K, B, X, float(grad_scale), _ptr(ws), ws.numel(),
                                       1 if scalars_on_device else 0, _stream()),
          "zs_iw_step_host_begin")
    _count()


def iw_step_host_wait(what):
    check(load().zs_iw_step_host_wait(int(what)), "zs_iw_step_host_wait")
