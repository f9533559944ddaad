"""zhusuan.utils — drop-in for zhusuan/utils.py of the reference."""
import torch

from . import _ops

__all__ = ['log_mean_exp', 'particle_linear']


def log_mean_exp(x, dim=None, keepdims=False):
    """Numerically stable log(mean(exp(x))) along `dim` (reference: zhusuan/utils.py:6-21),
    one kernel launch instead of six.  `dim=None` reduces over all elements."""
    x = torch.as_tensor(x)
    if dim is None:
        out = _ops.log_mean_exp(x.reshape(-1), 0, False)
        return out.reshape([1] * x.dim()) if keepdims else out
    if isinstance(dim, (list, tuple)):
        if len(dim) != 1:
            raise TypeError("log_mean_exp: a single reduction axis is supported")
        dim = dim[0]
    return _ops.log_mean_exp(x, int(dim), bool(keepdims))


def particle_linear(w, h, scale=True):
    """The per-particle layer of the reference's Bayesian-NN examples (examples/bayesian_neural_nets/bnn_vi.py:39-45,
    bnn_sgmcmc.py:47-53) without its [K, B, n_out, n_in+1] copy of the weights:

        w [K, n_out, n_in + 1]  weight particles / chains (the last input column is the bias)
        h [K, B, n_in] or [B, n_in]  activations (shared by all particles when 2-D)
        -> [K, B, n_out] = (h ++ 1) @ w_k^T / sqrt(n_in + 1)      (`scale=False` drops the division)

    The reference `repeat`s w over the batch and runs a [K,B,n_out,n_in+1] x [K,B,n_in+1,1] matmul, which makes
    B copies of every weight; this is one cuBLAS batched GEMM (torch.baddbmm) over the K particles, with the bias
    column folded into the addend.  SURVEY 8(f)-3: the GEMM-shaped caller next to the hot path; it is a library
    GEMM and not part of the HBM-bound claim."""
    w = torch.as_tensor(w)
    K, n_out, n_in1 = w.shape
    h = torch.as_tensor(h, dtype=w.dtype, device=w.device)
    if h.dim() == 2:
        h = h.unsqueeze(0).expand(K, h.shape[0], h.shape[1])
    if h.shape[0] != K or h.shape[-1] != n_in1 - 1:
        raise RuntimeError("particle_linear: w %s does not match h %s" % (tuple(w.shape), tuple(h.shape)))
    bias = w[:, :, -1].unsqueeze(1)                      # [K, 1, n_out]
    out = torch.baddbmm(bias, h, w[:, :, :-1].transpose(1, 2))   # [K, B, n_out]
    if scale:
        out = out / torch.sqrt(torch.as_tensor(float(n_in1), dtype=torch.float32, device=w.device)).to(w.dtype)
    return out
