"""Base class of the stochastic-gradient MCMC samplers (drop-in for zhusuan/mcmc/SGMCMC.py).

State machine of the reference (:38-59): `sample(bn, observed, resample=True)` runs the net once,
snapshots every unobserved node's `.tensor` (a fresh draw) as a detached leaf and performs NO
update; later calls run `step` updates.  The returned dict maps latent names to leaf tensors
(requires_grad=True) that are new objects after every update.
"""
import torch
import torch.nn as nn

from zhusuan import _ops, _rng

__all__ = ["SGMCMC"]


class SGMCMC(nn.Module):
    def __init__(self):
        super().__init__()
        self.t = 0

    def _update(self, bn, observed):
        raise NotImplementedError()

    # -- helpers shared by the concrete samplers ----------------------------------------------------
    def _gradients(self, bn, observed):
        """Gradient of the net's log joint wrt the current chain states (reference SGLD.py:43-48)."""
        observed_ = {**dict(zip(self._latent_k, self._var_list)), **observed}
        bn.forward(observed_)
        log_joint_ = bn.log_joint()
        return torch.autograd.grad(log_joint_, self._var_list)

    @staticmethod
    def _noise(like):
        """Injected Gaussian term for the next draw (parity tests), on the compute device, or None."""
        n = _rng.take_injected("normal")
        if n is None:
            return None
        return _ops.to_compute(torch.as_tensor(n)).to(like.dtype).reshape(like.shape).contiguous()

    @staticmethod
    def _leaf(t, home):
        t = _ops.back_home(t, home).detach()
        t.requires_grad = True
        return t

    @staticmethod
    def _on_device(t, dtype=None):
        """The chain-state / gradient tensor as the kernels read it: as is when it already is a contiguous CUDA
        tensor (the usual case: no per-tensor detach / copy), otherwise moved / made contiguous."""
        if isinstance(t, _ops.LazyDraw):
            t = t._zs_materialize()
        if t.is_cuda and t.is_contiguous() and (dtype is None or t.dtype == dtype):
            return t.detach() if t.requires_grad else t
        t = _ops.to_compute(t.detach())
        return (t if dtype is None else t.to(dtype)).contiguous()

    def _multi_step(self, algorithm, ws, gs=None, states=None, noises=None, **coef):
        """One launch over every chain-state tensor (zs_sgmcmc_multi_step); tensors of different dtypes (rare) go in
        one launch per dtype.  Returns the updated states in the order of `ws`."""
        from zhusuan import _backend as _be
        n = len(ws)
        out = [None] * n
        groups = {}
        for i, w in enumerate(ws):
            groups.setdefault((w.dtype, w.device), []).append(i)
        for (_, dev), idx in groups.items():
            pick = lambda lst: None if lst is None else [lst[i] for i in idx]
            nz = pick(noises)
            injected = nz is not None and all(v is not None for v in nz)
            kw = dict(seed=0, offset=0) if injected else _rng.draw_args(dev)
            res = _be.sgmcmc_multi_step(algorithm, pick(ws), pick(gs), pick(states), nz, None, **coef, **kw)
            for i, r in zip(idx, res):
                out[i] = r
        return out

    def forward(self, bn, observed, resample=False, step=1):
        if resample:
            self.t = 0
            with _ops.lazy_first_draws():  # the states are the `.tensor` reads below, not stochastic_node's returns
                bn.forward(observed)
            self.t += 1
            self._latent = {k: v.tensor for k, v in bn.nodes.items() if k not in observed.keys()}
            self._latent_k = self._latent.keys()
            self._var_list = [self._latent[k] for k in self._latent_k]
            sample_ = dict(zip(self._latent_k, self._var_list))
            for i in range(len(self._var_list)):
                self._var_list[i] = self._var_list[i].detach()
                self._var_list[i].requires_grad = True
            return sample_

        for _ in range(step):
            self._update(bn, observed)
            self.t += 1
        return dict(zip(self._latent_k, self._var_list))

    def initialize(self):
        self.t = 0

    def sample(self, bn, observed, resample=False, step=1):
        """Run one sampler call; see the class docstring for the resample protocol."""
        return self.forward(bn, observed, resample, step)
