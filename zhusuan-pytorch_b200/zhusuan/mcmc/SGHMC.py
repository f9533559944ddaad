"""SGHMC (Chen et al. 2014), first order and second order (symmetric splitting) — drop-in for
zhusuan/mcmc/SGHMC.py.

Per chain-state tensor an update is at most two kernels around the gradient evaluation:
  pre : optional velocity resample v ~ N(0, lr) and, for second order, the half step w += v/2
  post: v = (1-alpha) v + lr g + n ; w += v                      (first order,  reference :46-50)
        v = d (d v + lr g + n), d = exp(-alpha/2) ; w += v/2      (second order, reference :51-56)
with n ~ N(0, 2 (alpha - beta) lr) drawn in registers.  Velocities live on the compute device (the
reference keeps them, and every noise draw, on the CPU: :27,33,34).
Deviation from the reference, on purpose: it reuses the LAST variable's Gaussian term for every
variable (:34 vs :47,53), which raises for differently shaped latents; here every variable gets its
own draw.  With a single latent (the only case the reference can run) the two coincide.
"""
import math

import torch

from zhusuan.mcmc.SGMCMC import SGMCMC
from zhusuan import _backend as _be
from zhusuan import _ops, _rng

__all__ = ["SGHMC"]


class SGHMC(SGMCMC):
    def __init__(self, learning_rate, friction=0.25, variance_estimate=0., n_iter_resample_v=20, second_order=True):
        super(SGHMC, self).__init__()
        self.lr = learning_rate
        self.alpha = friction
        self.beta = variance_estimate
        if n_iter_resample_v is None:
            n_iter_resample_v = 0
        self.n_iter_resample_v = n_iter_resample_v
        self.second_order = second_order
        self.vs = None  # velocities, one per latent, on the compute device

    def _draw_velocity(self, like):
        v = self._noise(like)
        if v is None:
            seed, offset = _rng.next_philox(like.device)
            v = _be.philox_normal(like.numel(), like.dtype, 0.0, math.sqrt(self.lr), seed, offset,
                                  like.device).reshape(like.shape)
        return v

    def _update(self, bn, observed):
        states = [_ops.to_compute(q.detach()).contiguous() for q in self._var_list]
        homes = [q.device for q in self._var_list]
        if not self.vs:
            self.vs = [self._draw_velocity(w) for w in states]
        resample = self.n_iter_resample_v != 0 and self.t % self.n_iter_resample_v == 0
        gaussian = []
        for i, w in enumerate(states):
            # draw order per variable follows the reference: velocity resample first, then the term
            v_noise = self._noise(w) if resample else None
            seed, offset = (0, 0) if (v_noise is not None or not resample) else _rng.next_philox(w.device)
            half = _be.sghmc_pre(w, self.vs[i], self.lr, resample, self.second_order, v_noise=v_noise, seed=seed,
                                 offset=offset)
            gaussian.append(self._noise(w))
            if self.second_order:
                states[i] = half
                self._var_list[i] = self._leaf(half, homes[i])
        grad = self._gradients(bn, observed)
        for i, g in enumerate(grad):
            w = states[i]
            gd = _ops.to_compute(g.detach()).to(w.dtype).contiguous()
            n = gaussian[i]
            seed, offset = (0, 0) if n is not None else _rng.next_philox(w.device)
            new = _be.sghmc_post(w, self.vs[i], gd, self.lr, self.alpha, self.beta, self.second_order, noise=n,
                                 seed=seed, offset=offset)
            self._var_list[i] = self._leaf(new, homes[i])
