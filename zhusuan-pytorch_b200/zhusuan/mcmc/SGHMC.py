"""SGHMC (Chen et al. 2014), first order and second order (symmetric splitting) — drop-in for
zhusuan/mcmc/SGHMC.py.

An update is at most two launches around the gradient evaluation, each covering EVERY chain-state tensor
(zs_sgmcmc_multi_step):
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
            v = _be.philox_normal(like.numel(), like.dtype, 0.0, math.sqrt(self.lr), device=like.device,
                                  **_rng.draw_args(like.device)).reshape(like.shape)
        return v

    def _update(self, bn, observed):
        homes = [q.device for q in self._var_list]
        states = [self._on_device(q) for q in self._var_list]
        if not self.vs:
            self.vs = [self._draw_velocity(w) for w in states]  # velocities live on the compute device
        resample = self.n_iter_resample_v != 0 and self.t % self.n_iter_resample_v == 0
        # injected noise, per variable in the reference's draw order: resampled velocity first, then the Gaussian term
        v_noise, gaussian = [], []
        for w in states:
            v_noise.append(self._noise(w) if resample else None)
            gaussian.append(self._noise(w))
        if resample or self.second_order:
            states = self._multi_step(_be.ALG_SGHMC_PRE, states, None, self.vs, v_noise, lr=self.lr,
                                      resample=resample, second_order=self.second_order)
            if self.second_order:
                self._var_list = [self._leaf(w, h) for w, h in zip(states, homes)]
                states = [self._on_device(q) for q in self._var_list]
        grad = self._gradients(bn, observed)
        gs = [self._on_device(g, w.dtype) for g, w in zip(grad, states)]
        new = self._multi_step(_be.ALG_SGHMC_POST, states, gs, self.vs, gaussian, lr=self.lr, a=self.alpha,
                               b=self.beta, second_order=self.second_order)
        self._var_list = [self._leaf(w, h) for w, h in zip(new, homes)]
