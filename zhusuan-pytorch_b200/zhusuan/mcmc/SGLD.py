"""SGLD and preconditioned SGLD (drop-in for zhusuan/mcmc/SGLD.py).

Each chain-state tensor is updated by ONE kernel that reads w and the gradient, draws the Gaussian
term from Philox in registers and writes the new state (12 B per element for SGLD).  The reference
draws the noise on the CPU, copies it to the device and runs 3 elementwise kernels (:50-52).
"""
import math

import torch

from zhusuan.mcmc.SGMCMC import SGMCMC
from zhusuan import _backend as _be
from zhusuan import _ops, _rng


class SGLD(SGMCMC):
    """Stochastic Gradient Langevin Dynamics (Welling & Teh 2011), eq. (3):
    w <- w + lr/2 * grad log p(w, data) + N(0, lr)."""

    def __init__(self, learning_rate):
        super().__init__()
        self.lr = torch.as_tensor(learning_rate)  # a 0-d tensor, as in the reference (:20)
        self.lr_min = torch.as_tensor(1e-4)
        self._device = torch.device('cpu')

    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return self._device

    def to(self, device):
        self._device = device
        return super().to(device)

    def _update(self, bn, observed):
        grad = self._gradients(bn, observed)
        lr = float(self.lr)
        for i, g in enumerate(grad):
            w = self._var_list[i]
            home = w.device
            wd = _ops.to_compute(w.detach()).contiguous()
            gd = _ops.to_compute(g.detach()).to(wd.dtype).contiguous()
            noise = self._noise(wd)
            seed, offset = (0, 0) if noise is not None else _rng.next_philox(wd.device)
            new = _be.sgld_step(wd, gd, lr, noise=noise, seed=seed, offset=offset)
            self._var_list[i] = self._leaf(new, home)


class PSGLD(SGLD):
    """SGLD with an RMSprop preconditioner (Li et al. 2016).  `aux` holds the running second moment
    per chain state; an injected noise tensor is interpreted as UNIT normals (the reference's
    torch.normal(0, std_tensor) equals std_tensor * xi, :79)."""

    def __init__(self, learning_rate, decay=0.9, epsilon=1e-3):
        super().__init__(learning_rate)
        self.aux = None
        self.decay = decay
        self.epsilon = epsilon

    def _update(self, bn, observed):
        if not self.aux:
            self.aux = [torch.zeros_like(_ops.to_compute(q.detach())) for q in self._var_list]
        grad = self._gradients(bn, observed)
        lr = float(self.lr)
        for i, g in enumerate(grad):
            w = self._var_list[i]
            home = w.device
            wd = _ops.to_compute(w.detach()).contiguous()
            gd = _ops.to_compute(g.detach()).to(wd.dtype).contiguous()
            unit = self._noise(wd)
            seed, offset = (0, 0) if unit is not None else _rng.next_philox(wd.device)
            new = _be.psgld_step(wd, self.aux[i], gd, lr, self.decay, self.epsilon, noise_unit=unit, seed=seed,
                                 offset=offset)
            self._var_list[i] = self._leaf(new, home)
