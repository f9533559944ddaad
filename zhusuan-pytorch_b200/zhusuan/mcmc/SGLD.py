"""SGLD and preconditioned SGLD (drop-in for zhusuan/mcmc/SGLD.py).

ALL chain-state tensors of the net are updated by ONE launch (zs_sgmcmc_multi_step) that reads w and the
gradient, draws the Gaussian term from Philox in registers and writes the new state (12 B per element for SGLD).
The reference loops over the latents in Python and, per tensor, draws the noise on the CPU, copies it to the device
and runs 3 elementwise kernels (:49-54).
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
        homes = [w.device for w in self._var_list]
        ws = [self._on_device(w) for w in self._var_list]
        gs = [self._on_device(g, w.dtype) for g, w in zip(grad, ws)]
        noises = [self._noise(w) for w in ws]  # injected Gaussian terms (parity tests), in variable order
        new = self._multi_step(_be.ALG_SGLD, ws, gs, None, noises, lr=float(self.lr))
        self._var_list = [self._leaf(w, h) for w, h in zip(new, homes)]


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
        homes = [w.device for w in self._var_list]
        ws = [self._on_device(w) for w in self._var_list]
        if not self.aux:
            self.aux = [torch.zeros_like(w) for w in ws]  # running second moments, kept on the compute device
        grad = self._gradients(bn, observed)
        gs = [self._on_device(g, w.dtype) for g, w in zip(grad, ws)]
        units = [self._noise(w) for w in ws]
        new = self._multi_step(_be.ALG_PSGLD, ws, gs, self.aux, units, lr=float(self.lr), a=self.decay, b=self.epsilon)
        self._var_list = [self._leaf(w, h) for w, h in zip(new, homes)]
