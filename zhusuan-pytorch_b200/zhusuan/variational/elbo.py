"""Evidence lower bound objective — drop-in for zhusuan/variational/elbo.py of the reference.

`forward` keeps the reference's protocol (:81-132): run the variational net, feed its samples to the
generator as observations, sum the nodes' log-probabilities, apply the estimator.  Any object with
`.tensor` and `.log_prob()` is accepted as a node (the reference's tests plant such objects).
The `sgvb` estimator over [K, ...] log-probabilities runs as one kernel (ZS_EST_ELBO).
"""
import torch
import torch.nn as nn

from zhusuan import _ops
from zhusuan import _backend as _be

__all__ = ['ELBO', 'EvidenceLowerBoundObjective']


class ELBO(nn.Module):
    def __init__(self, generator, variational, estimator='sgvb', transform=None, transform_var=[], auxillary_var=[]):
        super(ELBO, self).__init__()
        self.generator = generator
        self.variational = variational
        if estimator not in ('sgvb', 'reinforce'):
            raise NotImplementedError()
        self.estimator = estimator
        if estimator == 'reinforce':
            # same buffers as the reference (:45-49), so state_dicts are interchangeable
            self.register_buffer('moving_mean', torch.zeros(size=[1], dtype=torch.float32))
            self.register_buffer('local_step', torch.zeros(size=[1], dtype=torch.int32))
        if transform:
            self.transform = transform
            self.transform_var = transform_var
            self.auxillary_var = auxillary_var
        else:
            self.transform = None

    def log_joint(self, nodes):
        """Sum of `node.log_prob()` over `nodes` (a dict), in insertion order."""
        total = None
        for name in nodes.keys():
            lp = nodes[name].log_prob()
            total = lp if total is None else total + lp
        return total

    def forward(self, observed, reduce_mean=True, **kwargs):
        # host-resident parameters cross PCIe once per step; large host-bound results are copied back lazily
        with _ops.upload_memo(), _ops.lazy_host_results():
            return self._forward(observed, reduce_mean, **kwargs)

    def _forward(self, observed, reduce_mean=True, **kwargs):
        with _ops.lazy_first_draws():  # draw #1 is materialised only if the variational net's forward uses it
            self.variational(observed)
        nodes_q = self.variational.nodes
        log_det = None
        latents = {}
        if self.transform is not None:
            # a flow (any callable returning (outputs, log_det)) rewrites the named latents (:90-119)
            flow_inputs = []
            for k in self.transform_var:
                assert k not in observed.keys()
                assert k in nodes_q.keys()
                flow_inputs.append(nodes_q[k].tensor)
            for k in self.auxillary_var:
                flow_inputs.append(self.variational.cache[k])
            output, log_det = self.transform(tuple(flow_inputs))
            assert len(output) == len(self.transform_var)
            for k in self.transform_var:
                latents[k] = output[k]
        for k, v in nodes_q.items():
            if k not in latents:
                latents[k] = v.tensor
        self.generator({**latents, **observed})
        nodes_p = self.generator.nodes
        logpxz = self.log_joint(nodes_p)
        logqz = self.log_joint(nodes_q)
        if self.estimator == "sgvb":
            return self.sgvb(logpxz, logqz, reduce_mean, log_det)
        return self.reinforce(logpxz, logqz, reduce_mean, **kwargs)

    def sgvb(self, logpxz, logqz, reduce_mean=True, log_det=None):
        """-(E[log p - log q]) with the pathwise ("reparameterisation") gradient (:134-161).
        The mean runs over ALL axes, particles included, as in the reference (:155-156)."""
        if torch.is_tensor(logqz) and logqz.dim() > 0 and reduce_mean:
            lp, lq = torch.broadcast_tensors(torch.as_tensor(logpxz, dtype=logqz.dtype, device=logqz.device), logqz)
            if (log_det is not None and torch.is_tensor(log_det) and log_det.numel() > 0
                    and lq.dtype in (torch.float32, torch.float64) and lq.numel() > 0):
                # the flow's term folded into the objective's reduction (zs_combine_sums): two launches in all
                return _ops.elbo_with_log_det(lp, lq, log_det)
            cost = _ops.iw_objective(lp.reshape(lp.shape[0], -1), lq.reshape(lq.shape[0], -1), 0, _be.ELBO, True)
        else:
            cost = -(logpxz - logqz)
        if log_det is not None:
            cost = cost - torch.mean(torch.sum(log_det)).squeeze()
        return cost

    def reinforce(self, logpxz, logqz, reduce_mean=True, baseline=None, variance_reduction=True, decay=0.8):
        """Score-function estimator with a moving-mean baseline (:163-238).  The stored moving mean is
        bias-corrected in place every call, exactly as the reference does (:221-224)."""
        reduce = torch.is_tensor(logqz) and logqz.dim() > 0 and reduce_mean
        if (reduce and variance_reduction and baseline is None and torch.is_tensor(logpxz)
                and logpxz.dtype == logqz.dtype and logqz.dtype in (torch.float32, torch.float64)
                and logqz.numel() > 0):
            # the form the examples use: one kernel for the signal, the moving-mean update (in place, on the
            # device), the surrogate and both gradients
            return _ops.reinforce(logpxz, logqz, self.moving_mean, self.local_step, decay)
        signal = (logpxz - logqz).detach()
        baseline_cost = None
        if variance_reduction:
            if baseline is not None:
                baseline_cost = 0.5 * torch.square(signal - baseline)
                if reduce:
                    baseline_cost = torch.mean(baseline_cost)
                signal = signal - baseline
            centre = torch.mean(signal) if reduce else signal
            mm = self.moving_mean
            with torch.no_grad():
                mm -= ((mm - centre.detach().to(mm.device, mm.dtype)) * (1.0 - decay)).reshape(mm.shape)
                self.local_step += 1
                bias = 1 - torch.pow(torch.ones(size=[1], dtype=torch.float32, device=mm.device) * decay,
                                     self.local_step)
                mm /= bias
            signal = signal - mm.detach().to(signal.device, signal.dtype)
        signal = signal.detach()
        cost = -(logpxz + signal * logqz)
        if baseline_cost is not None:
            loss = torch.mean(cost + baseline_cost) if reduce else cost + baseline_cost
            return loss, torch.mean(logpxz - logqz)
        if reduce:
            cost = torch.mean(cost)
        return cost


class EvidenceLowerBoundObjective(ELBO):
    """Alias of :class:`ELBO` (reference :241-253)."""

    def __init__(self, generator, variational, estimator='sgvb', transform=None, transform_var=[], auxillary_var=[]):
        super().__init__(generator, variational, estimator, transform, transform_var, auxillary_var)
