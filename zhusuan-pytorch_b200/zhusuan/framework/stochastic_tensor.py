"""StochasticTensor — a named stochastic node of a BayesianNet.

Drop-in for zhusuan/framework/stochastic_tensor.py of the reference: same constructor, same
properties, and the same (stateful) semantics:
  * `.tensor` returns the observation when the node's name is observed in its net (and records it as
    the distribution's `sample_cache`), otherwise it draws a NEW sample on every access (:114-127);
  * `.log_prob()` evaluates at `sample_cache` when called without an argument and then applies
    reduce_mean_dims -> reduce_sum_dims -> squeeze -> multiplier (:160-181).
B200 build: trailing `reduce_sum_dims` are folded into the distribution's log-density kernel (one
pass over the [K,B,X] tensor produces the [K,B] result) instead of materialising the elementwise
log-probabilities and reducing them afterwards.
"""
import weakref

import torch
from zhusuan._shapes import broadcast_shapes as _bshapes

from zhusuan.distributions.base import Distribution

__all__ = ['StochasticTensor']


class StochasticTensor(object):
    def __init__(self, bn, name, dist, observation=None, n_samples=None, **kwargs):
        # weak back-reference: bn._nodes[name] -> node -> bn would otherwise be a reference cycle that
        # keeps every step's samples (device memory) alive until the cyclic GC happens to run
        self._bn_ref = weakref.ref(bn) if bn is not None else None
        self._name = name
        self._dist = dist
        self._dtype = dist.dtype
        self._n_samples = n_samples
        self._observation = observation
        self._reduce_mean_dims = kwargs.get("reduce_mean_dims", None)
        self._reduce_sum_dims = kwargs.get("reduce_sum_dims", None)
        self._multiplier = kwargs.get("multiplier", None)

    @property
    def bn(self):
        return self._bn_ref() if self._bn_ref is not None else None

    @property
    def _bn(self):
        return self.bn

    @property
    def name(self):
        return self._name

    @property
    def dtype(self):
        return self._dtype

    @property
    def dist(self):
        return self._dist

    def is_observed(self):
        return self._observation is not None

    def _observed_value(self):
        obs = self._bn.observed
        return obs[self._name] if self._name in obs.keys() else None

    @property
    def tensor(self):
        value = self._observed_value()
        if value is not None:
            self._dist.sample_cache = value
            return value
        return self._draw()

    def _draw(self):
        """A new sample; when the node's reductions end in an event sum the kernels can take, log q at that sample
        comes out of the same launch and is kept for `log_prob()` (distribution.sample_for_node)."""
        dist = self._dist
        K = self._n_samples
        try:
            shape = (((int(K),) if K is not None and int(K) > 1 else ()) + tuple(dist.batch_shape))
            n_event, _, _ = self.reduction_plan(shape)
        except Exception:
            return dist.sample(n_samples=K)
        return dist.sample_for_node(K, n_event)

    def first_draw(self):
        """The value `stochastic_node` hands to the user's `forward` (reference framework/bn.py:158): the observation,
        or draw #1.  Inside `_ops.lazy_first_draws()` (the objectives and samplers, which read `.tensor` again and
        use THAT draw) an unobserved node returns a LazyDraw -- the sampling launch runs only if `forward` actually
        uses the value (zhusuan/_lazy.py).  Injected noise keeps draws eager so it is consumed in reference order."""
        from zhusuan import _ops, _rng
        value = self._observed_value()
        if value is not None or not _ops.lazy_active() or _rng.has_injected():
            return self.tensor
        dist = self._dist
        K = self._n_samples
        try:
            shape = (((int(K),) if K is not None and int(K) > 1 else ()) + tuple(dist.batch_shape))
            device = dist.sample_device()
        except Exception:
            return self.tensor
        lazy = _ops.LazyDraw(self._materialize_lazy, shape, dist.dtype, device)
        dist.sample_cache = lazy
        dist._logq_cache = None
        return lazy

    def _materialize_lazy(self, lazy):
        """Run the deferred draw #1.  A newer draw / observation that has become the distribution's current value in
        the meantime stays current (`log_prob(None)` must keep evaluating at it)."""
        dist = self._dist
        current = (dist.sample_cache, dist._logq_cache)
        z = self._draw()
        if current[0] is not lazy:
            dist.sample_cache, dist._logq_cache = current
        return z

    def sample(self, force=False):
        value = None if force else self._observed_value()
        if value is not None:
            self._dist.sample_cache = value
            return value
        return self._draw()

    @property
    def shape(self):
        return self.tensor.shape

    def get_shape(self):
        return self.shape

    # -- reduction plan -------------------------------------------------------------------------
    def reduction_plan(self, value_shape):
        """Split the node's reductions into (n_event, mean_dims, sum_dims): `n_event` trailing axes
        are summed inside the log-density kernel, the rest is applied to its (small) output.
        Axes index the tensor left after the distribution's own group_ndims sum, as in the reference."""
        g = self._dist.group_ndims
        full = _bshapes(tuple(value_shape), tuple(self._dist.batch_shape))
        nd = len(full) - g
        mean_dims = sorted(set(d % nd for d in (self._reduce_mean_dims or []))) if nd > 0 else []
        sum_dims = sorted(set(d % nd for d in (self._reduce_sum_dims or []))) if nd > 0 else []
        fold = 0
        if sum_dims and sum_dims == list(range(nd - len(sum_dims), nd)) and not (set(sum_dims) & set(mean_dims)):
            fold = len(sum_dims)
            sum_dims = []
        return g + fold, mean_dims, sum_dims

    def log_prob(self, sample=None):
        dist = self._dist
        given = dist._given(sample)
        n_event, mean_dims, sum_dims = self.reduction_plan(given.shape)
        lp = dist.cached_log_prob(given, n_event) if sample is None else None
        if lp is None:
            lp = dist._log_prob_event(given, n_event)
        else:
            from zhusuan import _ops
            lp = _ops.back_home(lp, given.device)
        if mean_dims:
            lp = torch.mean(lp, mean_dims, keepdim=True)
        if sum_dims:
            lp = torch.sum(lp, sum_dims, keepdim=True)
        for d in sorted(set(mean_dims + sum_dims), reverse=True):
            lp = torch.squeeze(lp, d)
        if self._multiplier:
            lp = lp * self._multiplier
        return lp
