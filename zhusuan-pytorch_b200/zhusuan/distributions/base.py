"""Distribution base class — same constructor contract and public surface as the reference's
zhusuan/distributions/base.py (sample :132-153, log_prob :161-178, properties :93-123)."""
import torch

__all__ = ['Distribution']

# The reference's constructors default to device=torch.device('cpu') and move every parameter there
# (normal.py:50-58).  This object is that default; when a constructor receives exactly this object
# (i.e. the caller did not pass `device=`) parameters that already live on a CUDA device stay there,
# instead of being copied to the host and shuttled back for every kernel.
DEFAULT_DEVICE = torch.device('cpu')


def resolve_device(device, *params):
    """The device the parameters are kept on: an explicit `device=` wins, otherwise the device of the
    first tensor parameter, otherwise the CPU."""
    if device is not DEFAULT_DEVICE:
        return device
    for p in params:
        if torch.is_tensor(p):
            return p.device
    return device


class Distribution(object):
    """Base of the stochastic-node distributions.

    Samples have shape ``([n_samples] +) batch_shape``; ``log_prob(given)`` returns
    ``(... +) batch_shape[:-group_ndims]``.  Subclasses implement ``_sample``, ``_log_prob``,
    ``_prob``, ``_batch_shape`` and keep ``sample_cache`` up to date (the value ``log_prob(None)``
    is evaluated at, stochastic_tensor.py:123,137 of the reference).

    B200 build: ``_log_prob_event(given, n_event)`` lets a subclass fuse the event-axis sum (the
    ``group_ndims`` axes plus any trailing ``reduce_sum_dims`` of the owning StochasticTensor) into
    its log-density kernel.
    """

    def __init__(self, dtype, is_continuous, is_reparameterized, use_path_derivative=False, group_ndims=0,
                 device=DEFAULT_DEVICE, **kwargs):
        # unknown kwargs are accepted and ignored, as in the reference (base.py:77): user models pass
        # reduce_mean_dims / multiplier / check_numerics / n_samples through the distribution ctor
        self._dtype = dtype
        self._is_continuous = is_continuous
        self._is_reparameterized = is_reparameterized
        self._use_path_derivative = use_path_derivative
        self._device = device
        self.sample_cache = None
        self._logq_cache = None  # (sample, n_event, log q) of the last fused draw, see sample_for_node
        if isinstance(group_ndims, int):
            if group_ndims < 0:
                raise ValueError("group_ndims must be non-negative.")
            self._group_ndims = group_ndims
        else:
            self._group_ndims = int(group_ndims)

    @property
    def dtype(self):
        return self._dtype

    @property
    def device(self):
        return self._device

    @property
    def is_continuous(self):
        return self._is_continuous

    @property
    def is_reparameterized(self):
        return self._is_reparameterized

    @property
    def group_ndims(self):
        return self._group_ndims

    @property
    def batch_shape(self):
        return self._batch_shape()

    def _batch_shape(self):
        raise NotImplementedError()

    def sample(self, n_samples=None, **kwargs):
        """``n_samples=None`` draws one sample without a leading axis; an int n prepends ``[n]``
        (n == 1 also has no leading axis, as in the reference: normal.py:96-99)."""
        if n_samples is None:
            return self._sample(n_samples=1, **kwargs)
        if isinstance(n_samples, int):
            return self._sample(n_samples, **kwargs)
        return self._sample(int(n_samples), **kwargs)

    def _sample(self, n_samples, **kwargs):
        raise NotImplementedError()

    # -- B200 build: a node that knows its event reduction can ask for the sample and its log q together
    def _sample_fused(self, n_samples, n_event):
        """Return (sample, log q(sample) summed over the last n_event axes) from one launch, or None when the
        distribution / shape has no fused form."""
        return None

    def sample_for_node(self, n_samples, n_event):
        """`sample()` for a StochasticTensor that will evaluate `log_prob()` at this sample with `n_event`
        trailing axes summed: the log-density is produced by the sampling launch and kept for that call."""
        K = 1 if n_samples is None else int(n_samples)
        from zhusuan import distributions as _d
        fused = self._sample_fused(K, n_event) if _d.FUSED_LATENT else None
        if fused is None:
            self._logq_cache = None
            return self.sample(n_samples)
        z, logq = fused
        self.sample_cache = z
        self._logq_cache = (z, n_event, logq)
        return z

    def sample_device(self):
        """Device a sample of this distribution is returned on.  Distributions that support a deferred first draw
        (StochasticTensor.first_draw) override this; the default opts out."""
        raise NotImplementedError()

    def cached_log_prob(self, given, n_event):
        c = self._logq_cache
        if c is not None and c[0] is given and c[1] == n_event:
            return c[2]
        return None

    def log_prob(self, given):
        """Log density / mass at `given`, summed over the last ``group_ndims`` axes."""
        return self._log_prob_event(given, self._group_ndims)

    def _given(self, given):
        if given is None:
            given = self.sample_cache
            if given is None:
                raise ValueError("log_prob(None) needs a cached sample: draw or observe a value first")
            return given
        return torch.as_tensor(given, dtype=self.dtype)

    def _log_prob_event(self, given, n_event):
        """Default: elementwise ``_log_prob`` then a torch sum.  Hot distributions override this with
        a single fused kernel."""
        log_p = self._log_prob(self._given(given))
        if n_event > 0:
            return torch.sum(log_p, [i for i in range(-n_event, 0)])
        return log_p

    def _log_prob(self, given):
        raise NotImplementedError()

    def prob(self, given):
        return self._prob(given)

    def _prob(self, given):
        raise NotImplementedError()
