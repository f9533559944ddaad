"""zhusuan — B200-native drop-in for the multi-particle stochastic-node hot path of
ZhuSuan-PyTorch (module layout and public names follow zhusuan/__init__.py:3-7 of the reference).

Only the hot path is provided: Normal / Bernoulli / Categorical stochastic nodes, BayesianNet,
ELBO / ImportanceWeightedObjective, SGLD / PSGLD / SGHMC.  Every numeric op runs in hand-written
sm_100a CUDA kernels behind a C ABI (include/zs_b200.h); there is no CPU fallback.
"""
__version__ = '0.0.1'

from . import distributions
from . import framework
from .utils import *
