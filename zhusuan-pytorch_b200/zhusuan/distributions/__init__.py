"""Hot-path distributions, plus the two location-scale families of SURVEY 8(f)-4 (Logistic, Laplace) on the same
kernel templates.  and Uniform.  The reference's remaining wrappers over torch.distributions (beta, gamma, poisson, studentT,
exponential; zhusuan/distributions/__init__.py:3-13) are outside the accelerated path and are not rebuilt
here."""
from .base import *
from .normal import *
from .bernoulli import *
from .categorical import *
from .logistic import *
from .laplace import *
from .uniform import *

# Latent nodes draw their sample and its log q in one launch when the node's event reduction allows it
# (zs_normal_latent_fwd / zs_bernoulli_latent_fwd); set to False to always use the separate kernels.
FUSED_LATENT = True
