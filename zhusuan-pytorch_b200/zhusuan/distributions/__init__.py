"""Hot-path distributions.  The reference's other wrappers over torch.distributions (beta, gamma,
laplace, ...; zhusuan/distributions/__init__.py:3-13) are outside the accelerated path and are not
rebuilt here."""
from .base import *
from .normal import *
from .bernoulli import *
from .categorical import *
