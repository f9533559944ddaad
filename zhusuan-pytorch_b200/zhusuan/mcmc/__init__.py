from .SGMCMC import SGMCMC
from .SGLD import SGLD, PSGLD
from .SGHMC import SGHMC
