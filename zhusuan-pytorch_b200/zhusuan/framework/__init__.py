from .bn import *
from .stochastic_tensor import *
