from ._core import jit, vmap, Arr            # noqa: F401
from . import numpy, random, lax            # noqa: F401
__version__ = "0.3.23-numpy-shim"
