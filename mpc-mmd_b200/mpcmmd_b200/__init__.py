"""mpcmmd_b200 -- host side of the B200-native MPC-MMD trajectory optimizer (libmpcmmd.so + drop-in CEM class)."""
from .cem_impl import CEM  # noqa: F401
