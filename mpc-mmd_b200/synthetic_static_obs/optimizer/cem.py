"""Drop-in for the reference's `synthetic_static_obs/optimizer/cem.py`:

    sys.path.insert(1, "<repo>/mpc-mmd_b200/synthetic_static_obs")     # reference: sys.path.insert(1, 'path/to/optimizer')
    from optimizer import cem                                       # synthetic_static_obs/main_mpc.py:5-6
    prob = cem.CEM(num_reduced, num_obs, noise_level, num_prime, noise, acc_const_noise, steer_const_noise)

Same class name, constructor, `compute_cem_{mmd_opt,mmd_random,cvar,saa}` methods and attributes; the
"static" constants (y_lb/y_ub at cem.py:155, K_steer at cem_helper.py:24) are selected here.
"""
import os
import sys

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _PKG_ROOT not in sys.path:
    sys.path.insert(1, _PKG_ROOT)

from mpcmmd_b200.cem_impl import CEM as _CEM  # noqa: E402


class CEM(_CEM):
    def __init__(self, num_reduced, num_obs, noise_level, num_prime, noise, acc_const_noise, steer_const_noise, **kw):
        kw.setdefault("variant", "static")
        super().__init__(num_reduced, num_obs, noise_level, num_prime, noise, acc_const_noise, steer_const_noise, **kw)
