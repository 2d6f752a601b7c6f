"""Generates tests/golden/dynamic_scenes.npz by EXECUTING THE REFERENCE'S OWN scene generator
(synthetic_dynamic_obs/obs_data_generate_dynamic.py, class obs_data) on the NumPy stand-in for the JAX API
(tests/golden/jax_shim, see its _core.py header).  Build container only (needs /root/reference):

    python tests/golden/make_golden_scenes.py

Recorded per (num_obs, k): the initial obstacle states of `compute_obs_data` and the obstacle trajectories the reference's
main loop builds with `compute_boundary_vec` + `compute_obs_guess(..., seed = 43 k + 11 tt + 5)` (D/main_mpc.py:108-126), plus the
`idx_mpc` the loop draws (`np.random.seed(k); np.random.randint(1, 10000)`, :114,128-133).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "jax_shim"))
sys.path.insert(1, "/root/reference/synthetic_dynamic_obs")
sys.path.insert(1, ROOT)

from obs_data_generate_dynamic import obs_data  # noqa: E402  (the reference's module)

CASES = [(6, 0), (6, 1), (6, 17), (6, 199), (3, 5), (2, 42)]


def main():
    od = obs_data(1)
    out = {}
    for num_obs, k in CASES:
        x, y, vx, vy, psi = od.compute_obs_data(num_obs, k)
        np.random.seed(k)
        xt, yt = np.zeros((num_obs, 100), np.float32), np.zeros((num_obs, 100), np.float32)
        for tt in range(num_obs):
            bx, by = od.compute_boundary_vec(x[tt], vx[tt], 0., y[tt], vy[tt], 0.)
            _x, _y = od.compute_obs_guess(bx, by, -1.75 * np.ones(1), 43 * k + 11 * tt + 5)
            xt[tt] = np.asarray(_x).reshape(-1); yt[tt] = np.asarray(_y).reshape(-1)
        idx = np.random.randint(1, 10000)
        tag = f"o{num_obs}_k{k}_"
        for n, v in (("x", x), ("y", y), ("vx", vx), ("vy", vy), ("psi", psi), ("xt", xt), ("yt", yt), ("idx", idx)):
            out[tag + n] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "dynamic_scenes.npz"), **out)
    print("wrote dynamic_scenes.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
