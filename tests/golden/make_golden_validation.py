"""Generates tests/golden/validation_ref.npz by EXECUTING THE REFERENCE'S OWN `compute_stats` (and the helpers it calls).

Build container only (needs /root/reference):   python tests/golden/make_golden_validation.py

validation.py runs argparse and imports matplotlib / scipy-spline plotting code at import time, so the module cannot be imported.  The
function definitions `compute_rollout_one_step, compute_rollout_complete, compute_f_bar_temp, compute_lane_bar, compute_controls,
compute_stats` are cut out of the file with `ast` and exec'ed UNMODIFIED in a namespace holding `np`, `_num_batch = 1000` and `prob` = the
reference's own `optimizer.cem.CEM` object, built on the NumPy stand-in for JAX (tests/golden/jax_shim).  One adjustment mirrors real JAX:
`prob.Pdot_jax / Pddot_jax` are handed over as plain float32 ndarrays, because a real DeviceArray is not an ndarray subclass and
`np.dot(DeviceArray, float64 ndarray)` yields a plain float64 ndarray (the shim's ndarray subclass would keep rounding to float32).
Everything else in these functions is NumPy float64 + the legacy MT19937 stream, i.e. the real arithmetic of the reference.

Trajectories fed in: oracle solves of sweep episodes (realistic, mostly collision free) and deliberately bad straight-line plans through
the obstacles (non-zero counts).
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
FUNCS = ("compute_rollout_one_step", "compute_rollout_complete", "compute_f_bar_temp", "compute_lane_bar", "compute_controls", "compute_stats")


def load_reference(variant_dir, cem_args):
    # the reference imports its pieces as TOP-LEVEL modules (`from cem_helper import Helper`): purge them all, or the second variant would
    # silently reuse the first variant's cem_helper (K_steer 0.01 vs 0.05)
    for m in [m for m in sys.modules if m == "optimizer" or m.startswith("optimizer.") or m in (
            "compute_beta", "kernel_computation", "bernstein_coeff_order10_arbitinterval", "cem_helper", "projection", "costs", "cem")]:
        del sys.modules[m]
    ref = os.path.join("/root/reference", variant_dir)
    sys.path[:] = [p for p in sys.path if not p.startswith("/root/reference")]
    sys.path.insert(0, os.path.join(HERE, "jax_shim"))
    sys.path.insert(1, ref); sys.path.insert(1, os.path.join(ref, "optimizer")); sys.path.insert(1, ROOT)
    from optimizer import cem
    prob = cem.CEM(*cem_args)
    assert os.path.dirname(sys.modules["cem_helper"].__file__).startswith(ref), "stale reference module"
    prob.Pdot_jax = np.array(np.asarray(prob.Pdot_jax).view(np.ndarray), dtype=np.float32)
    prob.Pddot_jax = np.array(np.asarray(prob.Pddot_jax).view(np.ndarray), dtype=np.float32)
    src = open(os.path.join(ref, "validation.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "prob": prob, "_num_batch": 1000}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in FUNCS:
            exec(compile(ast.Module([node], []), os.path.join(ref, "validation.py"), "exec"), ns)
    return prob, ns


def straight_plan(P64, x_end, y_lane):
    """Bernstein coefficients of a straight constant-speed line (least squares on the float64 basis)"""
    t = np.linspace(0, 1, 100)
    cx = np.linalg.lstsq(P64, x_end * t, rcond=None)[0]
    cy = np.linalg.lstsq(P64, y_lane * np.ones(100), rcond=None)[0]
    return cx, cy


def main():
    sys.path.insert(1, ROOT); sys.path.insert(1, os.path.join(ROOT, "mpc-mmd_b200"))
    from oracle import oracle as O
    from mpcmmd_b200 import scenes
    out = {}
    cases = []
    # ---- static: cfg2-shaped (beta 0.3, 4 obstacles, num_prime 50) and a gaussian one with common-mode noise
    for tag, args in (("S_beta", (5, 4, 0.3, 50, "beta", 0.0, 0.0)), ("S_gauss", (5, 2, 0.1, 30, "gaussian", 0.05, 0.01))):
        prob, ns = load_reference("synthetic_static_obs", args)
        ora = O.OracleCEM(*args, variant="static")
        init_state, mean, cov, v_des = O.driver_inputs("static")
        for k in (0, 3):
            sc, idx = O.static_episode(args[1], k)
            xt, yt, _ = ora.compute_obs_trajectories(*sc)
            r = ora.solve("cvar", idx, init_state, mean, cov, xt, yt, v_des)
            cases.append((f"{tag}_solve{k}", ns, args, "static", np.asarray(r["cx"], np.float64), np.asarray(r["cy"], np.float64), init_state, sc, None, k))
        sc, _ = O.static_episode(args[1], 1)
        cx, cy = straight_plan(np.asarray(prob.P, np.float64), 75.0, float(sc[1][0]))
        cases.append((f"{tag}_straight", ns, args, "static", cx, cy, init_state, sc, None, 7))
    # ---- dynamic: cfg3-shaped
    for tag, args in (("D_gauss", (5, 6, 0.1, 60, "gaussian", 0.0, 0.0)), ("D_beta", (5, 3, 0.3, 40, "beta", 0.02, 0.0))):
        prob, ns = load_reference("synthetic_dynamic_obs", args)
        ora = O.OracleCEM(*args, variant="dynamic")
        init_state, mean, cov, v_des = O.driver_inputs("dynamic")
        for k in (0, 2):
            sc, idx, xt, yt = scenes.dynamic_scene(args[1], k)
            r = ora.solve("cvar", idx, init_state, mean, cov, xt, yt, v_des)
            cases.append((f"{tag}_solve{k}", ns, args, "dynamic", np.asarray(r["cx"], np.float64), np.asarray(r["cy"], np.float64), init_state, sc, (xt, yt), k))
        sc, idx, xt, yt = scenes.dynamic_scene(args[1], 1)
        cx, cy = straight_plan(np.asarray(prob.P, np.float64), 60.0, 0.0)
        cases.append((f"{tag}_straight", ns, args, "dynamic", cx, cy, init_state, sc, (xt, yt), 5))
    names = []
    for name, ns, args, variant, cx, cy, init_state, sc, traj, key in cases:
        nr, num_obs, noise_level, num_prime, noise, acn, scn = args
        ist = np.asarray(init_state, np.float64)
        if variant == "static":
            x, y, vx, vy = (np.asarray(a, np.float64) for a in sc[:4])
            count, count_lane, x_roll, y_roll, _, _ = ns["compute_stats"](cx, cy, ist, x, y, vx, vy, num_prime, noise_level, noise, num_obs, key)
            out[name + "_x_obs"] = x; out[name + "_y_obs"] = y; out[name + "_vx_obs"] = vx; out[name + "_vy_obs"] = vy
        else:
            xt, yt = (np.asarray(a, np.float64) for a in traj)            # what np.load of the reference's .npz returns
            count, count_lane, x_roll, y_roll, _, _ = ns["compute_stats"](cx, cy, ist, xt, yt, num_prime, noise_level, noise, num_obs, key)
            out[name + "_x_obs_traj"] = xt.astype(np.float32); out[name + "_y_obs_traj"] = yt.astype(np.float32)
        out[name + "_cx"] = cx; out[name + "_cy"] = cy; out[name + "_init_state"] = ist
        out[name + "_args"] = np.array([nr, num_obs, noise_level, num_prime, 0 if noise == "gaussian" else 1, acn, scn, key], np.float64)
        out[name + "_count"] = np.array([int(count), int(count_lane)])
        out[name + "_roll_head"] = np.stack([np.asarray(x_roll)[:4], np.asarray(y_roll)[:4]])           # first 4 rollouts, full horizon
        out[name + "_roll_sum"] = np.array([np.asarray(x_roll).sum(), np.asarray(y_roll).sum()])
        names.append(name)
        print(name, "count", int(count), "lane", int(count_lane))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "validation_ref.npz"), **out)
    print("wrote validation_ref.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
