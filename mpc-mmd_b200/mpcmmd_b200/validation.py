"""Monte-Carlo validation of the planned trajectories: the reference's `validation.py` workflow (synthetic_static_obs/validation.py,
synthetic_dynamic_obs/validation.py) with the 1000-rollout evaluation of `compute_stats` on the GPU.

What stays on the host, and why: the reference seeds NumPy's legacy global stream per trajectory (`np.random.seed(key)`, validation.py:43)
and draws `multivariate_normal` / `beta` from it (:65-84).  Bit-identical collision counts need those exact MT19937 draws, so the noise is
drawn here with the same calls in the same order, the controls are perturbed in float64 exactly as the reference does, and the perturbed
controls go to `mpcmmd_validate_host` (csrc/k_validate.cuh), which rolls the 1000 bicycle models out and counts intersections.

Same command line as the reference, same input files (`./data/{noise}_noise/noise_{L}/ts_{np}/{cost}_{nr}_samples_{O}_obs.npz`), same output
(`./stats/{noise}_noise/noise_{L}/ts_{np}/{nr}_samples_{O}_obs.npz` with coll_cvar, coll_cvar_lane, coll_mmd_opt, coll_mmd_opt_lane,
coll_mmd_random, coll_mmd_random_lane).  Trajectory pairs are enumerated exactly like the reference does -- `set(cvar rows) & set(mmd_opt rows)`
in Python set-iteration order, position k in that order is the RNG seed of the pair -- so the per-pair draws are the same.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os

import numpy as np

from . import binding as B
from .cem_impl import CEM

NUM_ROLLOUTS = 1000          # _num_batch, validation.py:173


def compute_controls(prob, cx, cy):
    """acc (101,), steer (100,) of the planned trajectory in float64 (validation.py:126-132,141-148): np.dot of the float32 basis with the
    float64 coefficients promotes to float64."""
    cx, cy = np.asarray(cx, np.float64).reshape(-1), np.asarray(cy, np.float64).reshape(-1)
    Pdot, Pddot = np.asarray(prob.Pdot_jax), np.asarray(prob.Pddot_jax)
    xdot, xddot = np.dot(Pdot, cx), np.dot(Pddot, cx)
    ydot, yddot = np.dot(Pdot, cy), np.dot(Pddot, cy)
    v = np.sqrt(xdot ** 2 + ydot ** 2)
    v = np.hstack((v, v[-1]))
    acc = np.diff(v) / prob.t
    acc = np.hstack((acc, acc[-1]))
    curvature = (yddot * xdot - ydot * xddot) / ((xdot ** 2 + ydot ** 2) ** (1.5))
    steer = np.arctan(curvature * prob.wheel_base)
    return acc, steer


def perturbed_controls(prob, acc, steer, noise_level, num_prime, noise, key, n_roll=NUM_ROLLOUTS):
    """the noisy control sequences of `compute_rollout_complete` (validation.py:40-90): legacy-stream draws in the reference's order,
    float64.  acc, steer: the first num_prime planned controls.  Returns (n_roll, num_prime) x2."""
    np.random.seed(key)
    if noise == "gaussian":
        z_acc = np.random.multivariate_normal(np.zeros(num_prime), np.eye(num_prime), (n_roll,))
        z_steer = np.random.multivariate_normal(np.zeros(num_prime), np.eye(num_prime), (n_roll,))
        acc_pert = noise_level * np.abs(acc) * z_acc
        steer_pert = noise_level * np.abs(steer) * z_steer
    else:
        b_acc = np.random.beta(prob.beta_a * np.abs(acc), prob.beta_b * np.abs(acc), (n_roll, num_prime))
        b_steer = np.random.beta(prob.beta_a * np.abs(steer) + 1e-5, prob.beta_b * np.abs(steer) + 1e-5, (n_roll, num_prime))
        acc_pert = noise_level * (2 * b_acc - 1)
        steer_pert = prob.cem_helper.K_steer * noise_level * (2 * b_steer - 1)
    z = np.random.multivariate_normal(np.zeros(num_prime), np.eye(num_prime), (n_roll,))
    return acc + acc_pert + prob.acc_const_noise * z, steer + steer_pert + prob.steer_const_noise * z


def compute_stats_batch(prob, items, num_prime, noise_level, noise, num_obs, want_rollouts=False, device=None, n_roll=NUM_ROLLOUTS):
    """`compute_stats` (validation.py:134-171) for a list of trajectories in ONE device call.

    items: dicts with cx, cy (11,), init_state (6,), key, and either x_obs, y_obs, vx_obs, vy_obs (static variant) or
    x_obs_traj, y_obs_traj (num_obs, 100) (dynamic variant).  Returns count (n,), count_lane (n,) int arrays [, x_roll, y_roll]."""
    n = len(items)
    if n == 0:
        z = np.zeros(0, np.int64)
        return (z, z, None, None) if want_rollouts else (z, z)
    acc_all = np.empty((n, n_roll, num_prime)); steer_all = np.empty((n, n_roll, num_prime))
    state0 = np.empty((n, 5)); xo = np.empty((n, num_obs, num_prime)); yo = np.empty((n, num_obs, num_prime))
    static = "x_obs_traj" not in items[0]
    for i, it in enumerate(items):
        acc, steer = compute_controls(prob, it["cx"], it["cy"])
        acc_all[i], steer_all[i] = perturbed_controls(prob, acc[0:num_prime], steer[0:num_prime], noise_level, num_prime, noise, it["key"], n_roll)
        s = np.asarray(it["init_state"], np.float64).reshape(-1)
        state0[i] = [s[0], s[1], s[2], s[3], np.arctan2(s[3], s[2])]
        if static:
            x_obs, y_obs = np.asarray(it["x_obs"]).reshape(-1), np.asarray(it["y_obs"]).reshape(-1)
            vx_obs, vy_obs = np.asarray(it["vx_obs"]).reshape(-1), np.asarray(it["vy_obs"]).reshape(-1)
            xt, yt, _ = prob.cem_helper.compute_obs_trajectories(x_obs, y_obs, vx_obs, vy_obs, np.arctan2(vy_obs, vx_obs))   # float32
        else:
            xt = np.asarray(it["x_obs_traj"]).reshape(num_obs, prob.num); yt = np.asarray(it["y_obs_traj"]).reshape(num_obs, prob.num)
        xo[i] = xt[:, 0:num_prime]; yo[i] = yt[:, 0:num_prime]
    count = np.zeros(n, np.int32); lane = np.zeros(n, np.int32)
    xr = np.empty((n, n_roll, num_prime)) if want_rollouts else None
    yr = np.empty((n, n_roll, num_prime)) if want_rollouts else None
    dev = prob.device.index if device is None else int(device)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    B.check(B.load().mpcmmd_validate_host(dev or 0, n, n_roll, num_prime, num_obs, 1 if static else 0, float(prob.t), float(prob.wheel_base),
                                          float(prob.a_obs), float(prob.b_obs), float(prob.y_lb), float(prob.y_ub), p(acc_all), p(steer_all),
                                          p(state0), p(xo), p(yo), p(count), p(lane), p(xr), p(yr)))
    return (count, lane, xr, yr) if want_rollouts else (count, lane)


def _load(root, noise, noise_level, num_prime, cost, num_reduced, num_obs):
    return np.load(root + "/{}_noise/noise_{}/ts_{}/{}_{}_samples_{}_obs.npz".format(noise, int(noise_level * 100), num_prime, cost, num_reduced, num_obs))


def _matrix(d, num_obs):
    return np.hstack((np.asarray(d["init_state"]), np.asarray(d["x_obs"])[:, 0:num_obs], np.asarray(d["y_obs"])[:, 0:num_obs],
                      np.asarray(d["vx_obs"])[:, 0:num_obs], np.asarray(d["vy_obs"])[:, 0:num_obs]))


def matched_pairs(data_cvar, data_mmd_opt, num_obs):
    """[(k, idx_cvar, idx_mmd_opt)]: scenes solved by BOTH costs, in the reference's order (validation.py:290-313): Python set iteration
    order of `cset & dset`; the first matching row of each file."""
    cvar_matrix, mmd_opt_matrix = _matrix(data_cvar, num_obs), _matrix(data_mmd_opt, num_obs)
    cset = set([tuple(x) for x in cvar_matrix])
    dset = set([tuple(x) for x in mmd_opt_matrix])
    eset = np.array([x for x in cset & dset])
    out = []
    for k in range(0, eset.shape[0]):
        idx_cvar = np.where(np.all(eset[k] == cvar_matrix, axis=1))[0]
        idx_mmd_opt = np.where(np.all(eset[k] == mmd_opt_matrix, axis=1))[0]
        out.append((k, int(idx_cvar[0]), int(idx_mmd_opt[0])))
    return out


def _item(d, idx, key, dynamic):
    it = dict(cx=np.asarray(d["cx"])[idx], cy=np.asarray(d["cy"])[idx], init_state=np.asarray(d["init_state"])[idx], key=key)
    if dynamic:
        it.update(x_obs_traj=np.asarray(d["x_obs_traj"])[idx], y_obs_traj=np.asarray(d["y_obs_traj"])[idx])
    else:
        it.update(x_obs=np.asarray(d["x_obs"])[idx], y_obs=np.asarray(d["y_obs"])[idx], vx_obs=np.asarray(d["vx_obs"])[idx], vy_obs=np.asarray(d["vy_obs"])[idx])
    return it


def run_validation(args, variant="static", device=0, log=print):
    written = []
    dynamic = variant == "dynamic"
    for noise in args.noises:
        for noise_level in args.noise_levels:
            for num_prime in args.num_prime:
                for num_obs in args.num_obs:
                    for num_reduced in args.num_reduced_sets:
                        prob = CEM(num_reduced, num_obs, noise_level, num_prime, noise, args.acc_const_noise, args.steer_const_noise,
                                   variant=variant, max_episodes=1, device=device)
                        data_mmd_opt = _load(args.root, noise, noise_level, num_prime, "mmd_opt", num_reduced, num_obs)
                        data_cvar = _load(args.root, noise, noise_level, num_prime, "cvar", num_reduced, num_obs)
                        pairs = matched_pairs(data_cvar, data_mmd_opt, num_obs)
                        log((len(pairs), _matrix(data_cvar, num_obs).shape[1]))                     # the reference prints eset.shape
                        # both costs of a scene share the seed k (validation.py:315-336)
                        items = [_item(data_mmd_opt, im, k, dynamic) for k, _, im in pairs] + [_item(data_cvar, ic, k, dynamic) for k, ic, _ in pairs]
                        count, lane = compute_stats_batch(prob, items, num_prime, noise_level, noise, num_obs)
                        m = len(pairs)
                        f = lambda a: np.asarray(a, np.float64)                                   # np.append([], int) yields float64 arrays
                        path = "{}/{}_noise/noise_{}/ts_{}/{}_samples_{}_obs".format(args.stats_root, noise, int(noise_level * 100), num_prime, num_reduced, num_obs)
                        os.makedirs(os.path.dirname(path), exist_ok=True)
                        np.savez(path, coll_cvar=f(count[m:]), coll_cvar_lane=f(lane[m:]), coll_mmd_opt=f(count[:m]), coll_mmd_opt_lane=f(lane[:m]),
                                 coll_mmd_random=[], coll_mmd_random_lane=[])
                        written.append(path + ".npz")
                        del prob
    return written


def build_parser():
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument("--noise_levels", type=float, nargs="+", required=True)
    p.add_argument("--num_reduced_sets", type=int, nargs="+", required=True)
    p.add_argument("--num_obs", type=int, nargs="+", required=True)
    p.add_argument("--num_prime", type=int, nargs="+", required=True)
    p.add_argument("--noises", type=str, nargs="+", required=True)
    p.add_argument("--acc_const_noise", type=float, required=True)
    p.add_argument("--steer_const_noise", type=float, required=True)
    p.add_argument("--root", type=str, default="./data")
    p.add_argument("--stats_root", type=str, default="./stats")
    p.add_argument("--variant", type=str, default="static", choices=["static", "dynamic"])
    return p


def main(argv=None, variant=None):
    args = build_parser().parse_args(argv)
    run_validation(args, variant or args.variant)


if __name__ == "__main__":
    main()
