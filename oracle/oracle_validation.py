"""CPU oracle of the Monte-Carlo validation (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

NumPy float64 restatement of `compute_stats` and its helpers:
    compute_rollout_one_step   S/validation.py:21-38        compute_rollout_complete   S/validation.py:40-105
    compute_f_bar_temp         S/validation.py:107-114      compute_lane_bar           S/validation.py:116-124
    compute_controls           S/validation.py:126-132      compute_stats              S/validation.py:134-171 (D/validation.py:129-165)
Pinned by tests/golden/validation_ref.npz, which tests/golden/make_golden_validation.py produces by executing the reference's OWN
function bodies (extracted from validation.py with `ast`, since the module runs argparse and imports matplotlib at import time).

dtype notes (what real JAX does at these call sites): `np.dot(prob.Pdot_jax, cx)` converts the float32 jax array through `__array__` and
returns a plain float64 ndarray, so controls and rollouts are float64.  In the STATIC variant `x_obs_traj` is a float32 jax array and
`x - x_obs[...]` dispatches to jax (`__array_priority__`), which rounds the float64 rollouts to float32: the obstacle cost is float32
there.  In the DYNAMIC variant the trajectories are NumPy float64 arrays from the .npz and the cost is float64.
"""
import numpy as np

f32 = np.float32
NUM_ROLLOUTS = 1000


def controls(Pdot32, Pddot32, cx, cy, t, wheel_base):
    xdot, xddot = np.dot(Pdot32, cx), np.dot(Pddot32, cx)
    ydot, yddot = np.dot(Pdot32, cy), np.dot(Pddot32, cy)
    v = np.sqrt(xdot ** 2 + ydot ** 2)
    v = np.hstack((v, v[-1]))
    acc = np.diff(v) / t
    acc = np.hstack((acc, acc[-1]))
    steer = np.arctan(((yddot * xdot - ydot * xddot) / ((xdot ** 2 + ydot ** 2) ** 1.5)) * wheel_base)
    return acc, steer


def noisy_controls(acc, steer, noise_level, num_prime, noise, key, K_steer, acc_const, steer_const, beta_a=2, beta_b=5, n_roll=NUM_ROLLOUTS):
    np.random.seed(key)
    eye, zero = np.eye(num_prime), np.zeros(num_prime)
    if noise == "gaussian":
        za = np.random.multivariate_normal(zero, eye, (n_roll,)); zs = np.random.multivariate_normal(zero, eye, (n_roll,))
        pa, ps = noise_level * np.abs(acc) * za, noise_level * np.abs(steer) * zs
    else:
        ba = np.random.beta(beta_a * np.abs(acc), beta_b * np.abs(acc), (n_roll, num_prime))
        bs = np.random.beta(beta_a * np.abs(steer) + 1e-5, beta_b * np.abs(steer) + 1e-5, (n_roll, num_prime))
        pa, ps = noise_level * (2 * ba - 1), K_steer * noise_level * (2 * bs - 1)
    z = np.random.multivariate_normal(zero, eye, (n_roll,))
    return acc + pa + acc_const * z, steer + ps + steer_const * z


def rollouts(acc, steer, state0, t, wheel_base):
    n, num_prime = acc.shape
    x_roll, y_roll = np.zeros((n, num_prime)), np.zeros((n, num_prime))
    x, y, vx, vy, psi = (np.full(n, float(s)) for s in state0)
    for i in range(num_prime):
        x_roll[:, i], y_roll[:, i] = x, y
        v = np.sqrt(vx ** 2 + vy ** 2)
        v = v + acc[:, i] * t
        psi = psi + (v * np.tan(steer[:, i]) / wheel_base) * t
        vx, vy = v * np.cos(psi), v * np.sin(psi)
        x, y = x + vx * t, y + vy * t
    return x_roll, y_roll


def counts(x_roll, y_roll, x_obs_traj, y_obs_traj, num_prime, a_obs, b_obs, y_lb, y_ub, obs_f32):
    xo, yo = np.asarray(x_obs_traj)[:, 0:num_prime][:, None], np.asarray(y_obs_traj)[:, 0:num_prime][:, None]
    if obs_f32:
        wc, ws = x_roll.astype(f32) - xo.astype(f32), y_roll.astype(f32) - yo.astype(f32)
        cost = -(wc * wc) / f32(a_obs ** 2) - (ws * ws) / f32(b_obs ** 2) + f32(1.0)
    else:
        wc, ws = x_roll - xo, y_roll - yo
        cost = -(wc ** 2) / (a_obs ** 2) - (ws ** 2) / (b_obs ** 2) + 1.0
    bar = np.maximum(np.zeros_like(cost), cost).transpose(0, 2, 1)             # num_obs x timesteps x rollouts
    count = int(np.max(np.max(np.count_nonzero(bar, axis=2), axis=1)))
    lb, ub = np.maximum(0.0, -y_roll + y_lb).T, np.maximum(0.0, y_roll - y_ub).T
    count_lane = int(np.max(np.count_nonzero(lb, axis=1))) + int(np.max(np.count_nonzero(ub, axis=1)))
    return count, count_lane


def compute_stats(c, cx, cy, init_state, x_obs_traj, y_obs_traj, num_prime, noise_level, noise, key, obs_f32, n_roll=NUM_ROLLOUTS):
    """c: dict(Pdot, Pddot (float32), t, wheel_base, a_obs, b_obs, y_lb, y_ub, K_steer, acc_const, steer_const)"""
    cx, cy = np.asarray(cx, np.float64).reshape(-1), np.asarray(cy, np.float64).reshape(-1)
    s = np.asarray(init_state, np.float64).reshape(-1)
    state0 = np.asarray([s[0], s[1], s[2], s[3], np.arctan2(s[3], s[2])])
    acc, steer = controls(c["Pdot"], c["Pddot"], cx, cy, c["t"], c["wheel_base"])
    a, st = noisy_controls(acc[0:num_prime], steer[0:num_prime], noise_level, num_prime, noise, key, c["K_steer"], c["acc_const"], c["steer_const"], n_roll=n_roll)
    x_roll, y_roll = rollouts(a, st, state0, c["t"], c["wheel_base"])
    count, count_lane = counts(x_roll, y_roll, x_obs_traj, y_obs_traj, num_prime, c["a_obs"], c["b_obs"], c["y_lb"], c["y_ub"], obs_f32)
    return count, count_lane, x_roll, y_roll
