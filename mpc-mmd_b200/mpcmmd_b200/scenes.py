"""Synthetic episode inputs of the reference drivers, restated for the harness / bench:
obstacle scenes of `compute_obs_data` (synthetic_static_obs/main_mpc.py:10-21), the fixed initial state / CEM mean /
covariance / v_des (main_mpc.py:46-74; synthetic_dynamic_obs/main_mpc.py:34-62) and the per-episode `idx_mpc`
draw (`np.random.randint(1,10000)` right after the scene draws, main_mpc.py:113-119)."""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from . import jaxrng
from .constants import NUM, NVAR, T_FIN, bernstein_coeff_order10_new

F32 = np.float32
_X_CHOICES = np.array([35, 40, 45, 50, 55, 60, 65, 70, 75])
_Y_CHOICES = np.array([-1.75, 1.75])


def static_scene(num_obs: int, k: int):
    """episode k of the static sweep: (x_obs, y_obs, vx_obs, vy_obs, psi_obs), idx_mpc -- legacy NumPy global stream seeded with k"""
    np.random.seed(k)
    x = np.random.choice(_X_CHOICES, (num_obs,), replace=False)
    y = np.random.choice(_Y_CHOICES, (num_obs,))
    idx_mpc = int(np.random.randint(1, 10000))
    z = np.zeros(num_obs)
    return (x, y, z.copy(), z.copy(), z.copy()), idx_mpc


def scaled_scene(num_obs: int, k: int):
    """harness-defined scene for num_obs > 9 (the reference's `choice(..., replace=False)` cannot draw more than 9):
    x uniform on [15, 75], lane +-1.75, static obstacles, Generator seeded with the episode index."""
    g = np.random.default_rng(k)
    x = g.uniform(15.0, 75.0, num_obs)
    y = g.choice(_Y_CHOICES, num_obs)
    z = np.zeros(num_obs)
    return (x, y, z.copy(), z.copy(), z.copy()), int(g.integers(1, 10000))


@lru_cache(maxsize=1)
def _obs_guess_maps():
    """Constant affine maps of `obs_data.compute_obs_guess` (synthetic_dynamic_obs/obs_data_generate_dynamic.py:71-108): whole-horizon
    PD tracking QP (k_p_v = k_p = 2, smoothness weight 100) with the boundary rows A_eq_x (3) / A_eq_y (4); the KKT matrices are
    constant, so they are inverted once in float64 (the reference LU-solves them in float32 on every call)."""
    tot_time = np.linspace(0, T_FIN, NUM)
    P64, Pd64, Pdd64 = bernstein_coeff_order10_new(10, tot_time[0], tot_time[-1], tot_time)
    P, Pd, Pdd = (a.astype(F32).astype(np.float64) for a in (P64, Pd64, Pdd64))          # jnp.asarray -> float32 (:19)
    A_eq_x = np.vstack((P[0], Pd[0], Pdd[0])); A_eq_y = np.vstack((P[0], Pd[0], Pdd[0], Pd[-1]))
    A_vd, A_pd = Pdd - 2.0 * Pd, Pdd - 2.0 * P
    Mx = np.linalg.inv(np.block([[100.0 * Pdd.T @ Pdd + A_vd.T @ A_vd, A_eq_x.T], [A_eq_x, np.zeros((3, 3))]]))
    My = np.linalg.inv(np.block([[100.0 * Pdd.T @ Pdd + A_pd.T @ A_pd, A_eq_y.T], [A_eq_y, np.zeros((4, 4))]]))
    # -lincost = A^T b with b = -k_p * des * 1  (:78-92)  ->  c = M[:11,:11] @ (-(-2) ... ) folded below
    gx = Mx[:NVAR, :NVAR] @ (-2.0 * A_vd.T @ np.ones(NUM)); gy = My[:NVAR, :NVAR] @ (-2.0 * A_pd.T @ np.ones(NUM))
    return P, gx, Mx[:NVAR, NVAR:], gy, My[:NVAR, NVAR:]


def dynamic_scene(num_obs: int, k: int):
    """episode k of the dynamic sweep (synthetic_dynamic_obs/main_mpc.py:108-126, obs_data_generate_dynamic.py:112-148):
    returns (x_obs, y_obs, vx_obs, vy_obs, psi_obs) initial values, idx_mpc, x_obs_traj (num_obs,100), y_obs_traj (num_obs,100).
    Obstacle tt starts in the left lane (y = 1.75, cut-in scenario) and tracks lane -1.75 at speed 6 + 0.1 N(0,1), the normal
    drawn with `PRNGKey(43 k + 11 tt + 5)`."""
    key = jaxrng.prng_key(k)
    x0 = jaxrng.choice_no_replace(key, jaxrng.linspace32(15, 45, 30), num_obs).astype(F32)
    vx0 = jaxrng.choice_no_replace(key, jaxrng.linspace32(0.5, 5, 15), num_obs).astype(F32)
    y0 = (F32(1.75) * np.ones(num_obs, F32)).astype(F32)
    z = np.zeros(num_obs, F32)
    P, gx, Bx, gy, By = _obs_guess_maps()
    xt, yt = np.zeros((num_obs, NUM), F32), np.zeros((num_obs, NUM), F32)
    for tt in range(num_obs):
        v_des = F32(jaxrng.normal(jaxrng.prng_key(43 * k + 11 * tt + 5), 1)[0] * F32(0.1) + F32(6.0))      # sampling_param (:112-116)
        cx = gx * (-float(v_des)) * -1.0 + Bx @ np.array([x0[tt], vx0[tt], 0.0], np.float64)
        cy = gy * (-(-1.75)) * -1.0 + By @ np.array([y0[tt], 0.0, 0.0, 0.0], np.float64)
        xt[tt] = (P @ cx).astype(F32); yt[tt] = (P @ cy).astype(F32)
    np.random.seed(k)                                                                  # main_mpc.py:114
    idx_mpc = int(np.random.randint(1, 10000))                                         # first legacy draw after the seed (:128-133)
    return (x0, y0, vx0, z.copy(), z.copy()), idx_mpc, xt, yt


def driver_inputs(variant: str = "static"):
    y0 = 1.75 if variant == "static" else -1.75
    init_state = np.array([0.0, y0, 5.0, 0.0, 0.0, 0.0], F32)       # x, y, vx, vy, ax, ay
    mean = np.array([15.0] * 4 + [0.0] * 4, F32)
    cov = np.diag(np.array([20.0] * 4 + [100.0] * 4)).astype(F32)
    return init_state, mean, cov, 15.0


def static_batch(prob, episodes, variant="static"):
    """stacked solve_batch inputs for the given episode indices (variant "dynamic": the cut-in scenes of `dynamic_scene`)"""
    init_state, mean, cov, v_des = driver_inputs(variant)
    idx, xo, yo = [], [], []
    for k in episodes:
        if variant == "dynamic":
            _, i, x, y = dynamic_scene(prob.num_obs, k)
        else:
            sc, i = static_scene(prob.num_obs, k) if prob.num_obs <= 9 else scaled_scene(prob.num_obs, k)
            x, y, _ = prob.cem_helper.compute_obs_trajectories(*sc)
        idx.append(i); xo.append(x); yo.append(y)
    E = len(idx)
    if E == 0:                      # a rank whose shard is empty (more ranks than episodes) still takes part in the gather
        xo = [np.zeros((0, prob.num_obs, NUM), F32)]; yo = [np.zeros((0, prob.num_obs, NUM), F32)]
        return dict(idx_mpc=np.zeros(0, np.int32), init_state=np.zeros((0, 6), F32), mean_param=np.zeros((0, 8), F32), cov_param=np.zeros((0, 8, 8), F32),
                    x_obs_traj=xo[0], y_obs_traj=yo[0], v_des=np.zeros(0, F32))
    return dict(idx_mpc=np.asarray(idx, np.int32), init_state=np.repeat(init_state[None], E, 0), mean_param=np.repeat(mean[None], E, 0),
                cov_param=np.repeat(cov[None], E, 0), x_obs_traj=np.stack(xo).astype(F32), y_obs_traj=np.stack(yo).astype(F32),
                v_des=np.full(E, v_des, F32))


def flops_per_sample(cost: str, num_reduced: int, num_prime: int, num_obs: int, num_samples_cem=100, maxiter_beta_cem=20, survey_count=False):
    """Algorithmic FLOPs (FMA = 2) of ONE CEM sample in ONE outer iteration, split by kernel -- the minimal formulation of
    SURVEY.md section 8(d).  Returns dict(project=..., risk=...).  `survey_count=True` reproduces SURVEY's figure literally (S evaluations in every
    inner iteration); the default counts what the algorithm needs: the elites keep their cost, so S + (iters - 1)(S - ne) rows are evaluated."""
    from math import log2
    nr, np_, O, S, T, n = num_reduced, num_prime, num_obs, num_samples_cem, 100, 11
    nm = nr * nr
    f_guess, f_proj, f_ctrl = 176, 76094, 25 * T
    opt = cost == "mmd_opt"
    R = nm if opt else nr
    f_roll = 30 * R * np_ + 6 * nr * np_
    f_fit = 2 * nm * (2 * np_ * n + 2 * n * n) if opt else 0
    f_rs = 0
    if opt:
        ne = max(int(0.1 * S) + 1, 3)                       # compute_beta.py:26
        f_eval = 2 * nm * log2(nm) + 2 * (nr * nr + nr * nm) + nr * nm + (2.0 / 3.0) * (nr + 1) ** 3 + 2 * (nr + 1) ** 2 + 2 * nr * nr + 2 * nr
        f_upd = 2 * S * log2(S) + ne * (nm + 1) ** 2 + (nm + 1) ** 3 / 3.0 + (S - ne) * (nm + 1) ** 2
        # the elites keep their cost from the iteration that produced them (same row => same value), so only the S - ne resampled rows
        # are evaluated after iteration 0: S + (iters - 1) (S - ne) evaluations, not iters * S
        n_eval = maxiter_beta_cem * S if survey_count else S + (maxiter_beta_cem - 1) * (S - ne)
        f_rs = 33 * nm * nm + n_eval * f_eval + maxiter_beta_cem * f_upd
    f_cost = 8 * nr * O * np_ + 5 * nr * nr + 3 * nr
    return dict(project=float(f_guess + f_proj + f_ctrl), risk=float(f_roll + f_fit + f_rs + f_cost))
