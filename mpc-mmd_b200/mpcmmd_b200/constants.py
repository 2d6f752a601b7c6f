"""Host-side constants of the trajectory optimizer (what ``CEM.__init__`` builds in the reference,
synthetic_static_obs/optimizer/cem.py:16-199), evaluated once in float64 NumPy and handed to
libmpcmmd.so as float32 matrices.

Folded constant solves (mathematically identical to the reference, which re-factorises the same
constant matrices in float32 on every call):
  * x_guess KKT systems (cem_helper.py:169-230)  ->  affine maps Gx (11,7), Gy (11,8)
  * projection KKT systems (projection.py:145-168) -> inverse rows Kx (11,14), Ky (11,15)
  * ridge fit of rollouts (cem_helper.py:553-564)  -> Wfit (11, num_prime)
"""
from __future__ import annotations

from dataclasses import dataclass
from math import comb

import numpy as np

NUM, NVAR, T_FIN = 100, 11, 15           # cem.py:37-38,50
F32 = np.float32

# the two copies of optimizer/ in the reference differ only in these constants
VARIANT_CONSTANTS = {
    "static": {"y_lb": -2.25, "y_ub": 2.25, "K_steer": 0.01},     # synthetic_static_obs/optimizer/cem.py:155, cem_helper.py:24
    "dynamic": {"y_lb": -2.25, "y_ub": -1.25, "K_steer": 0.05},   # synthetic_dynamic_obs/optimizer/cem.py:155, cem_helper.py:24
}


def _basis(n: int, k: int, t: np.ndarray) -> np.ndarray:
    """Bernstein polynomial B_{k,n}(t); zero outside 0 <= k <= n."""
    if k < 0 or k > n:
        return np.zeros_like(t)
    return comb(n, k) * (1.0 - t) ** (n - k) * t ** k


def bernstein_coeff_order10_new(n, tmin, tmax, t_actual):
    """Same contract as bernstein_coeff_order10_arbitinterval.bernstein_coeff_order10_new (:13-103):
    returns P, Pdot, Pddot of shape (len(t), n+1) on [tmin, tmax]; derivatives via the Bernstein
    difference identities instead of the reference's expanded polynomials."""
    t_actual = np.asarray(t_actual, dtype=np.float64).reshape(-1)
    span = float(np.asarray(tmax).reshape(-1)[0]) - float(np.asarray(tmin).reshape(-1)[0])
    t = (t_actual - float(np.asarray(tmin).reshape(-1)[0])) / span
    P = np.stack([_basis(n, k, t) for k in range(n + 1)], axis=1)
    Pdot = np.stack([n * (_basis(n - 1, k - 1, t) - _basis(n - 1, k, t)) for k in range(n + 1)], axis=1) / span
    Pddot = np.stack([n * (n - 1) * (_basis(n - 2, k - 2, t) - 2.0 * _basis(n - 2, k - 1, t) + _basis(n - 2, k, t))
                      for k in range(n + 1)], axis=1) / (span ** 2)
    return P, Pdot, Pddot


def _integer_pow_f32(x: np.ndarray, y: int) -> np.ndarray:
    """x**y for float32 by binary exponentiation (how `jnp_array ** int` is lowered)."""
    x = x.astype(F32)
    if y == 0:
        return np.ones_like(x)
    acc = None
    while y > 0:
        if y & 1:
            acc = x if acc is None else (acc * x).astype(F32)
        y >>= 1
        if y > 0:
            x = (x * x).astype(F32)
    return acc


def bernstein_P_prime_f32(num_prime: int, t_fin_prime: float, n: int = 10) -> np.ndarray:
    """P_prime of cem_helper.py:112-118: the reference evaluates it through jnp, i.e. in float32."""
    div = num_prime - 1
    step = (np.arange(div, dtype=F32) / F32(div)).astype(F32)
    stop = F32(t_fin_prime)
    tt = (F32(0.0) * (F32(1.0) - step) + stop * step).astype(F32)
    tt = np.concatenate([tt, np.array([stop], dtype=F32)])
    span = F32(tt[-1] - tt[0])
    t = ((tt - tt[0]) / span).astype(F32)
    one_minus_t = (F32(1.0) - t).astype(F32)
    cols = [((F32(comb(n, k)) * _integer_pow_f32(one_minus_t, n - k)).astype(F32) * _integer_pow_f32(t, k)).astype(F32)
            for k in range(n + 1)]
    return np.stack(cols, axis=1).astype(F32)


@dataclass
class HostConstants:
    P: np.ndarray
    Pdot: np.ndarray
    Pddot: np.ndarray
    P64: np.ndarray
    Pdot64: np.ndarray
    Pddot64: np.ndarray
    Gx: np.ndarray
    Gy: np.ndarray
    Kx: np.ndarray
    Ky: np.ndarray
    Wfit: np.ndarray
    P_prime: np.ndarray
    tot_time: np.ndarray


def build_constants(num_prime: int) -> HostConstants:
    tot_time = np.linspace(0, T_FIN, NUM)                                               # cem.py:42
    P64, Pd64, Pdd64 = bernstein_coeff_order10_new(10, tot_time[0], tot_time[-1], tot_time)   # cem.py:46
    P, Pd, Pdd = (np.ascontiguousarray(a.astype(F32)) for a in (P64, Pd64, Pdd64))     # cem.py:48 (jnp.asarray -> f32)
    Pf, Pdf, Pddf = (a.astype(np.float64) for a in (P, Pd, Pdd))
    A_eq_x = np.vstack((Pf[0], Pdf[0], Pddf[0]))                                        # cem.py:55
    A_eq_y = np.vstack((Pf[0], Pdf[0], Pddf[0], Pdf[-1]))                               # cem.py:56

    # x_guess: k_p_v = k_p = 2, smoothness weight 100, four 25-knot quarters (cem_helper.py:183-214)
    Qx = 100.0 * Pddf.T @ Pddf
    Qy = 100.0 * Pddf.T @ Pddf
    gx, gy = [], []
    for q in range(4):
        rows = slice(25 * q, 25 * q + 25)
        A_vd = Pddf[rows] - 2.0 * Pdf[rows]
        A_pd = Pddf[rows] - 2.0 * Pf[rows]
        Qx = Qx + A_vd.T @ A_vd
        Qy = Qy + A_pd.T @ A_pd
        gx.append(-2.0 * A_vd.T @ np.ones(25))
        gy.append(-2.0 * A_pd.T @ np.ones(25))
    Mx = np.linalg.inv(np.block([[Qx, A_eq_x.T], [A_eq_x, np.zeros((3, 3))]]))          # cem_helper.py:216,222
    My = np.linalg.inv(np.block([[Qy, A_eq_y.T], [A_eq_y, np.zeros((4, 4))]]))          # cem_helper.py:217,223
    Gx = np.hstack((Mx[:NVAR, :NVAR] @ np.stack(gx, 1), Mx[:NVAR, NVAR:]))
    Gy = np.hstack((My[:NVAR, :NVAR] @ np.stack(gy, 1), My[:NVAR, NVAR:]))

    # projection: all rho = 1, A_projection = I, A_lane_bound = [P[1:]; -P[1:]] (projection.py:145-155, cem.py:126-134)
    A_lane = np.vstack((Pf[1:], -Pf[1:]))
    cost_x = np.eye(NVAR) + Pddf.T @ Pddf + Pdf.T @ Pdf
    cost_y = cost_x + A_lane.T @ A_lane
    Kx = np.linalg.inv(np.block([[cost_x, A_eq_x.T], [A_eq_x, np.zeros((3, 3))]]))[:NVAR, :]
    Ky = np.linalg.inv(np.block([[cost_y, A_eq_y.T], [A_eq_y, np.zeros((4, 4))]]))[:NVAR, :]

    # ridge fit on the float32 P_prime (cem_helper.py:112-118, 553-564)
    P_prime = bernstein_P_prime_f32(num_prime, num_prime * (T_FIN / NUM))
    Pp = P_prime.astype(np.float64)
    Wfit = np.linalg.inv(Pp.T @ Pp + 0.05 * np.eye(NVAR)) @ Pp.T

    c32 = lambda a: np.ascontiguousarray(a.astype(F32))
    return HostConstants(P=P, Pdot=Pd, Pddot=Pdd, P64=P64, Pdot64=Pd64, Pddot64=Pdd64, Gx=c32(Gx), Gy=c32(Gy), Kx=c32(Kx), Ky=c32(Ky),
                         Wfit=c32(Wfit), P_prime=P_prime, tot_time=tot_time)
