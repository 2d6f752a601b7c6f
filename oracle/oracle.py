"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product).

Python face of the plain-C oracle (oracle_mpc.c): builds the constants of ``CEM.__init__``
(reference S/optimizer/cem.py:16-199) in NumPy, binds liboracle.so through ctypes and restates the
reference's scene generators (S/main_mpc.py:10-21, D/obs_data_generate_dynamic.py) so tests and
bench.py's ``cpu_baseline`` leg can run the same episodes as the CUDA path.

Parity pins (see oracle_mpc.c header / DESIGN.md section 4): RNG known answers, the reference's own Bernstein file, and
tests/golden/ref_stages.npz (stage vectors recorded from the reference's own optimizer source running on a NumPy
stand-in for JAX).  jax.random.beta's rejection sampler is the one piece that remains restated-but-unpinned.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from math import comb

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

T, NV, NL, NPAR = 100, 11, 99, 8
COST_KINDS = {"mmd_opt": 0, "mmd_random": 1, "cvar": 2, "saa": 3}
NOISE_KINDS = {"gaussian": 0, "beta": 1}
f32 = np.float32


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_mpc.c", "oracle_math.h", "oracle_rng.h", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.oracle_sizeof_cfg() == C.sizeof(OCfg), (_lib.oracle_sizeof_cfg(), C.sizeof(OCfg))
        assert _lib.oracle_sizeof_proj() == C.sizeof(OProjOut)
        assert _lib.oracle_sizeof_risk() == C.sizeof(ORiskOut)
        assert _lib.oracle_sizeof_solve() == C.sizeof(OSolveOut)
    return _lib


FP = C.POINTER(C.c_float)


class OCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "np", "nr", "nm", "O", "iters", "n_el", "n_el_cost", "noise_kind",
                                         "S_in", "iters_in", "n_el_in", "naive", "pad_")] + \
               [(n, C.c_float) for n in ("sigma_acc", "sigma_steer", "ksig_steer", "acc_const", "steer_const", "beta_a", "beta_b",
                                         "v_min", "v_max", "a_max", "b_lane_ub", "b_lane_lb", "y_lb", "y_ub", "a2_obs", "b2_obs",
                                         "wheel_base", "dt", "steer_max", "steer_rate_pen", "alpha_quant", "ker_wt",
                                         "w_obs", "lam_inv", "one_m_alpha_mean", "alpha_mean", "one_m_alpha_cov", "alpha_cov",
                                         "sigma_clip", "inv_nm", "m2_inv_nm", "beta_del", "sigma_random")] + \
               [(n, FP) for n in ("P", "Pd", "Pdd", "Gx", "Gy", "Kx", "Ky", "Wfit", "z_init", "theta0", "zb_iter")]


class OProjOut(C.Structure):
    _fields_ = [("cx", C.c_float * NV), ("cy", C.c_float * NV),
                ("xd", C.c_float * T), ("yd", C.c_float * T), ("xdd", C.c_float * T), ("ydd", C.c_float * T), ("y", C.c_float * T),
                ("res_norm", C.c_float), ("acc", C.c_float * T), ("steer", C.c_float * T), ("cost_base", C.c_float)]


class ORiskOut(C.Structure):
    _fields_ = [("risk", C.c_float), ("lane", C.c_float), ("beta", C.c_float * 64), ("sigma", C.c_float),
                ("res_beta", C.c_float * 64), ("red_cost", C.c_float * 64), ("red_idx", C.c_int * 64)]


class OSolveOut(C.Structure):
    _fields_ = [("cx", C.c_float * NV), ("cy", C.c_float * NV), ("cost_lane", C.c_float), ("cost_obs", C.c_float),
                ("beta", C.c_float * 64), ("sigma", C.c_float), ("res_beta", C.c_float * 64), ("sel_last", C.c_int32)]


class ONoise(C.Structure):
    _fields_ = [("z1", FP), ("z2", FP), ("z3", FP), ("k1", C.c_uint32 * 2), ("k2", C.c_uint32 * 2), ("b1", FP), ("b2", FP)]


class OSelectOut(C.Structure):
    _fields_ = [("sel", C.c_int), ("cost_min", C.c_float), ("top", C.c_int * 32), ("elite", C.c_int * 8)]


class OTrace(C.Structure):
    _fields_ = [("params", FP), ("res_norm", FP), ("risk", FP), ("lane", FP), ("cost_base", FP), ("mean", FP), ("cov", FP),
                ("sel", C.POINTER(C.c_int32)), ("cxy", FP)]


def _fp(a):
    return a.ctypes.data_as(FP)


def set_threads(n: int):
    """host threads oracle_solve splits the B samples of an iteration over (results are independent of n)"""
    lib().oracle_set_threads(int(n))


# ------------------------------------------------------------------------------------------------
# RNG (JAX 0.3.23 protocol restated in oracle_rng.h)

def prng_key(seed: int):
    return (0, int(seed) & 0xFFFFFFFF)


def split0(key):
    o = (C.c_uint32 * 4)()
    lib().oracle_rng_split(C.c_uint32(key[0]), C.c_uint32(key[1]), 2, o)
    return (int(o[0]), int(o[1]))


def split(key, m=2):
    o = (C.c_uint32 * (2 * m))()
    lib().oracle_rng_split(C.c_uint32(key[0]), C.c_uint32(key[1]), m, o)
    return np.array(list(o), dtype=np.uint32).reshape(m, 2)


def normal(key, n):
    out = np.empty(n, dtype=f32)
    lib().oracle_rng_normal(C.c_uint32(key[0]), C.c_uint32(key[1]), int(n), _fp(out))
    return out


def uniform(key, n, lo=0.0, hi=1.0):
    out = np.empty(n, dtype=f32)
    lib().oracle_rng_uniform(C.c_uint32(key[0]), C.c_uint32(key[1]), int(n), C.c_float(lo), C.c_float(hi), _fp(out))
    return out


def bits(key, n):
    out = np.empty(n, dtype=np.uint32)
    lib().oracle_rng_bits(C.c_uint32(key[0]), C.c_uint32(key[1]), int(n), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out


def beta(key, a, b):
    a = np.ascontiguousarray(a, dtype=f32).ravel(); b = np.ascontiguousarray(b, dtype=f32).ravel()
    out = np.empty(a.size, dtype=f32)
    lib().oracle_rng_beta(C.c_uint32(key[0]), C.c_uint32(key[1]), _fp(a), _fp(b), int(a.size), _fp(out))
    return out


def threefry(k0, k1, x0, x1):
    o = (C.c_uint32 * 2)()
    lib().oracle_threefry(C.c_uint32(k0), C.c_uint32(k1), C.c_uint32(x0), C.c_uint32(x1), o)
    return int(o[0]), int(o[1])


MATH_FN = {"exp": 0, "log": 1, "log1p": 2, "sin": 3, "cos": 4, "tan": 5, "atan": 6, "atan2": 7, "erfinv": 8, "lap": 12}


def math_vec(fn, x, y=None):
    x = np.ascontiguousarray(x, dtype=f32)
    y = x if y is None else np.ascontiguousarray(y, dtype=f32)
    out = np.empty_like(x)
    lib().oracle_math_vec(MATH_FN[fn], _fp(x), _fp(y), _fp(out), int(x.size))
    return out


# ------------------------------------------------------------------------------------------------
# Bernstein basis (S/bernstein_coeff_order10_arbitinterval.py:13-103), restated through the
# derivative identities  B'_{k,n} = n (B_{k-1,n-1} - B_{k,n-1}),  B''_{k,n} = n(n-1)(B_{k-2,n-2} - 2 B_{k-1,n-2} + B_{k,n-2}).

def _bern(n, k, t):
    if k < 0 or k > n:
        return np.zeros_like(t)
    return comb(n, k) * (1.0 - t) ** (n - k) * t ** k


def bernstein_basis(tmin, tmax, t_actual, n=10):
    t_actual = np.asarray(t_actual, dtype=np.float64).reshape(-1)
    l = float(tmax) - float(tmin)
    t = (t_actual - float(tmin)) / l
    P = np.stack([_bern(n, k, t) for k in range(n + 1)], axis=1)
    Pd = np.stack([n * (_bern(n - 1, k - 1, t) - _bern(n - 1, k, t)) for k in range(n + 1)], axis=1) / l
    Pdd = np.stack([n * (n - 1) * (_bern(n - 2, k - 2, t) - 2.0 * _bern(n - 2, k - 1, t) + _bern(n - 2, k, t))
                    for k in range(n + 1)], axis=1) / (l ** 2)
    return P, Pd, Pdd


def _ipow32(x, y):
    """lax.integer_pow for float32 arrays: binary exponentiation (what jnp `x**int` lowers to)."""
    x = x.astype(f32)
    if y == 0:
        return np.ones_like(x)
    acc = None
    while y > 0:
        if y & 1:
            acc = x if acc is None else (acc * x).astype(f32)
        y >>= 1
        if y > 0:
            x = (x * x).astype(f32)
    return acc


def bernstein_P_f32(num_prime, t_fin_prime, n=10):
    """P_prime of cem_helper.py:112-118: the basis evaluated in float32 through jnp [Q14]."""
    div = num_prime - 1
    step = (np.arange(div, dtype=f32) / f32(div)).astype(f32)
    stop = f32(t_fin_prime)
    tt = (f32(0.0) * (f32(1.0) - step) + stop * step).astype(f32)          # jnp.linspace, endpoint recovered exactly
    tt = np.concatenate([tt, np.array([stop], dtype=f32)])
    tmin, tmax = tt[0], tt[-1]
    l = f32(tmax - tmin)
    t = ((tt - tmin) / l).astype(f32)
    omt = (f32(1.0) - t).astype(f32)
    cols = [((f32(comb(n, k)) * _ipow32(omt, n - k)).astype(f32) * _ipow32(t, k)).astype(f32) for k in range(n + 1)]
    return np.stack(cols, axis=1).astype(f32)


# ------------------------------------------------------------------------------------------------
# constants of CEM.__init__ (cem.py:16-199) and the folded solves [D2]

VARIANTS = {"static": dict(y_lb=-2.25, y_ub=2.25, K_steer=0.01),      # S/optimizer/cem.py:155, cem_helper.py:24
            "dynamic": dict(y_lb=-2.25, y_ub=-1.25, K_steer=0.05)}    # D/optimizer/cem.py:155, cem_helper.py:24


class OracleCEM:
    """Constants + ctypes config for one (num_reduced, num_obs, noise_level, num_prime, noise, ...) setting."""

    def __init__(self, num_reduced, num_obs, noise_level, num_prime, noise, acc_const_noise, steer_const_noise,
                 variant="static", num_batch=100, maxiter_cem=20, num_samples_cem=100, maxiter_beta_cem=20, naive=False):
        v = VARIANTS[variant]
        self.variant = variant
        self.num_reduced, self.num_obs, self.num_prime, self.noise = num_reduced, num_obs, num_prime, noise
        self.num_mother = num_reduced ** 2
        self.num_batch, self.maxiter_cem = num_batch, maxiter_cem
        self.num, self.nvar, self.t_fin = T, NV, 15
        self.t = self.t_fin / self.num
        self.y_lb, self.y_ub, self.K_steer = v["y_lb"], v["y_ub"], v["K_steer"]
        self.a_obs, self.b_obs, self.wheel_base, self.ker_wt = 4.25, 2.75, 2.5, 1000.0
        self.beta_a, self.beta_b = 2, 5
        self.acc_const_noise, self.steer_const_noise = acc_const_noise, steer_const_noise
        self.tot_time = np.linspace(0, self.t_fin, self.num)
        P, Pd, Pdd = bernstein_basis(self.tot_time[0], self.tot_time[-1], self.tot_time)   # cem.py:42-46 (float64)
        self.P64, self.Pd64, self.Pdd64 = P, Pd, Pdd
        self.P, self.Pd, self.Pdd = (np.ascontiguousarray(a.astype(f32)) for a in (P, Pd, Pdd))   # cem.py:48
        Pf, Pdf, Pddf = (a.astype(np.float64) for a in (self.P, self.Pd, self.Pdd))
        A_eq_x = np.vstack((Pf[0], Pdf[0], Pddf[0]))                                           # cem.py:55
        A_eq_y = np.vstack((Pf[0], Pdf[0], Pddf[0], Pdf[-1]))                                  # cem.py:56
        # x_guess (cem_helper.py:169-230): k_p_v = k_p = 2, weight_smoothness = 100, rho = 1, 4 quarters of 25 rows
        Qx = 100.0 * Pddf.T @ Pddf
        Qy = 100.0 * Pddf.T @ Pddf
        gx, gy = [], []
        for q in range(4):
            sl = slice(25 * q, 25 * q + 25)
            A_vd = Pddf[sl] - 2.0 * Pdf[sl]
            A_pd = Pddf[sl] - 2.0 * Pf[sl]
            Qx = Qx + A_vd.T @ A_vd
            Qy = Qy + A_pd.T @ A_pd
            gx.append(-2.0 * A_vd.T @ np.ones(25))      # -lincost_x = sum_q v_q * gx[q]
            gy.append(-2.0 * A_pd.T @ np.ones(25))
        Mx = np.linalg.inv(np.block([[Qx, A_eq_x.T], [A_eq_x, np.zeros((3, 3))]]))
        My = np.linalg.inv(np.block([[Qy, A_eq_y.T], [A_eq_y, np.zeros((4, 4))]]))
        self.Gx = np.ascontiguousarray(np.hstack((Mx[:NV, :NV] @ np.stack(gx, 1), Mx[:NV, NV:])).astype(f32))   # (11,7)
        self.Gy = np.ascontiguousarray(np.hstack((My[:NV, :NV] @ np.stack(gy, 1), My[:NV, NV:])).astype(f32))   # (11,8)
        # projection KKT (projection.py:145-168): rho's = 1, A_projection = I, A_lane_bound = [P[1:]; -P[1:]]
        A_lane = np.vstack((Pf[1:], -Pf[1:]))
        cost_x = np.eye(NV) + Pddf.T @ Pddf + Pdf.T @ Pdf
        cost_y = cost_x + A_lane.T @ A_lane
        Kx = np.linalg.inv(np.block([[cost_x, A_eq_x.T], [A_eq_x, np.zeros((3, 3))]]))
        Ky = np.linalg.inv(np.block([[cost_y, A_eq_y.T], [A_eq_y, np.zeros((4, 4))]]))
        self.Kx = np.ascontiguousarray(Kx[:NV, :].astype(f32))    # (11,14)
        self.Ky = np.ascontiguousarray(Ky[:NV, :].astype(f32))    # (11,15)
        # ridge fit (cem_helper.py:553-564) on the float32 P_prime of cem_helper.py:112-118
        self.P_prime = bernstein_P_f32(num_prime, num_prime * self.t)
        Pp = self.P_prime.astype(np.float64)
        self.Wfit = np.ascontiguousarray((np.linalg.inv(Pp.T @ Pp + 0.05 * np.eye(NV)) @ Pp.T).astype(f32))   # (11,np)
        # constant normal tables [Q8]
        key0 = prng_key(0)
        key_init = split0(key0)                                                  # cem_helper.py:86,125
        self.z_init = normal(key_init, num_batch * NPAR).reshape(num_batch, NPAR)
        d = self.num_mother + 1
        S, ne = num_samples_cem, max(int(0.1 * num_samples_cem) + 1, 3)          # compute_beta.py:14,26
        z0 = normal(split0(key_init), S * d).reshape(S, d)                       # compute_beta.py:108,44-46
        th0 = (np.sqrt(f32(20.0)).astype(f32) * z0).astype(f32)                  # chol(20 I) = sqrt(20) I
        th0[:, -1] = np.maximum(th0[:, -1], f32(0.01))                           # compute_beta.py:47
        self.theta0 = np.ascontiguousarray(th0)
        zb = np.empty((maxiter_beta_cem, S - ne, d), dtype=f32)
        carry = key_init
        for it in range(maxiter_beta_cem):
            carry = split0(carry)                                                # compute_beta.py:131
            zb[it] = normal(split0(carry), (S - ne) * d).reshape(S - ne, d)      # compute_beta.py:54,63
        self.zb_iter = zb
        self.S_in, self.n_el_in, self.iters_in = S, ne, maxiter_beta_cem

        c = OCfg()
        c.B, c.np, c.nr, c.nm, c.O = num_batch, num_prime, num_reduced, self.num_mother, num_obs
        c.iters, c.n_el, c.n_el_cost = maxiter_cem, 5, min(20, num_batch)
        c.noise_kind = NOISE_KINDS[noise]
        c.S_in, c.iters_in, c.n_el_in, c.naive = S, maxiter_beta_cem, ne, int(naive)
        c.sigma_acc = c.sigma_steer = noise_level
        c.ksig_steer = self.K_steer * noise_level           # python-double product, then float32 (cem_helper.py:436)
        c.acc_const, c.steer_const = acc_const_noise, steer_const_noise
        c.beta_a, c.beta_b = 2.0, 5.0
        c.v_min, c.v_max, c.a_max = 0.1, 30.0, 18.0
        c.b_lane_ub, c.b_lane_lb = 1.0 * self.y_ub, -1.0 * self.y_lb       # projection.py:127-128, gamma = 1
        c.y_lb, c.y_ub = self.y_lb, self.y_ub
        c.a2_obs, c.b2_obs = self.a_obs ** 2, self.b_obs ** 2
        c.wheel_base, c.dt, c.steer_max, c.steer_rate_pen = 2.5, self.t, 0.6, 0.05
        c.alpha_quant, c.ker_wt = 0.98, 1000.0
        c.lam_inv = 1 / 0.9
        c.one_m_alpha_mean, c.alpha_mean, c.one_m_alpha_cov, c.alpha_cov = 1 - 0.6, 0.6, 1 - 0.6, 0.6
        c.sigma_clip = 0.01
        c.inv_nm = 1 / self.num_mother                    # compute_beta.py:77
        c.m2_inv_nm = -2 * (1 / self.num_mother)          # compute_beta.py:86
        c.beta_del = 1 / num_reduced                      # kernel_computation.py:72 / cem.py:355
        c.sigma_random = 0.01
        c.P, c.Pd, c.Pdd = _fp(self.P), _fp(self.Pd), _fp(self.Pdd)
        c.Gx, c.Gy, c.Kx, c.Ky, c.Wfit = _fp(self.Gx), _fp(self.Gy), _fp(self.Kx), _fp(self.Ky), _fp(self.Wfit)
        c.z_init, c.theta0, c.zb_iter = _fp(self.z_init), _fp(self.theta0), _fp(self.zb_iter)
        self.cfg = c

    # weights on the obstacle risk inside compute_cost (cem.py:161-163 with :294, 423, 549, 675)
    W_OBS = {"mmd_opt": 1e3, "mmd_random": 1e3, "cvar": 1e3, "saa": 1e6}

    def _cfg_for(self, cost):
        self.cfg.w_obs = self.W_OBS[cost]
        return C.byref(self.cfg)

    def compute_obs_trajectories(self, x_obs, y_obs, vx_obs, vy_obs, psi_obs):
        """cem_helper.py:366-378 (float32 like jnp)."""
        tt = self.tot_time.astype(f32)[:, None]
        x = (np.asarray(x_obs, f32) + np.asarray(vx_obs, f32) * tt).T
        y = (np.asarray(y_obs, f32) + np.asarray(vy_obs, f32) * tt).T
        psi = np.tile(np.asarray(psi_obs, f32), (self.num, 1)).T
        return np.ascontiguousarray(x, f32), np.ascontiguousarray(y, f32), np.ascontiguousarray(psi, f32)

    def solve(self, cost, idx_mpc, init_state, mean, cov, x_obs_traj, y_obs_traj, v_des, trace=False):
        init_state = np.ascontiguousarray(init_state, f32); mean = np.ascontiguousarray(mean, f32)
        cov = np.ascontiguousarray(cov, f32); xo = np.ascontiguousarray(x_obs_traj, f32); yo = np.ascontiguousarray(y_obs_traj, f32)
        assert xo.shape == (self.num_obs, T) and yo.shape == (self.num_obs, T)
        out = OSolveOut()
        tr, trp = None, None
        if trace:
            I, B = self.maxiter_cem, self.num_batch
            tr = dict(params=np.zeros((I + 1, B, NPAR), f32), res_norm=np.zeros((I, B), f32), risk=np.zeros((I, B), f32),
                      lane=np.zeros((I, B), f32), cost_base=np.zeros((I, B), f32), mean=np.zeros((I + 1, NPAR), f32),
                      cov=np.zeros((I + 1, NPAR * NPAR), f32), sel=np.zeros(I, np.int32), cxy=np.zeros((I, 2 * NV), f32))
            ot = OTrace(*[(_fp(tr[k]) if k != "sel" else tr[k].ctypes.data_as(C.POINTER(C.c_int32)))
                          for k in ("params", "res_norm", "risk", "lane", "cost_base", "mean", "cov", "sel", "cxy")])
            trp = C.byref(ot)
        lib().oracle_solve(self._cfg_for(cost), COST_KINDS[cost], C.c_int32(int(idx_mpc)), _fp(init_state), _fp(mean), _fp(cov),
                           _fp(xo), _fp(yo), C.c_float(v_des), C.byref(out), trp)
        nr = self.num_reduced
        res = dict(cx=np.array(out.cx, f32), cy=np.array(out.cy, f32), cost_lane=f32(out.cost_lane), cost_obs=f32(out.cost_obs),
                   beta=np.array(out.beta[:nr], f32), sigma=f32(out.sigma), res_beta=np.array(out.res_beta[:self.iters_in], f32),
                   sel=int(out.sel_last))
        if trace:
            res["trace"] = tr
        return res

    # ---- stage entry points (teacher-forced tests) -------------------------------------------
    def project(self, param, beq_x, beq_y, v_des, lam_x, lam_y, s_lane):
        """one sample; lam_x, lam_y (11,), s_lane (198,) are updated in place. Returns dict of outputs."""
        o = OProjOut()
        param = np.ascontiguousarray(param, f32); bx = np.ascontiguousarray(beq_x, f32); by = np.ascontiguousarray(beq_y, f32)
        lib().oracle_project(C.byref(self.cfg), _fp(param), _fp(bx), _fp(by), C.c_float(v_des), _fp(lam_x), _fp(lam_y), _fp(s_lane), C.byref(o))
        return {k: (np.array(getattr(o, k), f32) if k not in ("res_norm", "cost_base") else f32(getattr(o, k)))
                for k in ("cx", "cy", "xd", "yd", "xdd", "ydd", "y", "res_norm", "acc", "steer", "cost_base")}

    def noise_tables(self, idx_mpc, it):
        n = self.num_reduced * self.num_prime
        z1, z2, z3 = (np.empty(n, f32) for _ in range(3))
        zc = np.empty(((self.num_batch - 5) * NPAR), f32)
        keys = (C.c_uint32 * 4)()
        lib().oracle_noise_tables(C.byref(self.cfg), C.c_int32(int(idx_mpc)), C.c_int32(int(it)), _fp(z1), _fp(z2), _fp(z3), _fp(zc), keys)
        return z1, z2, z3, zc.reshape(-1, NPAR), [int(k) for k in keys]

    def risk(self, cost, acc, steer, st0, noise, x_obs_traj, y_obs_traj, want_rollouts=False, beta_draws=None):
        """`beta_draws` = (b_acc, b_steer), each (nr, np): the two `jax.random.beta` samples of cem_helper.py:427-436 injected instead of drawn"""
        z1, z2, z3, _, keys = noise
        b1 = b2 = None
        if beta_draws is not None:
            b1, b2 = (np.ascontiguousarray(b, f32).reshape(-1) for b in beta_draws)
        nz = ONoise(_fp(z1), _fp(z2), _fp(z3), (C.c_uint32 * 2)(keys[0], keys[1]), (C.c_uint32 * 2)(keys[2], keys[3]),
                    _fp(b1) if b1 is not None else None, _fp(b2) if b2 is not None else None)
        o = ORiskOut()
        acc = np.ascontiguousarray(acc, f32); steer = np.ascontiguousarray(steer, f32); st0 = np.ascontiguousarray(st0, f32)
        xo = np.ascontiguousarray(x_obs_traj, f32); yo = np.ascontiguousarray(y_obs_traj, f32)
        R = self.num_mother if cost == "mmd_opt" else self.num_reduced
        xr = np.zeros((R, self.num_prime), f32); yr = np.zeros((R, self.num_prime), f32)
        lib().oracle_risk(self._cfg_for(cost), COST_KINDS[cost], _fp(acc), _fp(steer), _fp(st0), C.byref(nz), _fp(xo), _fp(yo), C.byref(o),
                          _fp(xr) if want_rollouts else None, _fp(yr) if want_rollouts else None)
        nr = self.num_reduced
        d = dict(risk=f32(o.risk), lane=f32(o.lane), beta=np.array(o.beta[:nr], f32), sigma=f32(o.sigma),
                 res_beta=np.array(o.res_beta[:self.iters_in], f32), red_cost=np.array(o.red_cost[:nr], f32),
                 red_idx=np.array(o.red_idx[:nr], np.int32))
        if want_rollouts:
            d["x_roll"], d["y_roll"] = xr, yr
        return d

    def sample_params(self, mean, cov, z):
        z = np.ascontiguousarray(z, f32); out = np.empty_like(z)
        mean = np.ascontiguousarray(mean, f32); cov = np.ascontiguousarray(cov, f32)
        lib().oracle_sample_params(C.byref(self.cfg), _fp(mean), _fp(cov), _fp(z), int(z.shape[0]), _fp(out))
        return out

    def select(self, cost, res_norm, risk, cost_base, params, mean, cov, z_cem):
        """returns (params_next, mean_new, cov_new, info)"""
        res_norm, risk, cost_base, params, z_cem = (np.ascontiguousarray(a, f32) for a in (res_norm, risk, cost_base, params, z_cem))
        mean = np.array(mean, f32).copy(); cov = np.array(cov, f32).reshape(-1).copy()
        nxt = np.empty_like(params); o = OSelectOut()
        lib().oracle_select(self._cfg_for(cost), _fp(res_norm), _fp(risk), _fp(cost_base), _fp(params), _fp(nxt), _fp(mean), _fp(cov), _fp(z_cem), C.byref(o))
        info = dict(sel=int(o.sel), cost_min=f32(o.cost_min), top=np.array(o.top[:self.cfg.n_el_cost], np.int32), elite=np.array(o.elite[:5], np.int32))
        return nxt, mean, cov.reshape(NPAR, NPAR), info


# ------------------------------------------------------------------------------------------------
# scene generators and driver inputs

def static_scene(num_obs, seed):
    """compute_obs_data (S/main_mpc.py:10-21), legacy np.random stream."""
    np.random.seed(seed)
    x = np.random.choice(np.array([35, 40, 45, 50, 55, 60, 65, 70, 75]), (num_obs,), replace=False)
    y = np.random.choice(np.array([-1.75, 1.75]), (num_obs,))
    z = np.zeros(num_obs)
    return x, y, z.copy(), z.copy(), z.copy()


def driver_inputs(variant="static"):
    """init_state / mean / cov / v_des of S/main_mpc.py:46-74 (D/main_mpc.py:34-62 has y_init = -1.75)."""
    y0 = 1.75 if variant == "static" else -1.75
    init_state = np.array([0.0, y0, 5.0, 0.0, 0.0, 0.0], f32)
    mean = np.array([15.0] * 4 + [0.0] * 4, f32)
    cov = np.diag(np.array([20.0] * 4 + [100.0] * 4)).astype(f32)
    return init_state, mean, cov, 15.0


def static_episode(num_obs, k):
    """scene k and its idx_mpc exactly as the loop body of S/main_mpc.py:106-119 draws them."""
    xo, yo, vx, vy, psi = static_scene(num_obs, k)
    idx_mpc = int(np.random.randint(1, 10000))
    return (xo, yo, vx, vy, psi), idx_mpc
