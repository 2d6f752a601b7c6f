"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, BIT EXACT.

The arithmetic contract makes every float32 result reproducible, so the bar
here is `np.array_equal`, which is stricter than the 1e-4 relative tolerance north_star asks for
(and implies identical elite index sets and identical downstream collision counts)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


def _eq(a, b, what=""):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if not np.array_equal(a, b, equal_nan=True):
        bad = np.flatnonzero(~((a == b) | (np.isnan(a) & np.isnan(b))).ravel())
        raise AssertionError(f"{what}: {bad.size}/{a.size} elements differ; first at {bad[:5]}: got {a.ravel()[bad[:5]]} want {b.ravel()[bad[:5]]}")


@pytest.fixture(scope="module")
def mods(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpcmmd_b200 import cem_impl
    from oracle import oracle as O
    return cem_impl, O


MATH_RANGES = {"exp": (-100.0, 90.0), "log": (1e-42, 1e6), "log1p": (-0.9999999, 10.0), "sin": (-700.0, 700.0), "cos": (-700.0, 700.0),
               "tan": (-1.6, 1.6), "atan": (-1e4, 1e4), "erfinv": (-0.99999994, 0.99999994)}


def test_ieee_shortcuts_exhaustive(mods):
    """dm::sqrt_rcp (one range check for the pivot's square root and its reciprocal) and dm::div10 (covariance / (num_elite - 1)) are the IEEE operations of
    the contract: compared on the device with sqrtf, 1.0f / sqrtf and x / 10.0f over ALL 2^32 float bit patterns (NaN == NaN), zero mismatches."""
    cem_impl, _ = mods
    assert cem_impl.selfcheck_ieee(0) == (0, 0, 0)


def test_device_laplace_kernel_entry_bit_exact(mods):
    """dm::lap_ (csrc/dmath.cuh) == om_lap (oracle/oracle_math.h): the Laplace kernel entry of the reduced-set inner CEM, including capped, zero, infinite and NaN arguments"""
    cem_impl, O = mods
    rng = np.random.default_rng(5)
    d = np.concatenate([rng.uniform(0, 60, 400000), rng.uniform(0, 1e4, 100000), [0.0, -0.0, np.inf, np.nan, 1.0, 1.0, 3.0]]).astype(f32)
    sig = np.concatenate([rng.uniform(0.01, 4.0, 500000), [1.0, 1.0, 0.01, 1.0, np.nan, np.inf, 0.0]]).astype(f32)
    _eq(cem_impl.math_vec(12, d, sig), O.math_vec("lap", d, sig), "lap")


@pytest.mark.parametrize("fn", list(MATH_RANGES))
def test_device_math_bit_exact(mods, fn):
    cem_impl, O = mods
    rng = np.random.default_rng(1)
    lo, hi = MATH_RANGES[fn]
    x = rng.uniform(lo, hi, 400000).astype(f32)
    if fn == "log":
        x = np.concatenate([x, np.exp(rng.uniform(-95, 20, 200000)).astype(f32), [0.0, np.inf]]).astype(f32)
    if fn in ("sin", "cos"):
        x = np.concatenate([x, rng.uniform(-7, 7, 400000).astype(f32)])
    x = np.concatenate([x, np.array([0.0, -0.0, np.nan], f32)])
    _eq(cem_impl.math_vec(O.MATH_FN[fn], x), O.math_vec(fn, x), fn)
    if fn == "exp":      # the hot-loop variant on its domain
        xn = x[(x <= 0) | np.isnan(x)]
        _eq(cem_impl.math_vec(9, xn), O.math_vec("exp", xn), "exp_nonpos")
    if fn == "sin":
        _eq(cem_impl.math_vec(10, x), O.math_vec("sin", x), "sincos.s")
    if fn == "cos":
        _eq(cem_impl.math_vec(11, x), O.math_vec("cos", x), "sincos.c")


def test_device_atan2_bit_exact(mods):
    cem_impl, O = mods
    rng = np.random.default_rng(2)
    x = rng.normal(0, 10, 300000).astype(f32); y = rng.normal(0, 10, 300000).astype(f32)
    x[:100] = 0.0; y[50:150] = 0.0
    _eq(cem_impl.math_vec(7, x, y), O.math_vec("atan2", x, y), "atan2")


def test_device_rng_normal_matches_known_answers(mods):
    cem_impl, O = mods
    got = cem_impl.rng_normal((0, 0), 1)
    assert abs(float(got[0]) - (-0.20584226)) < 2e-7          # jax.random.normal(PRNGKey(0), (1,)) from the JAX docs
    for key, n in (((0, 0), 1), ((0, 42), 7), ((123, 456), 800), ((4146024105, 967050713), 2501)):
        _eq(cem_impl.rng_normal(key, n), O.normal(key, n), f"normal{key}")


def test_device_rng_beta_bit_exact(mods):
    cem_impl, O = mods
    rng = np.random.default_rng(3)
    a = np.abs(rng.normal(0, 2, 3000)).astype(f32) * 2; b = a * f32(2.5)
    a[:5] = 0.0; b[:5] = 0.0                                   # alpha = 0 edge (acc[99] = 0 in the reference)
    a[5:10] = 1e-6; b[5:10] = 2.5e-6
    got = cem_impl.rng_beta((7, 9), a, b); want = O.beta((7, 9), a, b)
    _eq(got, want, "beta")
    ok = ~np.isnan(want)
    assert ok.sum() > 2900 and np.all((want[ok] >= 0) & (want[ok] <= 1))


SMALL = dict(num_batch=24, maxiter_cem=3, num_samples_cem=40, maxiter_beta_cem=4)


def _pair(mods, args, variant="static", **kw):
    cem_impl, O = mods
    return cem_impl.CEM(*args, variant=variant, max_episodes=kw.pop("max_episodes", 4), **kw), O.OracleCEM(*args, variant=variant, **kw)


def test_constant_tables_bit_exact(mods):
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0))
    zi, th, zb = prob.tables()
    _eq(zi, ora.z_init, "z_init"); _eq(th, ora.theta0, "theta0"); _eq(zb, ora.zb_iter, "zb_iter")


@pytest.mark.parametrize("variant", ["static", "dynamic"])
def test_stage_project_bit_exact(mods, variant):
    cem_impl, O = mods
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), variant=variant)
    rng = np.random.default_rng(5)
    n = 37
    params = np.concatenate([rng.uniform(0.1, 30, (n, 4)), rng.normal(0, 6, (n, 4))], 1).astype(f32)
    beq_x = np.array([0.0, 5.0, 0.3], f32); beq_y = np.array([1.75 if variant == "static" else -1.75, 0.2, -0.1, 0.0], f32)
    lam_x = rng.normal(0, 0.5, (n, 11)).astype(f32); lam_y = rng.normal(0, 0.5, (n, 11)).astype(f32)
    s_lane = np.abs(rng.normal(0, 1, (n, 198))).astype(f32)
    lam_x[:10] = 0; lam_y[:10] = 0; s_lane[:10] = 0
    got = prob.stage_project(params, beq_x, beq_y, 15.0, lam_x, lam_y, s_lane)
    for i in range(n):
        lx, ly, sl = lam_x[i].copy(), lam_y[i].copy(), s_lane[i].copy()
        ref = ora.project(params[i], beq_x, beq_y, 15.0, lx, ly, sl)
        for k in ("cx", "cy", "res_norm", "acc", "steer", "cost_base"):
            _eq(got[k][i], ref[k], f"project[{i}].{k}")
        _eq(got["lam_x"][i], lx, "lam_x"); _eq(got["lam_y"][i], ly, "lam_y"); _eq(got["s_lane"][i], sl, "s_lane")


def _norm_err(got, ref):
    ref = np.asarray(ref, np.float64); got = np.asarray(got, np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


@pytest.mark.parametrize("variant,hard", [("static", False), ("static", True), ("dynamic", True)])
def test_stage_project_tensor_core_within_tolerance(mods, monkeypatch, variant, hard):
    """k_project_tc (MPCMMD_PROJ=tc: tcgen05 kind::tf32 products on an error-free hi/lo split, TMEM accumulators) against the oracle.
    Not bit exact by construction; the bar is north_star's 1e-4 relative (max error over the array / max magnitude of the array) on
    trajectories (cx, cy), controls, costs, multipliers and slacks.  res_norm of a feasible sample is pure round-off (~1e-5 in the
    reference's own float32), hence the absolute floor.  300 samples = two full 128-sample tiles and a ragged one."""
    cem_impl, O = mods
    monkeypatch.setenv("MPCMMD_PROJ", "tc-always")
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), variant=variant)
    rng = np.random.default_rng(11)
    n = 300
    params = np.concatenate([rng.uniform(0.1, 45 if hard else 30, (n, 4)), rng.normal(0, 12 if hard else 6, (n, 4))], 1).astype(f32)
    beq_x = np.array([0.0, 5.0, 0.3], f32); beq_y = np.array([1.75 if variant == "static" else -1.75, 0.2, -0.1, 0.0], f32)
    lam_x = rng.normal(0, 0.5, (n, 11)).astype(f32); lam_y = rng.normal(0, 0.5, (n, 11)).astype(f32)
    s_lane = np.abs(rng.normal(0, 1, (n, 198))).astype(f32)
    lam_x[:10] = 0; lam_y[:10] = 0; s_lane[:10] = 0
    got = prob.stage_project(params, beq_x, beq_y, 15.0, lam_x, lam_y, s_lane)
    ref = {k: [] for k in got}
    for i in range(n):
        lx, ly, sl = lam_x[i].copy(), lam_y[i].copy(), s_lane[i].copy()
        r = ora.project(params[i], beq_x, beq_y, 15.0, lx, ly, sl)
        for k in ("cx", "cy", "res_norm", "acc", "steer", "cost_base"):
            ref[k].append(np.asarray(r[k]))
        ref["lam_x"].append(lx); ref["lam_y"].append(ly); ref["s_lane"].append(sl)
    for k in ("cx", "cy", "acc", "steer", "cost_base", "lam_x", "lam_y", "s_lane"):
        assert not np.isnan(got[k]).any(), k
        assert _norm_err(got[k], ref[k]) <= 1e-4, (k, _norm_err(got[k], ref[k]))
    rn = np.asarray(ref["res_norm"], np.float64)
    err = np.abs(got["res_norm"] - rn)
    assert err.max() <= 1e-4 * np.abs(rn).max() + 5e-5, (float(err.max()), int(err.argmax()), float(rn[err.argmax()]))


@pytest.mark.parametrize("n", [1, 37, 129])
def test_stage_project_tensor_core_ragged_tiles(mods, monkeypatch, n):
    """k_project_tc with fewer samples than a 128-sample tile, and one sample past a full tile: idle TMEM lanes shadow the last
    sample and store nothing; every live row still meets the 1e-4 bar."""
    cem_impl, O = mods
    monkeypatch.setenv("MPCMMD_PROJ", "tc-always")
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0))
    rng = np.random.default_rng(100 + n)
    params = np.concatenate([rng.uniform(0.1, 35, (n, 4)), rng.normal(0, 8, (n, 4))], 1).astype(f32)
    beq_x = np.array([0.0, 5.0, 0.0], f32); beq_y = np.array([1.75, 0.0, 0.0, 0.0], f32)
    lam_x = rng.normal(0, 0.5, (n, 11)).astype(f32); lam_y = rng.normal(0, 0.5, (n, 11)).astype(f32)
    s_lane = np.abs(rng.normal(0, 1, (n, 198))).astype(f32)
    got = prob.stage_project(params, beq_x, beq_y, 15.0, lam_x, lam_y, s_lane)
    for i in range(n):
        lx, ly, sl = lam_x[i].copy(), lam_y[i].copy(), s_lane[i].copy()
        r = ora.project(params[i], beq_x, beq_y, 15.0, lx, ly, sl)
        for k, rv in (("cx", r["cx"]), ("cy", r["cy"]), ("acc", r["acc"]), ("steer", r["steer"]), ("lam_x", lx), ("lam_y", ly), ("s_lane", sl)):
            assert _norm_err(got[k][i], rv) <= 1e-4, (i, k, _norm_err(got[k][i], rv))
        assert abs(float(got["cost_base"][i]) - float(r["cost_base"])) <= 1e-4 * abs(float(r["cost_base"])) + 5e-5


def test_solve_with_tensor_core_projection(mods, monkeypatch):
    """MPCMMD_PROJ=tc over the whole 200-episode sweep of configs[1] (cvar + mmd_opt), against the exact projection on the same episodes.

    What can and cannot be compared: the reference's first elite stage keeps the 20 smallest projection residuals (cem.py:233), and for feasible
    samples those residuals are pure float32 round-off (~5e-6), so ANY change of rounding -- XLA vs NumPy vs this tf32-split kernel -- keeps other
    rows and the CEM iterates of one episode diverge.  Per-episode trajectories are therefore not comparable beyond the
    boundary conditions; what a user of the sweep consumes is comparable: WHICH episodes are accepted (main_mpc.py:121-128) and the distribution
    of the planned motion.  This test pins, at sweep level: boundary coefficients to 1e-5, accepted sets (Jaccard >= 0.9, counts within 4 %),
    and the 10 / 50 / 90 % quantiles of the distance covered and of the final lateral position."""
    cem_impl, O = mods
    args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
    ora = O.OracleCEM(*args, variant="static")
    init_state, mean, cov, v_des = O.driver_inputs("static")
    E = 200
    idx, xo, yo = _episodes(O, ora, E, 4)
    res = {}
    for tag in ("exact", "tc"):
        if tag == "tc":
            monkeypatch.setenv("MPCMMD_PROJ", "tc-always")
        else:
            monkeypatch.delenv("MPCMMD_PROJ", raising=False)
        prob = cem_impl.CEM(*args, variant="static", max_episodes=E)
        res[tag] = {c: prob.solve_batch(c, idx, np.stack([init_state] * E), np.stack([mean] * E), np.stack([cov] * E), xo, yo, [v_des] * E)
                    for c in ("cvar", "mmd_opt")}
        del prob
    P_end = np.ones(11); P_end[:-1] = 0.0                                    # Bernstein basis at t_fin: the last coefficient is the end point
    for c, thr in (("cvar", 1e-5), ("mmd_opt", -999.0)):                    # main_mpc.py:88-97
        t, x = res["tc"][c], res["exact"][c]
        for k in ("cx", "cy", "cost_obs", "cost_lane"):
            assert np.isfinite(t[k]).all(), (c, k)
        np.testing.assert_allclose(t["cx"][:, :3], x["cx"][:, :3], rtol=1e-5, atol=1e-5)      # x0, v0, a0 boundary conditions
        np.testing.assert_allclose(t["cy"][:, :3], x["cy"][:, :3], rtol=1e-5, atol=1e-5)
        at, ax = t["cost_obs"] <= thr, x["cost_obs"] <= thr
        jac = (at & ax).sum() / max((at | ax).sum(), 1)
        print(f"{c}: accepted exact {ax.sum()} tc {at.sum()} Jaccard {jac:.3f}")
        assert jac >= 0.9 and abs(int(at.sum()) - int(ax.sum())) <= 0.04 * E, (c, jac, at.sum(), ax.sum())
        qt, qx = np.quantile(t["cx"][:, -1], [0.1, 0.5, 0.9]), np.quantile(x["cx"][:, -1], [0.1, 0.5, 0.9])
        print(f"{c}: distance covered quantiles exact {qx} tc {qt}")
        assert np.abs(qt - qx).max() <= 0.03 * np.abs(x["cx"][:, -1]).max(), (c, qt, qx)
        # the final lateral position is bimodal (the plan ends in the left or in the right lane), so compare the split and the spread, not the median
        lt, lx = float((t["cy"][:, -1] > 0).mean()), float((x["cy"][:, -1] > 0).mean())
        qt, qx = np.quantile(np.abs(t["cy"][:, -1]), [0.1, 0.5, 0.9]), np.quantile(np.abs(x["cy"][:, -1]), [0.1, 0.5, 0.9])
        print(f"{c}: ends left of the centre line exact {lx:.3f} tc {lt:.3f}; |y_end| quantiles exact {qx} tc {qt}")
        assert abs(lt - lx) <= 0.08 and np.abs(qt - qx).max() <= 0.25, (c, lt, lx, qt, qx)


def test_stage_noise_bit_exact(mods):
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0))
    for idx_mpc, it in ((6745, 0), (1, 19), (9999, 7)):
        g = prob.stage_noise(idx_mpc, it); r = ora.noise_tables(idx_mpc, it)
        for a, b, nm in zip(g[:4], r[:4], ("z1", "z2", "z3", "zcem")):
            _eq(a, b, nm)
        assert list(g[4]) == list(r[4])


def _controls(ora, rng, n):
    out = []
    beq_x = np.array([0.0, 5.0, 0.0], f32); beq_y = np.array([1.75, 0.0, 0.0, 0.0], f32)
    for _ in range(n):
        p = np.concatenate([rng.uniform(2, 25, 4), rng.normal(0, 3, 4)]).astype(f32)
        o = ora.project(p, beq_x, beq_y, 15.0, np.zeros(11, f32), np.zeros(11, f32), np.zeros(198, f32))
        out.append((o["acc"], o["steer"]))
    return np.stack([a for a, _ in out]), np.stack([s for _, s in out])


@pytest.mark.parametrize("cost,noise,nr,npr,nobs", [
    ("cvar", "gaussian", 5, 30, 2), ("cvar", "beta", 5, 50, 4), ("saa", "gaussian", 4, 20, 3), ("mmd_random", "gaussian", 5, 30, 2),
    ("mmd_opt", "gaussian", 5, 30, 2), ("mmd_opt", "beta", 3, 20, 2), ("mmd_opt", "gaussian", 10, 40, 3), ("cvar", "gaussian", 10, 100, 6),
    ("mmd_opt", "beta", 6, 30, 2), ("mmd_opt", "gaussian", 8, 25, 3), ("mmd_opt", "gaussian", 7, 20, 2), ("mmd_opt", "beta", 9, 20, 2)])
def test_stage_risk_bit_exact(mods, cost, noise, nr, npr, nobs):
    kw = dict(num_samples_cem=40, maxiter_beta_cem=4) if cost == "mmd_opt" else {}
    prob, ora = _pair(mods, (nr, nobs, 0.3 if noise == "beta" else 0.1, npr, noise, 0.05, 0.01), **kw)
    rng = np.random.default_rng(11)
    n = 6 if cost == "mmd_opt" else 12
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(4242, 3)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(nobs, 5)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    xo = xo.copy(); xo[0] = np.linspace(2, 60, 100)      # make sure some rollouts actually collide
    yo = yo.copy(); yo[0] = 1.75
    got = prob.stage_risk(cost, acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk(cost, acc[i], steer[i], st0, noise_t, xo, yo)
        _eq(got["risk"][i], ref["risk"], f"risk[{i}]"); _eq(got["lane"][i], ref["lane"], f"lane[{i}]")
        if cost in ("mmd_opt", "mmd_random"):
            _eq(got["beta"][i], ref["beta"], "beta"); _eq(got["sigma"][i], ref["sigma"], "sigma")
        if cost == "mmd_opt":
            _eq(got["res_beta"][i], ref["res_beta"], "res_beta")
    assert np.any(got["risk"] != got["risk"][0]) or cost == "saa"


@pytest.mark.parametrize("nr,npr,noise", [(6, 20, "gaussian"), (10, 30, "beta")])
def test_stage_risk_generic_kernel_reference_sizes(mods, nr, npr, noise):
    """num_reduced 6..10 with the reference's inner-CEM population (100 samples, 11 elites): the generic shared-memory kernel in its build with those sizes as
    compile-time constants (k_inner_cem<NR, 100, 11>); the other generic-kernel cases run reduced populations, i.e. the run-time-size build.  Bit exact vs the oracle."""
    cem_impl, O = mods
    prob, ora = _pair(mods, (nr, 2, 0.3 if noise == "beta" else 0.1, npr, noise, 0.05, 0.01), maxiter_beta_cem=3)
    rng = np.random.default_rng(23)
    n = 3
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(777, 2)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(2, 4)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    xo = xo.copy(); xo[0] = np.linspace(2, 60, 100)
    yo = yo.copy(); yo[0] = 1.75
    got = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk("mmd_opt", acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(got[k][i], ref[k], f"nr={nr} {k}[{i}]")


@pytest.mark.parametrize("cost,nr,npr", [("cvar", 5, 50), ("mmd_opt", 5, 30), ("mmd_random", 4, 20)])
def test_stage_risk_injected_beta_draws(mods, cost, nr, npr):
    """beta noise with the two jax.random.beta samples of cem_helper.py:427-436 INJECTED as tensors (mpcmmd_stage_risk_injected): no device RNG
    runs, CUDA and oracle consume the same draws and must agree bit for bit downstream of the sampler (rollouts, reduced-set CEM, risk).  This
    is the entry a machine with the reference's own jax==0.3.23 uses to take the restated Marsaglia-Tsang sampler out of the comparison."""
    kw = dict(num_samples_cem=40, maxiter_beta_cem=4) if cost == "mmd_opt" else {}
    prob, ora = _pair(mods, (nr, 3, 0.3, npr, "beta", 0.05, 0.01), **kw)
    rng = np.random.default_rng(29)
    n = 7
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(99, 1)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(3, 4)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    xo = xo.copy(); xo[0] = np.linspace(2, 60, 100); yo = yo.copy(); yo[0] = 1.75
    # any draws in (0, 1) do: here NumPy's Beta(2|u|, 5|u|) on the sample's own controls (|u| floored: NumPy rejects a = 0)
    b_acc = np.stack([rng.beta(2 * np.maximum(np.abs(acc[i][:npr]), 1e-3), 5 * np.maximum(np.abs(acc[i][:npr]), 1e-3), (nr, npr)) for i in range(n)]).astype(f32)
    b_steer = np.stack([rng.beta(2 * np.maximum(np.abs(steer[i][:npr]), 1e-3), 5 * np.maximum(np.abs(steer[i][:npr]), 1e-3), (nr, npr)) for i in range(n)]).astype(f32)
    got = prob.stage_risk_injected(cost, acc, steer, st0, noise_t[2], b_acc, b_steer, xo, yo)
    plain = prob.stage_risk(cost, acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk(cost, acc[i], steer[i], st0, noise_t, xo, yo, beta_draws=(b_acc[i], b_steer[i]))
        _eq(got["risk"][i], ref["risk"], f"risk[{i}]"); _eq(got["lane"][i], ref["lane"], f"lane[{i}]")
        if cost == "mmd_opt":
            _eq(got["beta"][i], ref["beta"], "beta"); _eq(got["sigma"][i], ref["sigma"], "sigma"); _eq(got["res_beta"][i], ref["res_beta"], "res_beta")
    assert np.isfinite(got["risk"]).all()
    if cost == "mmd_opt":
        assert not np.array_equal(got["res_beta"], plain["res_beta"])          # the injected draws really replaced the device sampler's


def test_two_handles_with_different_inner_sizes_alternate(mods):
    """two live handles on one device whose inner-CEM kernels need different amounts of dynamic shared memory (same kernel, same num_reduced):
    the opt-in limit is per (device, kernel) and must only ever be raised (a second, smaller handle used to lower it)"""
    cem_impl, O = mods
    args = (5, 2, 0.1, 30, "gaussian", 0.02, 0.01)
    big = dict(num_batch=24, maxiter_cem=2, num_samples_cem=100, maxiter_beta_cem=3)
    small = dict(num_batch=24, maxiter_cem=2, num_samples_cem=40, maxiter_beta_cem=3)
    pb = cem_impl.CEM(*args, max_episodes=2, **big)
    ps = cem_impl.CEM(*args, max_episodes=2, **small)          # created AFTER the big one
    ob, os_ = O.OracleCEM(*args, **big), O.OracleCEM(*args, **small)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    idx, xo, yo = _episodes(O, ob, 2, 2)
    stack = lambda a: np.stack([a] * 2)
    for rnd in range(2):
        for prob, ora in ((pb, ob), (ps, os_), (pb, ob)):
            got = prob.solve_batch("mmd_opt", idx, stack(init_state), stack(mean), stack(cov), xo, yo, [v_des] * 2)
            if rnd == 0:
                ref = ora.solve("mmd_opt", idx[0], init_state, mean, cov, xo[0], yo[0], v_des)
                for k in ("cx", "cy", "cost_obs", "res_beta"):
                    _eq(got[k][0], ref[k], f"S={prob.num_samples_cem} {k}")


def test_stage_select_bit_exact(mods):
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0))
    rng = np.random.default_rng(13)
    B = 100
    for trial in range(4):
        res = np.abs(rng.normal(0, 1e-5, B)).astype(f32); risk = np.where(rng.random(B) < 0.6, 0.0, rng.random(B)).astype(f32)
        if trial == 1:
            res[10:30] = res[10]                 # ties in the first key too
        base = rng.uniform(5, 50, B).astype(f32)
        params = np.concatenate([rng.uniform(0.1, 30, (B, 4)), rng.normal(0, 6, (B, 4))], 1).astype(f32)
        mean = rng.normal(5, 3, 8).astype(f32); A = rng.normal(0, 1, (8, 8)); cov = (A @ A.T + 8 * np.eye(8)).astype(f32)
        z = rng.normal(0, 1, (95, 8)).astype(f32)
        g = prob.stage_select("cvar", res, risk, base, params, mean, cov, z)
        r = ora.select("cvar", res, risk, base, params, mean, cov, z)
        _eq(g[0], r[0], "params_next"); _eq(g[1], r[1], "mean"); _eq(g[2], r[2], "cov")
        assert g[3] == r[3]["sel"]


# ---- the three inner-CEM kernels (k_inner_cem_warp: one warp per chain, persistent; k_inner_cem_fast: one CTA per chain; k_inner_cem: generic)
@pytest.mark.parametrize("mode", ["lat512", "pipe", "split", "warp", "cta", "lat", "generic"])
@pytest.mark.parametrize("nr,npr,noise,small", [(5, 30, "gaussian", True), (5, 50, "beta", False), (4, 20, "gaussian", True), (3, 20, "beta", True), (2, 25, "gaussian", True)])
def test_inner_cem_kernel_variants_bit_exact(mods, monkeypatch, mode, nr, npr, noise, small):
    monkeypatch.setenv("MPCMMD_INNER_CEM", mode)
    kw = dict(num_samples_cem=40, maxiter_beta_cem=4) if small else {}          # small=False: reference sizes (100 samples, 20 iterations, 11 elites)
    prob, ora = _pair(mods, (nr, 3, 0.3 if noise == "beta" else 0.1, npr, noise, 0.05, 0.01), **kw)
    rng = np.random.default_rng(17 + nr)
    n = 9
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(777, 2)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(3, 2)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    xo = xo.copy(); xo[0] = np.linspace(2, 60, 100); yo = yo.copy(); yo[0] = 1.75
    got = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk("mmd_opt", acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(got[k][i], ref[k], f"{mode} nr={nr} chain {i} {k}")


@pytest.mark.parametrize("mode", ["cta", "pipe", "split"])
def test_inner_cem_repeatable_under_load(mods, monkeypatch, mode):
    """Race detector of last resort (compute-sanitizer is closed on this GPU pool, profiles/r02_sanitizer.md): the kernels that alias live
    shared-memory regions across barriers (k_inner_cem_fast), hand-roll a named-barrier producer/consumer protocol (k_inner_cem_pipe) or pass
    chain state through L2 between launches (k_icem_*) run a launch of several waves of CTAs four times; every output bit must repeat, and
    sampled chains must equal the oracle.  A data race or a missing fence shows up as run-to-run or chain-to-chain differences."""
    monkeypatch.setenv("MPCMMD_INNER_CEM", mode)
    kw = dict(num_samples_cem=40, maxiter_beta_cem=5)
    prob, ora = _pair(mods, (5, 2, 0.1, 20, "gaussian", 0.0, 0.0), max_episodes=60, **kw)
    rng = np.random.default_rng(41)
    base_acc, base_steer = _controls(ora, rng, 12)
    n = 5003                                                   # odd: the two-chain CTAs of the pipelined kernel end on a single chain
    pick = rng.integers(0, 12, n)
    acc = (base_acc[pick] * rng.uniform(0.5, 1.5, (n, 1))).astype(f32); steer = (base_steer[pick] * rng.uniform(0.5, 1.5, (n, 1))).astype(f32)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(57, 3)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(2, 1)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    first = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
    for rep in range(3):
        again = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(again[k], first[k], f"{mode} repeat {rep} {k}")
    for i in list(range(0, n, 211)) + [n - 1]:
        ref = ora.risk("mmd_opt", acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(first[k][i], ref[k], f"{mode} chain {i} {k}")


@pytest.mark.parametrize("nr,npr,noise,kw,n", [
    (12, 20, "gaussian", dict(num_samples_cem=40, maxiter_beta_cem=4), 5),
    (20, 60, "gaussian", dict(num_samples_cem=40, maxiter_beta_cem=3), 4),           # configs[3]: num_reduced 20, num_prime 60
    (20, 20, "beta", dict(num_samples_cem=100, maxiter_beta_cem=3), 2),              # reference inner-CEM population (100 samples, 11 elites)
    (40, 20, "gaussian", dict(num_samples_cem=40, maxiter_beta_cem=2), 2)])          # d = 1601: 10 MB covariance per chain
def test_stage_risk_large_reduced_set_mmd_opt(mods, nr, npr, noise, kw, n):
    """mmd_opt for num_reduced > 10 (BASELINE configs[3] sweeps {5, 10, 20, 40}): k_inner_cem_big keeps the chain state (distance table nm x nm, covariance
    (nm+1)^2, samples) in global memory and runs the same phases as the shared-memory kernels -- bit for bit the oracle, which is generic in num_reduced
    (compute_beta.py:51-68: jnp.cov + multivariate_normal of an (nr^2+1)^2 covariance)."""
    prob, ora = _pair(mods, (nr, 2, 0.3 if noise == "beta" else 0.1, npr, noise, 0.05, 0.01), **kw)
    rng = np.random.default_rng(31 + nr)
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(4242, 3)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(2, 5)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    xo = xo.copy(); xo[0] = np.linspace(2, 60, 100); yo = yo.copy(); yo[0] = 1.75
    got = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk("mmd_opt", acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(got[k][i], ref[k], f"nr={nr} chain {i} {k}")
    assert np.isfinite(got["res_beta"]).all()


def test_inner_cem_warp_persistent_grid_strides_over_chains(mods, monkeypatch):
    """more chains than persistent CTAs (SMs x resident warps): every CTA of k_inner_cem_warp runs several chains and reuses its stash"""
    monkeypatch.setenv("MPCMMD_INNER_CEM", "warp")
    kw = dict(num_samples_cem=40, maxiter_beta_cem=3)
    prob, ora = _pair(mods, (5, 2, 0.1, 20, "gaussian", 0.0, 0.0), max_episodes=80, **kw)
    rng = np.random.default_rng(23)
    base_acc, base_steer = _controls(ora, rng, 16)
    n = 7000
    pick = rng.integers(0, 16, n)
    acc = (base_acc[pick] * rng.uniform(0.5, 1.5, (n, 1))).astype(f32); steer = (base_steer[pick] * rng.uniform(0.5, 1.5, (n, 1))).astype(f32)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(31, 5)
    sc = __import__("oracle.oracle", fromlist=["x"]).static_scene(2, 1)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    got = prob.stage_risk("mmd_opt", acc, steer, st0, noise_t, xo, yo)
    for i in list(range(0, n, 97)) + [n - 1]:
        ref = ora.risk("mmd_opt", acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane", "beta", "sigma", "res_beta"):
            _eq(got[k][i], ref[k], f"chain {i} {k}")


def _episodes(O, ora, n, nobs):
    eps = [O.static_episode(nobs, k) for k in range(n)]
    tr = [ora.compute_obs_trajectories(*sc) for sc, _ in eps]
    return [i for _, i in eps], np.stack([t[0] for t in tr]), np.stack([t[1] for t in tr])


@pytest.mark.parametrize("cost,noise", [("mmd_opt", "gaussian"), ("cvar", "gaussian"), ("cvar", "beta"), ("mmd_random", "gaussian"), ("saa", "gaussian")])
def test_solve_small_bit_exact(mods, cost, noise):
    cem_impl, O = mods
    prob, ora = _pair(mods, (5, 2, 0.3 if noise == "beta" else 0.1, 30, noise, 0.02, 0.01), **SMALL)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    E = 3
    idx, xo, yo = _episodes(O, ora, E, 2)
    got = prob.solve_batch(cost, idx, np.stack([init_state] * E), np.stack([mean] * E), np.stack([cov] * E), xo, yo, [v_des] * E)
    for e in range(E):
        ref = ora.solve(cost, idx[e], init_state, mean, cov, xo[e], yo[e], v_des)
        for k in ("cx", "cy", "cost_obs", "cost_lane"):
            _eq(got[k][e], ref[k], f"{cost}/{noise} ep{e} {k}")
        if cost == "mmd_opt":
            for k in ("beta", "sigma", "res_beta"):
                _eq(got[k][e], ref[k], f"ep{e} {k}")


def test_solve_full_size_cvar_matches_oracle(mods):
    """BASELINE cfg2 shape (static, cvar, beta noise 0.3, 4 obstacles, num_prime 50), full CEM sizes, 6 episodes."""
    cem_impl, O = mods
    prob, ora = _pair(mods, (5, 4, 0.3, 50, "beta", 0.0, 0.0), max_episodes=6)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    E = 6
    idx, xo, yo = _episodes(O, ora, E, 4)
    got = prob.solve_batch("cvar", idx, np.stack([init_state] * E), np.stack([mean] * E), np.stack([cov] * E), xo, yo, [v_des] * E)
    for e in range(E):
        ref = ora.solve("cvar", idx[e], init_state, mean, cov, xo[e], yo[e], v_des)
        for k in ("cx", "cy", "cost_obs", "cost_lane"):
            _eq(got[k][e], ref[k], f"ep{e} {k}")


def test_solve_full_size_mmd_opt_matches_oracle(mods):
    """BASELINE cfg1: static, mmd_opt, gaussian 0.1, nr 5, 2 obstacles, num_prime 30, full sizes, episode 0 (oracle ~8 s)."""
    cem_impl, O = mods
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), max_episodes=1)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    idx, xo, yo = _episodes(O, ora, 1, 2)
    got = prob.compute_cem_mmd_opt(idx[0], init_state, mean, cov, xo[0], yo[0], v_des)
    ref = ora.solve("mmd_opt", idx[0], init_state, mean, cov, xo[0], yo[0], v_des)
    for g, k in zip(got, ("cx", "cy", "cost_lane", "cost_obs", "beta", "sigma", "res_beta")):
        _eq(g, ref[k], k)


def test_drop_in_module_surface(mods):
    """`from optimizer import cem` exactly as synthetic_static_obs/main_mpc.py:5-6 does, and the attributes callers read."""
    import importlib, os, sys
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpc-mmd_b200", "synthetic_dynamic_obs")
    sys.path.insert(1, root)
    try:
        sys.modules.pop("optimizer", None); sys.modules.pop("optimizer.cem", None)
        cem = importlib.import_module("optimizer.cem")
        prob = cem.CEM(5, 2, 0.1, 30, "gaussian", 0.0, 0.0)
        assert (prob.y_lb, prob.y_ub, prob.cem_helper.K_steer) == (-2.25, -1.25, 0.05)
        for a in ("nvar", "num_obs", "num", "num_prime", "ker_wt", "P", "Pdot", "Pdot_jax", "Pddot_jax", "t", "wheel_base", "beta_a", "beta_b",
                  "a_obs", "b_obs", "acc_const_noise", "steer_const_noise"):
            assert hasattr(prob, a), a
        x, y, psi = prob.cem_helper.compute_obs_trajectories(np.array([30.0, 40.0]), np.array([1.75, -1.75]), np.array([1.0, 0.0]), np.zeros(2), np.zeros(2))
        assert x.shape == (2, 100) and abs(x[0, -1] - 45.0) < 1e-4
    finally:
        sys.path.remove(root); sys.modules.pop("optimizer", None); sys.modules.pop("optimizer.cem", None)


@pytest.mark.parametrize("variant,num_obs,num_prime,noise", [("static", 4, 50, "beta"), ("dynamic", 3, 30, "gaussian")])
def test_sweep_driver_writes_reference_schema(mods, tmp_path, variant, num_obs, num_prime, noise):
    """mpcmmd_b200.driver (= main_mpc.py of the two variants): 4 episodes of a cvar sweep point, full solver sizes.  The saved arrays
    must equal the oracle's solve of every accepted episode, in episode order, with the reference's file name and keys."""
    cem_impl, O = mods
    from mpcmmd_b200 import driver, scenes
    argv = ["--costs", "cvar", "--noises", noise, "--noise_levels", "0.3", "--num_reduced_sets", "5", "--num_obs", str(num_obs),
            "--num_prime", str(num_prime), "--acc_const_noise", "0.0", "--steer_const_noise", "0.0", "--num_configs", "4",
            "--root", str(tmp_path), "--variant", variant]
    args = driver.build_parser().parse_args(argv)
    written = driver.run_sweep(args, log=lambda *a: None)
    assert written == [str(tmp_path / f"{noise}_noise" / "noise_30" / f"ts_{num_prime}" / f"cvar_5_samples_{num_obs}_obs.npz")]
    data = np.load(written[0])
    ora = O.OracleCEM(5, num_obs, 0.3, num_prime, noise, 0.0, 0.0, variant=variant)
    init_state, mean, cov, v_des = O.driver_inputs(variant)
    rows = []
    for k in range(4):
        if variant == "dynamic":
            sc, idx, xt, yt = scenes.dynamic_scene(num_obs, k)
        else:
            sc, idx = O.static_episode(num_obs, k)
            xt, yt, _ = ora.compute_obs_trajectories(*sc)
        ref = ora.solve("cvar", idx, init_state, mean, cov, xt, yt, v_des)
        if ref["cost_obs"] <= 1e-5:
            rows.append((k, ref, sc, xt, yt))
    assert data["cx"].shape == (len(rows), 11) and data["cx"].dtype == np.float64
    for r, (k, ref, sc, xt, yt) in enumerate(rows):
        _eq(data["cx"][r].astype(f32), ref["cx"], f"episode {k} cx"); _eq(data["cy"][r].astype(f32), ref["cy"], f"episode {k} cy")
        _eq(data["x_obs"][r], np.asarray(sc[0], np.float64), "x_obs"); _eq(data["vx_obs"][r], np.asarray(sc[2], np.float64), "vx_obs")
        if variant == "dynamic":
            _eq(data["x_obs_traj"][r].astype(f32), xt, "x_obs_traj"); _eq(data["y_obs_traj"][r].astype(f32), yt, "y_obs_traj")


def _scaled_episodes(ora, n, nobs):
    from mpcmmd_b200 import scenes
    eps = [scenes.scaled_scene(nobs, k) for k in range(n)]
    tr = [ora.compute_obs_trajectories(*sc) for sc, _ in eps]
    return [i for _, i in eps], np.stack([t[0] for t in tr]), np.stack([t[1] for t in tr])


@pytest.mark.parametrize("cost,noise,B,kw", [
    ("cvar", "beta", 2048, dict(maxiter_cem=20)),                                                   # large-batch selection path, full depth
    ("mmd_opt", "gaussian", 1100, dict(maxiter_cem=3, num_samples_cem=40, maxiter_beta_cem=4)),    # just above SEL_RANK_MAX, 27 500 inner chains
    ("cvar", "gaussian", 16384, dict(maxiter_cem=4)),                                               # BASELINE configs[4] shape: 16k samples x 32 obstacles x 100 steps
])
def test_solve_scaled_batch_bit_exact(mods, cost, noise, B, kw):
    """num_batch far above the reference's 100 (BASELINE configs[4] 'scaled synthetic': 16k CEM samples, 32 obstacles, num_prime 100):
    k_select takes its O(n_el_cost * B) path, k_project / k_rollouts run B-sized grids.  Bit exact vs the oracle."""
    cem_impl, O = mods
    E = 2
    nobs, npr = (32, 100) if B == 16384 else (12, 60)
    prob, ora = _pair(mods, (5, nobs, 0.3 if noise == "beta" else 0.1, npr, noise, 0.01, 0.0), max_episodes=E, num_batch=B, **kw)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    idx, xo, yo = _scaled_episodes(ora, E, nobs)
    got = prob.solve_batch(cost, idx, np.stack([init_state] * E), np.stack([mean] * E), np.stack([cov] * E), xo, yo, [v_des] * E)
    for e in range(E):
        ref = ora.solve(cost, idx[e], init_state, mean, cov, xo[e], yo[e], v_des)
        for k in ("cx", "cy", "cost_obs", "cost_lane") + (("beta", "sigma", "res_beta") if cost == "mmd_opt" else ()):
            _eq(got[k][e], ref[k], f"B={B} {cost} ep{e} {k}")


@pytest.mark.parametrize("cost,nobs,special", [("cvar", 32, "plain"), ("cvar", 32, "nan_obstacle"), ("mmd_random", 20, "moving"), ("mmd_opt", 32, "edge"),
                                               ("cvar", 64, "plain")])
def test_stage_risk_many_obstacles_sorted_window(mods, cost, nobs, special):
    """num_obs > 8 (configs[4]: 32 obstacles): the rollouts look only at the obstacles inside an x window around the vehicle, found by binary search in the
    per-knot sorted obstacle lists of k_obs_sort.  Skipped obstacles contribute exactly 0, so the result must equal the oracle's all-pairs maximum bit for bit --
    with static and moving obstacles, with obstacles placed right at the window edge (|dx| = a_obs +- ulps), and with a NaN obstacle coordinate (plain-loop fallback)."""
    kw = dict(num_samples_cem=40, maxiter_beta_cem=3) if cost == "mmd_opt" else {}
    prob, ora = _pair(mods, (5, nobs, 0.1, 60, "gaussian", 0.02, 0.01), max_episodes=1, **kw)
    rng = np.random.default_rng(300 + nobs)
    n = 24 if cost != "mmd_opt" else 6
    acc, steer = _controls(ora, rng, n)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(91, 4)
    t = np.linspace(0.0, 15.0, 100).astype(f32)
    x0 = rng.uniform(3.0, 90.0, nobs).astype(f32); y0 = rng.choice([-1.75, 1.75], nobs).astype(f32)
    vx = (rng.uniform(0.0, 6.0, nobs) if special == "moving" else np.zeros(nobs)).astype(f32)
    xo = (x0[:, None] + vx[:, None] * t[None]).astype(f32); yo = np.repeat(y0[:, None], 100, 1).astype(f32)
    if special == "edge":          # obstacles exactly a_obs = 4.25 (and one ulp either side) ahead of where a nominal rollout is at knots 5, 10, 15
        xr = ora.risk("cvar", acc[0], steer[0], st0, noise_t, xo, yo, want_rollouts=True)["x_roll"][0]
        for j, (kn, du) in enumerate(((5, 0), (10, 1), (15, -1), (20, 2))):
            xe = np.nextafter(f32(xr[kn] + 4.25), f32(np.inf) if du > 0 else f32(-np.inf)) if du else f32(xr[kn] + 4.25)
            xo[j, :] = xe; yo[j, :] = 1.75
    if special == "nan_obstacle":
        xo[3, 17] = np.nan; yo[5, 40] = np.nan
    got = prob.stage_risk(cost, acc, steer, st0, noise_t, xo, yo)
    for i in range(n):
        ref = ora.risk(cost, acc[i], steer[i], st0, noise_t, xo, yo)
        _eq(got["risk"][i], ref["risk"], f"{special} risk[{i}]"); _eq(got["lane"][i], ref["lane"], f"{special} lane[{i}]")
    if special != "nan_obstacle":
        assert np.isfinite(got["risk"]).all() and np.any(got["risk"] != got["risk"][0])


@pytest.mark.parametrize("nr,npr", [(10, 20), (20, 60), (40, 100), (40, 20)])
@pytest.mark.parametrize("cost", ["mmd_random", "cvar"])
def test_stage_risk_large_reduced_sets(mods, cost, nr, npr):
    """BASELINE configs[3]: the MMD / CVaR risk stage over num_reduced in {5,10,20,40} x num_prime in {20..100} (num_reduced rollouts per
    sample, warp-parallel Laplace-kernel MMD); bit exact vs the oracle on 64 samples"""
    cem_impl, O = mods
    prob, ora = _pair(mods, (nr, 6, 0.1, npr, "gaussian", 0.02, 0.01), max_episodes=1)
    rng = np.random.default_rng(nr * 1000 + npr)
    acc, steer = _controls(ora, rng, 64)
    st0 = np.array([0.0, 1.75, 5.0, 0.0, 0.0], f32)
    noise_t = ora.noise_tables(77, 2)
    from mpcmmd_b200 import scenes
    sc, _ = scenes.static_scene(6, 3)
    xo, yo, _ = ora.compute_obs_trajectories(*sc)
    got = prob.stage_risk(cost, acc, steer, st0, noise_t, xo, yo)
    for i in range(64):
        ref = ora.risk(cost, acc[i], steer[i], st0, noise_t, xo, yo)
        for k in ("risk", "lane"):
            _eq(got[k][i], ref[k], f"{cost} nr={nr} np={npr} sample {i} {k}")


def test_degenerate_inputs_propagate_nan_like_the_oracle(mods):
    """edge cases the reference handles silently ('Errors'): a covariance that is not positive definite (Cholesky -> NaN) and
    a vehicle at rest (zero speed -> 0/0 curvature).  NaN must sort last in every argsort and flow to the outputs exactly as in the oracle."""
    cem_impl, O = mods
    prob, ora = _pair(mods, (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), **SMALL)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    idx, xo, yo = _episodes(O, ora, 3, 2)
    bad_cov = cov.copy(); bad_cov[2, 2] = -1.0                       # not PD -> NaN rows from the third column on
    rest = init_state.copy(); rest[2] = 0.0                          # vx = vy = 0
    st = np.stack([init_state, init_state, rest]); cv = np.stack([cov, bad_cov, cov])
    for cost in ("cvar", "mmd_opt"):
        got = prob.solve_batch(cost, idx, st, np.stack([mean] * 3), cv, xo, yo, [v_des] * 3)
        for e in range(3):
            ref = ora.solve(cost, idx[e], st[e], mean, cv[e], xo[e], yo[e], v_des)
            for k in ("cx", "cy", "cost_obs", "cost_lane"):
                _eq(got[k][e], ref[k], f"{cost} ep{e} {k}")
    assert np.isnan(got["cx"][1]).any() or np.isfinite(got["cx"][1]).all()      # whichever it is, it equals the oracle (checked above)


@pytest.mark.parametrize("noise,level", [("gaussian", 0.1), ("beta", 0.3)])
def test_solve_full_size_dynamic_mmd_opt_matches_oracle(mods, noise, level):
    """BASELINE configs[2]: synthetic_dynamic_obs, mmd_opt, 6 obstacles, num_prime 60, full CEM sizes, episode 3 of the sweep (cut-in scene
    from the dynamic generator); oracle ~10 s"""
    cem_impl, O = mods
    from mpcmmd_b200 import scenes
    prob, ora = _pair(mods, (5, 6, level, 60, noise, 0.0, 0.0), variant="dynamic", max_episodes=1)
    init_state, mean, cov, v_des = O.driver_inputs("dynamic")
    _, idx, xt, yt = scenes.dynamic_scene(6, 3)
    got = prob.compute_cem_mmd_opt(idx, init_state, mean, cov, xt, yt, v_des)
    ref = ora.solve("mmd_opt", idx, init_state, mean, cov, xt, yt, v_des)
    for g, k in zip(got, ("cx", "cy", "cost_lane", "cost_obs", "beta", "sigma", "res_beta")):
        _eq(g, ref[k], f"{noise} {k}")


def test_full_sweep_is_batch_invariant(mods):
    """size-independent property at BASELINE configs[1] full size (200 episodes x 100 samples, both costs): an episode solved inside the
    200-episode batch equals the same episode solved alone, bit for bit (episodes are independent), and the episode solved
    alone equals the oracle (cvar)."""
    cem_impl, O = mods
    from mpcmmd_b200 import scenes
    args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
    big = cem_impl.CEM(*args, variant="static", max_episodes=200)
    one = cem_impl.CEM(*args, variant="static", max_episodes=1)
    batch = scenes.static_batch(big, list(range(200)))
    pick = [0, 57, 131, 199]
    for cost in ("cvar", "mmd_opt"):
        out = big.solve_batch(cost, **batch)
        assert out["cx"].shape == (200, 11) and np.isfinite(out["cx"]).all()
        for k in pick:
            solo = one.solve_batch(cost, **{n: v[k:k + 1] for n, v in batch.items()})
            for f in ("cx", "cy", "cost_obs", "cost_lane") + (("beta", "sigma", "res_beta") if cost == "mmd_opt" else ()):
                _eq(out[f][k], solo[f][0], f"{cost} episode {k} {f}: batch vs alone")
    ora = O.OracleCEM(*args, variant="static")
    init_state, mean, cov, v_des = O.driver_inputs("static")
    out = big.solve_batch("cvar", **batch)
    for k in pick:
        ref = ora.solve("cvar", int(batch["idx_mpc"][k]), init_state, mean, cov, batch["x_obs_traj"][k], batch["y_obs_traj"][k], v_des)
        _eq(out["cx"][k], ref["cx"], f"episode {k} cx vs oracle"); _eq(out["cost_obs"][k], ref["cost_obs"], f"episode {k} cost_obs vs oracle")


@pytest.mark.parametrize("E,groups", [(13, None), (25, None), (25, "5"), (7, "8")])
def test_solve_graph_episode_groups_are_invariant(mods, monkeypatch, E, groups):
    """strong-scaled shards: an mmd_opt solve of 13 / 25 episodes is captured as 2 / 3 episode groups on parallel graph branches (solve_groups; MPCMMD_GROUPS forces
    other counts, capped by the episode count: ragged groups, more groups than the automatic rule ever picks).  Every episode equals the same episode solved alone."""
    cem_impl, _ = mods
    from mpcmmd_b200 import scenes
    if groups:
        monkeypatch.setenv("MPCMMD_GROUPS", groups)
    args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
    big = cem_impl.CEM(*args, variant="static", max_episodes=E)
    monkeypatch.delenv("MPCMMD_GROUPS", raising=False)
    one = cem_impl.CEM(*args, variant="static", max_episodes=1)
    batch = scenes.static_batch(big, list(range(E)))
    out = big.solve_batch("mmd_opt", **batch)
    for k in sorted({0, E // 3, E // 2, (2 * E) // 3, E - 1}):
        solo = one.solve_batch("mmd_opt", **{n: v[k:k + 1] for n, v in batch.items()})
        for f in ("cx", "cy", "cost_obs", "cost_lane", "beta", "sigma", "res_beta"):
            _eq(out[f][k], solo[f][0], f"E={E} groups={groups} episode {k} {f}: grouped graph vs alone")


def test_full_sweep_matches_oracle_episode_by_episode(mods):
    """BASELINE configs[1] at full size: ALL 200 episodes of the cvar sweep and 12 episodes of the mmd_opt sweep, solved in one batch each,
    against the oracle episode by episode -- trajectories, costs, and the accepted set that main_mpc.py would write (cost_obs <= threshold)."""
    cem_impl, O = mods
    from mpcmmd_b200 import scenes
    args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
    big = cem_impl.CEM(*args, variant="static", max_episodes=200)
    ora = O.OracleCEM(*args, variant="static")
    O.set_threads(__import__("os").cpu_count() or 1)
    init_state, mean, cov, v_des = O.driver_inputs("static")
    batch = scenes.static_batch(big, list(range(200)))
    out = big.solve_batch("cvar", **batch)
    acc_gpu, acc_ref = [], []
    for k in range(200):
        ref = ora.solve("cvar", int(batch["idx_mpc"][k]), init_state, mean, cov, batch["x_obs_traj"][k], batch["y_obs_traj"][k], v_des)
        for f in ("cx", "cy", "cost_obs", "cost_lane"):
            _eq(out[f][k], ref[f], f"cvar episode {k} {f}")
        acc_gpu.append(bool(out["cost_obs"][k] <= 1e-5)); acc_ref.append(bool(ref["cost_obs"] <= 1e-5))
    assert acc_gpu == acc_ref and 100 < sum(acc_gpu) < 200
    sub = {n: v[:12] for n, v in batch.items()}
    out = big.solve_batch("mmd_opt", **sub)
    for k in range(12):
        ref = ora.solve("mmd_opt", int(sub["idx_mpc"][k]), init_state, mean, cov, sub["x_obs_traj"][k], sub["y_obs_traj"][k], v_des)
        for f in ("cx", "cy", "cost_obs", "cost_lane", "beta", "sigma", "res_beta"):
            _eq(out[f][k], ref[f], f"mmd_opt episode {k} {f}")
