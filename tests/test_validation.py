"""Monte-Carlo validation (`compute_stats`, S/validation.py:134-171): oracle vs the golden vectors recorded from the reference's own
functions (CPU), host logic of the drop-in (CPU), and the CUDA path vs both (GPU)."""
import os
import sys
import types

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "validation_ref.npz")
f32 = np.float32


def _cases():
    g = np.load(GOLD)
    return g, [str(n) for n in g["names"]]


def _consts(variant, args, ref_basis=False):
    from mpcmmd_b200.constants import VARIANT_CONSTANTS, build_constants
    hc = build_constants(int(args[3]))
    vc = VARIANT_CONSTANTS[variant]
    if ref_basis:
        # the reference's own basis matrices (tests/golden/bernstein_ref.npz): ours differ from them in 4 of 1100 Pddot entries, all
        # mathematically zero (t = 1/3, 2/3), by ~1e-17 -- enough to move a float64 rollout by one ulp, so the bit-level pin uses the reference's
        gb = np.load(os.path.join(HERE, "golden", "bernstein_ref.npz"))
        return dict(Pdot=gb["Pdot"].astype(f32), Pddot=gb["Pddot"].astype(f32), t=15 / 100, wheel_base=2.5, a_obs=4.25, b_obs=2.75, y_lb=vc["y_lb"],
                    y_ub=vc["y_ub"], K_steer=vc["K_steer"], acc_const=float(args[5]), steer_const=float(args[6])), hc
    return dict(Pdot=hc.Pdot, Pddot=hc.Pddot, t=15 / 100, wheel_base=2.5, a_obs=4.25, b_obs=2.75, y_lb=vc["y_lb"], y_ub=vc["y_ub"],
                K_steer=vc["K_steer"], acc_const=float(args[5]), steer_const=float(args[6])), hc


def _case_inputs(g, name, ref_basis=False):
    args = g[name + "_args"]
    variant = "static" if name.startswith("S_") else "dynamic"
    num_obs, noise_level, num_prime, noise, key = int(args[1]), float(args[2]), int(args[3]), ("gaussian", "beta")[int(args[4])], int(args[7])
    c, hc = _consts(variant, args, ref_basis)
    if variant == "static":
        tt = hc.tot_time.astype(f32)[:, None]                               # compute_obs_trajectories in float32 (cem_helper.py:366-378)
        xt = (g[name + "_x_obs"].astype(f32) + g[name + "_vx_obs"].astype(f32) * tt).T
        yt = (g[name + "_y_obs"].astype(f32) + g[name + "_vy_obs"].astype(f32) * tt).T
    else:
        xt, yt = g[name + "_x_obs_traj"].astype(np.float64), g[name + "_y_obs_traj"].astype(np.float64)
    return variant, c, num_obs, noise_level, num_prime, noise, key, xt, yt


@pytest.mark.parametrize("name", _cases()[1])
def test_oracle_matches_reference_compute_stats(built, name):
    """oracle/oracle_validation.py vs the reference's own compute_stats: counts identical, rollouts bit for bit (same NumPy ops)"""
    from oracle import oracle_validation as OV
    g, _ = _cases()
    variant, c, num_obs, noise_level, num_prime, noise, key, xt, yt = _case_inputs(g, name, ref_basis=True)
    count, lane, xr, yr = OV.compute_stats(c, g[name + "_cx"], g[name + "_cy"], g[name + "_init_state"], xt, yt, num_prime, noise_level, noise, key,
                                           obs_f32=(variant == "static"))
    assert [count, lane] == g[name + "_count"].tolist()
    assert np.array_equal(np.stack([xr[:4], yr[:4]]), g[name + "_roll_head"])
    assert np.array_equal(np.array([xr.sum(), yr.sum()]), g[name + "_roll_sum"])


def _fake_prob(variant, args):
    """the attributes mpcmmd_b200.validation reads from a CEM object, without a GPU"""
    c, hc = _consts(variant, args)
    from mpcmmd_b200.cem_impl import _HelperShim
    p = types.SimpleNamespace(Pdot_jax=hc.Pdot, Pddot_jax=hc.Pddot, t=15 / 100, wheel_base=2.5, a_obs=4.25, b_obs=2.75, y_lb=c["y_lb"], y_ub=c["y_ub"],
                              beta_a=2, beta_b=5, acc_const_noise=float(args[5]), steer_const_noise=float(args[6]), num=100,
                              tot_time=hc.tot_time, _K_steer=c["K_steer"])
    p.cem_helper = _HelperShim(p)
    return p


@pytest.mark.parametrize("name", ["S_beta_solve0", "S_gauss_solve3", "D_beta_solve2"])
def test_host_controls_and_noise_match_oracle(built, name):
    """the host half of the drop-in (planned controls + legacy-stream noise) is bit-identical to the oracle's restatement"""
    from mpcmmd_b200 import validation as V
    from oracle import oracle_validation as OV
    g, _ = _cases()
    variant, c, num_obs, noise_level, num_prime, noise, key, xt, yt = _case_inputs(g, name)
    prob = _fake_prob(variant, g[name + "_args"])
    acc, steer = V.compute_controls(prob, g[name + "_cx"], g[name + "_cy"])
    acc_o, steer_o = OV.controls(c["Pdot"], c["Pddot"], g[name + "_cx"], g[name + "_cy"], c["t"], c["wheel_base"])
    assert np.array_equal(acc, acc_o) and np.array_equal(steer, steer_o) and acc.shape == (101,) and acc[99] == 0.0 and acc[100] == 0.0
    a, s = V.perturbed_controls(prob, acc[:num_prime], steer[:num_prime], noise_level, num_prime, noise, key)
    a_o, s_o = OV.noisy_controls(acc_o[:num_prime], steer_o[:num_prime], noise_level, num_prime, noise, key, c["K_steer"], c["acc_const"], c["steer_const"])
    assert np.array_equal(a, a_o) and np.array_equal(s, s_o) and a.shape == (1000, num_prime)


def test_matched_pairs_follow_reference_enumeration(built):
    """scenes solved by both costs, enumerated as validation.py:290-313 does (set intersection order, first matching row)"""
    from mpcmmd_b200 import validation as V
    rng = np.random.default_rng(0)
    rows = rng.integers(0, 5, (9, 6 + 4 * 2)).astype(np.float64)
    mk = lambda idx: dict(init_state=rows[idx, :6], x_obs=rows[idx, 6:8], y_obs=rows[idx, 8:10], vx_obs=rows[idx, 10:12], vy_obs=rows[idx, 12:14])
    d_cvar, d_opt = mk([0, 1, 2, 3, 4, 5, 1]), mk([8, 5, 3, 1, 7])
    pairs = V.matched_pairs(d_cvar, d_opt, 2)
    cset = set(tuple(x) for x in rows[[0, 1, 2, 3, 4, 5, 1]]); dset = set(tuple(x) for x in rows[[8, 5, 3, 1, 7]])
    want = [x for x in cset & dset]
    assert len(pairs) == 3 and [k for k, _, _ in pairs] == [0, 1, 2]
    for (k, ic, im), w in zip(pairs, want):
        assert tuple(rows[[0, 1, 2, 3, 4, 5, 1][ic]]) == w and tuple(rows[[8, 5, 3, 1, 7][im]]) == w
    assert [ic for _, ic, _ in pairs if tuple(rows[1]) == tuple(rows[[0, 1, 2, 3, 4, 5, 1][ic]])] == [1]       # duplicate row -> first index


# ---- GPU -------------------------------------------------------------------------------------------------------------------------
def _items(g, names):
    items = []
    for name in names:
        it = dict(cx=g[name + "_cx"], cy=g[name + "_cy"], init_state=g[name + "_init_state"], key=int(g[name + "_args"][7]))
        if name.startswith("S_"):
            it.update(x_obs=g[name + "_x_obs"], y_obs=g[name + "_y_obs"], vx_obs=g[name + "_vx_obs"], vy_obs=g[name + "_vy_obs"])
        else:
            it.update(x_obs_traj=g[name + "_x_obs_traj"].astype(np.float64), y_obs_traj=g[name + "_y_obs_traj"].astype(np.float64))
        items.append(it)
    return items


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["S_beta", "S_gauss", "D_gauss", "D_beta"])
def test_cuda_compute_stats_matches_reference_golden(built, group):
    """mpcmmd_validate_host vs the counts the reference's own compute_stats produced: identical collision / lane counts; rollouts within
    1e-12 relative (CUDA's float64 sin/cos/tan are not glibc's: <= 2 ulp per call, accumulated over <= 60 steps)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpcmmd_b200 import CEM, validation as V
    g, names = _cases()
    names = [n for n in names if n.startswith(group)]
    args = g[names[0] + "_args"]
    variant = "static" if group.startswith("S_") else "dynamic"
    num_obs, noise_level, num_prime, noise = int(args[1]), float(args[2]), int(args[3]), ("gaussian", "beta")[int(args[4])]
    prob = CEM(int(args[0]), num_obs, noise_level, num_prime, noise, float(args[5]), float(args[6]), variant=variant)
    count, lane, xr, yr = V.compute_stats_batch(prob, _items(g, names), num_prime, noise_level, noise, num_obs, want_rollouts=True)
    for i, name in enumerate(names):
        assert [int(count[i]), int(lane[i])] == g[name + "_count"].tolist(), name
        head = np.stack([xr[i, :4], yr[i, :4]])
        assert np.allclose(head, g[name + "_roll_head"], rtol=1e-12, atol=1e-12), name
        assert np.allclose([xr[i].sum(), yr[i].sum()], g[name + "_roll_sum"], rtol=1e-12)


@pytest.mark.gpu
def test_cuda_compute_stats_matches_oracle_on_solver_output(built):
    """full loop at BASELINE cfg2 sizes: solve 6 static episodes on the GPU, validate the plans on the GPU, compare the counts with the CPU
    oracle of compute_stats (1000 rollouts x 50 steps x 4 obstacles per plan)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpcmmd_b200 import CEM, scenes, validation as V
    from oracle import oracle_validation as OV
    args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
    prob = CEM(*args, variant="static", max_episodes=6)
    eps = list(range(6))
    batch = scenes.static_batch(prob, eps)
    out = prob.solve_batch("cvar", **batch)
    init_state = scenes.driver_inputs("static")[0]
    items = []
    for k in eps:
        (x, y, vx, vy, _), _ = scenes.static_scene(4, k)
        items.append(dict(cx=out["cx"][k].astype(np.float64), cy=out["cy"][k].astype(np.float64), init_state=init_state.astype(np.float64), key=k,
                          x_obs=x.astype(np.float64), y_obs=y.astype(np.float64), vx_obs=vx, vy_obs=vy))
    count, lane = V.compute_stats_batch(prob, items, 50, 0.3, "beta", 4)
    c, _ = _consts("static", np.array([5, 4, 0.3, 50, 1, 0.0, 0.0, 0]))
    for i, it in enumerate(items):
        xt, yt, _ = prob.cem_helper.compute_obs_trajectories(it["x_obs"], it["y_obs"], it["vx_obs"], it["vy_obs"], np.zeros(4))
        co, lo, _, _ = OV.compute_stats(c, it["cx"], it["cy"], it["init_state"], xt, yt, 50, 0.3, "beta", it["key"], obs_f32=True)
        assert (int(count[i]), int(lane[i])) == (co, lo), (i, count[i], lane[i], co, lo)


@pytest.mark.gpu
def test_validation_workflow_end_to_end(built, tmp_path):
    """README workflow: main_mpc (sweep driver) -> validation, 5 static episodes, cvar + mmd_opt; the stats file has the reference's keys
    and its counts equal the CPU oracle's for every scene both costs solved"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from mpcmmd_b200 import driver, validation as V
    from oracle import oracle_validation as OV
    common = ["--noises", "beta", "--noise_levels", "0.3", "--num_reduced_sets", "5", "--num_obs", "4", "--num_prime", "50",
              "--acc_const_noise", "0.0", "--steer_const_noise", "0.0", "--root", str(tmp_path / "data")]
    driver.run_sweep(driver.build_parser().parse_args(common + ["--costs", "cvar", "mmd_opt", "--num_configs", "5"]), log=lambda *a: None)
    args = V.build_parser().parse_args(common + ["--stats_root", str(tmp_path / "stats")])
    written = V.run_validation(args, "static", log=lambda *a: None)
    assert written == [str(tmp_path / "stats" / "beta_noise" / "noise_30" / "ts_50" / "5_samples_4_obs.npz")]
    st = np.load(written[0])
    assert set(st.files) == {"coll_cvar", "coll_cvar_lane", "coll_mmd_opt", "coll_mmd_opt_lane", "coll_mmd_random", "coll_mmd_random_lane"}
    d_c = np.load(str(tmp_path / "data" / "beta_noise" / "noise_30" / "ts_50" / "cvar_5_samples_4_obs.npz"))
    d_o = np.load(str(tmp_path / "data" / "beta_noise" / "noise_30" / "ts_50" / "mmd_opt_5_samples_4_obs.npz"))
    pairs = V.matched_pairs(d_c, d_o, 4)
    assert len(pairs) >= 1 and st["coll_cvar"].shape == (len(pairs),) and st["coll_mmd_random"].shape == (0,)
    c, hc = _consts("static", np.array([5, 4, 0.3, 50, 1, 0.0, 0.0, 0]))
    tt = hc.tot_time.astype(f32)[:, None]
    for k, ic, im in pairs:
        for d, idx, key_c, key_l in ((d_c, ic, "coll_cvar", "coll_cvar_lane"), (d_o, im, "coll_mmd_opt", "coll_mmd_opt_lane")):
            xt = (d["x_obs"][idx].astype(f32) + d["vx_obs"][idx].astype(f32) * tt).T; yt = (d["y_obs"][idx].astype(f32) + d["vy_obs"][idx].astype(f32) * tt).T
            co, lo, _, _ = OV.compute_stats(c, d["cx"][idx], d["cy"][idx], d["init_state"][idx], xt, yt, 50, 0.3, "beta", k, obs_f32=True)
            assert (st[key_c][k], st[key_l][k]) == (co, lo)
