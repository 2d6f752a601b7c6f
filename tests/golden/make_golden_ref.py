"""Generates tests/golden/ref_stages.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden_ref.py

The reference needs jax==0.3.23 (absent, no network).  `tests/golden/jax_shim/jax` is a NumPy float32 stand-in for the
slice of the JAX API the optimizer uses (see its _core.py header for what it does and does not reproduce); with it on
sys.path the unmodified modules `optimizer/cem.py, cem_helper.py, projection.py, costs.py, compute_beta.py,
kernel_computation.py` of synthetic_static_obs/ and synthetic_dynamic_obs/ import and run.  This script runs
`CEM.compute_cem_*` for a few CEM iterations, records the inputs and outputs of every stage method, and stores them
for a subset of 20 samples (the reference's own top-20 of that iteration).  tests/test_reference_stages.py then
feeds the recorded INPUTS to the oracle (and, on the GPU box, to the CUDA stage entry points) and compares OUTPUTS at
the 1e-4 tolerance north_star states.  Full-solve outputs are deliberately NOT compared: the reference's first elite
stage sorts projection residuals that sit at float32 round-off (~5e-6), so its elite sets are rounding-noise
dependent and two float32 implementations legitimately diverge after one iteration.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
f32 = np.float32

# name, variant dir, CEM args, cost, iterations to run, iterations to keep
CASES = [
    ("A_static_gauss_cvar", "synthetic_static_obs", (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), "cvar", 2, (0, 1)),
    ("B_static_beta_cvar", "synthetic_static_obs", (5, 4, 0.3, 50, "beta", 0.0, 0.0), "cvar", 2, (1,)),
    ("C_dynamic_gauss_cvar", "synthetic_dynamic_obs", (5, 3, 0.1, 20, "gaussian", 0.02, 0.01), "cvar", 2, (1,)),
    ("D_static_gauss_saa", "synthetic_static_obs", (4, 2, 0.2, 30, "gaussian", 0.0, 0.0), "saa", 1, (0,)),
    ("E_static_gauss_mmd_random", "synthetic_static_obs", (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), "mmd_random", 1, (0,)),
]
# mmd_opt: the per-sample chain (rollouts of the num_reduced^2 mother set + ridge fit + reduced-set inner CEM + MMD)
OPT_CASES = [
    ("F_static_gauss_mmd_opt", "synthetic_static_obs", (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), 4),
    ("G_static_beta_mmd_opt", "synthetic_static_obs", (5, 4, 0.3, 50, "beta", 0.0, 0.0), 2),
    ("H_dynamic_gauss_mmd_opt_nr3", "synthetic_dynamic_obs", (3, 3, 0.1, 20, "gaussian", 0.02, 0.01), 3),
]


REAL_JAX = os.environ.get("MPCMMD_GOLDEN_JAX", "shim") == "real"     # make_golden_jax.py: the reference on its OWN runtime (jax==0.3.23), no stand-in
REF_ROOT = os.environ.get("MPCMMD_REFERENCE_ROOT", "/root/reference")
OUT_NAME = "ref_stages_jax.npz" if REAL_JAX else "ref_stages.npz"


def _setup(variant_dir):
    if not REAL_JAX:
        sys.path.insert(0, os.path.join(HERE, "jax_shim"))
    ref = os.path.join(REF_ROOT, variant_dir)
    sys.path.insert(1, ref)
    sys.path.insert(1, os.path.join(ref, "optimizer"))
    sys.path.insert(1, ROOT)
    from optimizer import cem            # the reference's module
    from oracle import oracle as O       # scene generators / driver inputs only
    return cem, O


def _record(obj, names, log):
    for n in names:
        def mk(fn, n):
            def w(*a, **k):
                r = fn(*a, **k)
                log.append((n, a, r))
                return r
            return w
        setattr(obj, n, mk(getattr(obj, n), n))


def _inputs(O, variant_dir, num_obs, episode):
    variant = "static" if "static" in variant_dir else "dynamic"
    init_state, mean, cov, v_des = O.driver_inputs(variant)
    sc, idx = O.static_episode(num_obs, episode)
    if variant == "dynamic":            # moving obstacles in the ego lane (D/main_mpc.py draws them from obs_data; any track exercises the path)
        x, y, vx, vy, psi = sc
        sc = (x, np.full_like(np.asarray(y, float), -1.75), np.full(num_obs, 3.0), np.zeros(num_obs), psi)
    return variant, init_state, mean, cov, v_des, sc, idx


def run_case(name, variant_dir, args, cost, iters, keep):
    cem, O = _setup(variant_dir)
    import jax.numpy as jnp
    prob = cem.CEM(*args)
    variant, init_state, mean, cov, v_des, sc, idx = _inputs(O, variant_dir, args[1], 3)
    xo, yo, _ = prob.cem_helper.compute_obs_trajectories(*[jnp.asarray(np.asarray(v, f32)) for v in sc])
    log = []
    _record(prob.cem_helper, ["sampling_param", "compute_x_guess", "compute_controls", "compute_rollout_baseline_vmap", "compute_cost",
                              "compute_ellite_samples", "compute_shifted_samples"], log)
    _record(prob.projection, ["compute_projection"], log)
    _record(prob.costs, ["compute_cvar_obs_vmap", "compute_cvar_lane_vmap", "compute_saa_obs_vmap", "compute_saa_lane_vmap",
                         "compute_mmd_obs_vmap"], log)
    prob.maxiter_cem = iters
    import contextlib
    import jax
    # real JAX: the solve is @jit-ed, so the stage hooks would see tracers; run it op by op instead (same XLA kernels, concrete arrays)
    with (jax.disable_jit() if REAL_JAX else contextlib.nullcontext()):
        out = getattr(prob, "compute_cem_" + cost)(idx, jnp.asarray(init_state), jnp.asarray(mean), np.asarray(cov, np.float64), xo, yo, v_des)
    g = {"meta.args": np.array([str(a) for a in args]), "meta.cost": np.array(cost), "meta.variant": np.array(variant), "meta.idx_mpc": np.array(idx),
         "meta.v_des": f32(v_des), "init_state": np.asarray(init_state, f32), "x_obs_traj": np.asarray(xo, f32), "y_obs_traj": np.asarray(yo, f32),
         "mean0": np.asarray(mean, f32), "cov0": np.asarray(cov, f32), "meta.iters": np.array(keep),
         "final.cx": np.asarray(out[0], f32), "final.cy": np.asarray(out[1], f32)}
    it, cur = -1, {}
    A = lambda v: np.asarray(v, f32)
    for nm, a, r in log:
        if nm == "sampling_param":
            g["sampling_param.out"] = A(r)
        elif nm == "compute_x_guess":
            it += 1
            cur = {"params": A(a[2]), "beq_x": A(a[0])[0], "beq_y": A(a[1])[0], "cbar_x": A(r[0]), "cbar_y": A(r[1])}
        elif nm == "compute_projection":
            cur.update(lam_x_in=A(a[4]), lam_y_in=A(a[5]), s_lane_in=A(a[10]))
            for k, v in zip(("cx", "cy", "x", "y", "xd", "yd", "xdd", "ydd", "res_norm", "lam_x_out", "lam_y_out", "s_lane_out"), r):
                cur[k] = A(v)
            cur["perm"] = np.argsort(cur["res_norm"], kind="stable")
        elif nm == "compute_controls":
            inv = np.empty(100, np.int64); inv[cur["perm"]] = np.arange(100)
            cur["acc"], cur["steer"] = A(r[0])[inv], A(r[1])[inv]                       # back to batch order; acc is (100,101)
        elif nm == "compute_rollout_baseline_vmap":
            inv = np.empty(100, np.int64); inv[cur["perm"]] = np.arange(100)
            cur["roll_key"] = np.asarray(a[3]).astype(np.uint32); cur["state0"] = A(a[2])
            cur["x_roll"], cur["y_roll"] = A(r[0])[inv], A(r[1])[inv]
        elif nm in ("compute_cvar_obs_vmap", "compute_saa_obs_vmap", "compute_mmd_obs_vmap"):
            inv = np.empty(100, np.int64); inv[cur["perm"]] = np.arange(100)
            rr = A(r) if nm != "compute_mmd_obs_vmap" else A(r).reshape(100)
            r = rr
            cur["risk"] = rr[inv]
            cur["top20"] = cur["perm"][np.argsort(A(r), kind="stable")[:20]]             # batch indices of the 20 rows the reference keeps
        elif nm in ("compute_cvar_lane_vmap", "compute_saa_lane_vmap"):
            cur["lane20"] = A(r)
        elif nm == "compute_cost":
            cur["cost20"] = A(r)
            # the risk-independent part of compute_cost for the same 20 rows, from the reference's own function
            z = jnp.zeros(20)
            cur["cost_base20"] = A(prob.cem_helper.__class__.compute_cost(prob.cem_helper, z, z, *a[2:]))
        elif nm == "compute_ellite_samples":
            cur["idx_ellite"] = np.asarray(r[1]).astype(np.int32)
        elif nm == "compute_shifted_samples":
            cur.update(sel_key=np.asarray(a[0]).astype(np.uint32), mean_prev=A(a[4]), cov_prev=A(a[5]), mean_new=A(r[0]), cov_new=A(r[1]), batch_new=A(r[2]))
            if it in keep:
                S = cur["top20"]
                per_sample = ("lam_x_in", "lam_y_in", "s_lane_in", "cbar_x", "cbar_y", "cx", "cy", "x", "y", "xd", "yd", "xdd", "ydd", "lam_x_out", "lam_y_out",
                              "s_lane_out", "acc", "steer", "x_roll", "y_roll")
                for k, v in cur.items():
                    g[f"it{it}.{k}"] = v[S] if k in per_sample else v
    np.savez_compressed(os.path.join(HERE, "_part_" + name + ".npz"), **g)


def run_opt_case(name, variant_dir, args, n_chains):
    """per-sample mmd_opt chain: Helper.compute_rollout_complete_opt (cem_helper.py:466-538) + Costs.compute_mmd_obs / compute_mmd_lane"""
    cem, O = _setup(variant_dir)
    import jax
    import jax.numpy as jnp
    prob = cem.CEM(*args)
    variant, init_state, mean, cov, v_des, sc, idx = _inputs(O, variant_dir, args[1], 5)
    xo, yo, _ = prob.cem_helper.compute_obs_trajectories(*[jnp.asarray(np.asarray(v, f32)) for v in sc])
    # one real projection step gives realistic controls
    h = prob.cem_helper
    params = h.sampling_param(jnp.asarray(mean), np.asarray(cov, np.float64))
    x0, y0, vx0, vy0, ax0, ay0 = jnp.asarray(init_state)
    bx, by = h.compute_boundary_vec(x0, vx0, ax0, y0, vy0, ay0)
    cbx, cby = h.compute_x_guess(bx, by, params)
    z = jnp.zeros((100, 11))
    pr = prob.projection.compute_projection(xo, yo, bx, by, z, z, cbx, cby, prob.a_obs, prob.b_obs, jnp.zeros((100, 198)))
    acc, steer = h.compute_controls(pr[4], pr[5], pr[6], pr[7])
    state0 = jnp.asarray([x0, y0, vx0, vy0, jnp.arctan2(vy0, vx0)])
    it = 2
    key = jax.random.PRNGKey(3 * idx + 5 * it + 7)
    key, _ = jax.random.split(key)
    rows = list(range(0, 100, 100 // n_chains))[:n_chains]
    npr = prob.num_prime
    g = {"meta.args": np.array([str(a) for a in args]), "meta.variant": np.array(variant), "meta.idx_mpc": np.array(idx), "meta.it": np.array(it),
         "state0": np.asarray(state0, f32), "x_obs_traj": np.asarray(xo, f32), "y_obs_traj": np.asarray(yo, f32), "roll_key": np.asarray(key).astype(np.uint32),
         "acc": np.asarray(acc, f32)[rows], "steer": np.asarray(steer, f32)[rows]}
    outs = {k: [] for k in ("x_red", "y_red", "beta", "sigma", "res_beta", "mmd_obs", "mmd_lane")}
    for b in rows:
        xr, yr, beta, sigma, res = h.compute_rollout_complete_opt(acc[b, 0:npr], steer[b, 0:npr], state0, key)
        outs["x_red"].append(xr); outs["y_red"].append(yr); outs["beta"].append(beta); outs["sigma"].append(sigma); outs["res_beta"].append(res)
        outs["mmd_obs"].append(prob.costs.compute_mmd_obs(beta, sigma, xr, yr, xo[:, 0:npr], yo[:, 0:npr]))
        outs["mmd_lane"].append(prob.costs.compute_mmd_lane(beta, sigma, yr))
    for k, v in outs.items():
        g[k] = np.asarray(np.stack([np.asarray(x) for x in v]), f32)
    np.savez_compressed(os.path.join(HERE, "_part_" + name + ".npz"), **g)


def main():
    if len(sys.argv) > 1:                        # child: one case per process (the two optimizer/ packages share module names)
        kind, i = sys.argv[1], int(sys.argv[2])
        (run_case(*CASES[i]) if kind == "case" else run_opt_case(*OPT_CASES[i]))
        return
    for kind, lst in (("case", CASES), ("opt", OPT_CASES)):
        for i in range(len(lst)):
            print("running", lst[i][0], flush=True)
            subprocess.check_call([sys.executable, os.path.abspath(__file__), kind, str(i)])
    merged = {}
    for lst in (CASES, OPT_CASES):
        for c in lst:
            p = os.path.join(HERE, "_part_" + c[0] + ".npz")
            with np.load(p) as z:
                for k in z.files:
                    merged[c[0] + "/" + k] = z[k]
            os.remove(p)
    if REAL_JAX:
        import jax
        merged["meta.jax_version"] = np.array(jax.__version__)
    np.savez_compressed(os.path.join(HERE, OUT_NAME), **merged)
    print("wrote %s: %d arrays, %.0f KB" % (OUT_NAME, len(merged), os.path.getsize(os.path.join(HERE, OUT_NAME)) / 1024))


if __name__ == "__main__":
    main()
