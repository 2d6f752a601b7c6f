"""Generates tests/golden/*.npz.  Run in the build container only (it imports the reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

bernstein_ref.npz : P, Pdot, Pddot (float64, (100,11)) computed by the REFERENCE's own
                    synthetic_static_obs/bernstein_coeff_order10_arbitinterval.py exactly as cem.py:42-46 calls it, plus the
                    same function on the num_prime grids of cem_helper.py:112-118 evaluated in float64 (np.linspace inputs).
                    This is the only piece of the hot path the reference can execute here (JAX is absent).
oracle_solves.npz : outputs of the CPU oracle for a few small solves (regression pins for the oracle itself).
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(1, ROOT)


def main():
    ref = "/root/reference/synthetic_static_obs/bernstein_coeff_order10_arbitinterval.py"
    spec = importlib.util.spec_from_file_location("ref_bernstein", ref)
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    tot_time = np.linspace(0, 15, 100)
    tt = tot_time.reshape(100, 1)
    P, Pd, Pdd = mod.bernstein_coeff_order10_new(10, tt[0], tt[-1], tt)
    out = dict(P=P, Pdot=Pd, Pddot=Pdd)
    for npr in (20, 30, 50, 60, 100):
        tp = np.linspace(0, npr * 0.15, npr).reshape(npr, 1)
        Pp, _, _ = mod.bernstein_coeff_order10_new(10, tp[0], tp[-1], tp)
        out[f"P_prime_{npr}"] = Pp
    np.savez_compressed(os.path.join(HERE, "bernstein_ref.npz"), **out)

    from oracle import oracle as O
    pins = {}
    init_state, mean, cov, v_des = O.driver_inputs("static")
    small = dict(num_batch=24, maxiter_cem=3, num_samples_cem=40, maxiter_beta_cem=4)
    for name, args, cost, variant in (("mmd_opt_g", (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), "mmd_opt", "static"),
                                      ("cvar_b", (5, 4, 0.3, 50, "beta", 0.0, 0.0), "cvar", "static"),
                                      ("saa_g_dyn", (4, 3, 0.1, 20, "gaussian", 0.02, 0.01), "saa", "dynamic"),
                                      ("mmd_random_g", (5, 2, 0.1, 30, "gaussian", 0.0, 0.0), "mmd_random", "static")):
        ora = O.OracleCEM(*args, variant=variant, **small)
        st, mn, cv, vd = O.driver_inputs(variant)
        sc, idx = O.static_episode(args[1], 3)
        xo, yo, _ = ora.compute_obs_trajectories(*sc)
        r = ora.solve(cost, idx, st, mn, cv, xo, yo, vd)
        for k in ("cx", "cy", "cost_obs", "cost_lane", "beta", "sigma"):
            pins[f"{name}.{k}"] = np.asarray(r[k])
    np.savez_compressed(os.path.join(HERE, "oracle_solves.npz"), **pins)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
