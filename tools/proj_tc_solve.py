"""Full cvar solves with the tensor-core projection (MPCMMD_PROJ=tc) next to the exact FP32 path: per-episode results side by side."""
import os, sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import cem_impl
from oracle import oracle as O
args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 6
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ora = O.OracleCEM(*args, variant="static")
init_state, mean, cov, v_des = O.driver_inputs("static")
eps = [O.static_episode(4, k) for k in range(E)]
tr = [ora.compute_obs_trajectories(*sc) for sc, _ in eps]
idx = [i for _, i in eps]; xo = np.stack([t[0] for t in tr]); yo = np.stack([t[1] for t in tr])
res = {}
for tag, env in (("fp32", ""), ("tc", "tc-always")):
    os.environ["MPCMMD_PROJ"] = env
    prob = cem_impl.CEM(*args, variant="static", max_episodes=E, maxiter_cem=IT)
    res[tag] = prob.solve_batch(os.environ.get("PROBE_COST", "cvar"), idx, np.stack([init_state] * E), np.stack([mean] * E), np.stack([cov] * E), xo, yo, [v_des] * E)
for e in range(min(E, 6)):
    for tag in ("fp32", "tc"):
        r = res[tag]
        print(e, tag, "cx", r["cx"][e][:5], "cy", r["cy"][e][:4], "obs", r["cost_obs"][e], "lane", r["cost_lane"][e], flush=True)
d = np.abs(res["tc"]["cx"] - res["fp32"]["cx"]).max(axis=1)
print("episodes with |cx_tc - cx_fp32| < 1e-2:", int((d < 1e-2).sum()), "of", E, " nan:", int(np.isnan(res["tc"]["cx"]).sum()))
