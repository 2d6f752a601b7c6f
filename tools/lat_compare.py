"""p50 latency of one mmd_opt solve batch of E episodes under different inner-CEM kernels (MPCMMD_INNER_CEM read at handle creation)."""
import os, sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np, torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
for E in (1, 2, 4):
    for mode in ("lat", "lat512", "cta"):
        os.environ["MPCMMD_INNER_CEM"] = mode
        prob = CEM(5, 4, 0.3, 50, "beta", 0.0, 0.0, variant="static", max_episodes=E)
        host = scenes.static_batch(prob, list(range(E)), "static")
        dev_in = {k: torch.as_tensor(host[k], device="cuda:0") for k in keys}
        for _ in range(5):
            prob.solve_batch_device("mmd_opt", *[dev_in[k] for k in keys])
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); prob.solve_batch_device("mmd_opt", *[dev_in[k] for k in keys]); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"E={E} {mode}: p50 {np.percentile(ts, 50):.3f} ms  by kernel {prob.profile_solve('mmd_opt', E)['ms']}", flush=True)
        del prob
