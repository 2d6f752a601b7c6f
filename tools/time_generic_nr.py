import sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np, torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
for nr in (6, 8, 10):
    for E in (1, 8):
        prob = CEM(nr, 6, 0.1, 60, "gaussian", 0.0, 0.0, max_episodes=E)
        host = scenes.static_batch(prob, list(range(E)), "static")
        dev_in = {k: torch.as_tensor(host[k], device="cuda:0") for k in keys}
        run = lambda: prob.solve_batch_device("mmd_opt", *[dev_in[k] for k in keys])
        run(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        print(f"nr={nr} E={E}: {np.median(ts)/E:.2f} ms per solve", flush=True)
        del prob
