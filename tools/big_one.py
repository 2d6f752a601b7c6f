"""one mmd_opt risk stage for a large reduced set (ncu target): usage big_one.py nr n_chains iters_in"""
import sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np, torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes
nr, n, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prob = CEM(nr, 2, 0.1, 20, "gaussian", 0.0, 0.0, max_episodes=1, num_batch=n, maxiter_cem=1, maxiter_beta_cem=iters)
host = scenes.static_batch(prob, [0], "static")
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
dev_in = {k: torch.as_tensor(host[k], device="cuda:0") for k in keys}
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); prob.solve_batch_device("mmd_opt", *[dev_in[k] for k in keys]); b.record(); torch.cuda.synchronize()
    print("nr", nr, "chains", n, "inner iterations", iters, "ms", a.elapsed_time(b), flush=True)
