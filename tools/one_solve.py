#!/usr/bin/env python
"""One mmd_opt (or cvar) solve batch of E episodes of configs[1] -- the target of ncu captures.  usage: one_solve.py [cost] [E] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-mmd_b200")):
    sys.path.insert(1, p)
import torch  # noqa: E402
import __graft_entry__ as G  # noqa: E402

G.build()
from mpcmmd_b200 import CEM, scenes  # noqa: E402

cost = sys.argv[1] if len(sys.argv) > 1 else "mmd_opt"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 200
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
iters = int(os.environ.get("MAXITER_CEM", "20"))
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
prob = CEM(5, 4, 0.3, 50, "beta", 0.0, 0.0, variant="static", max_episodes=E, device=0, maxiter_cem=iters)
host = scenes.static_batch(prob, list(range(E)), "static")
dev_in = {k: torch.as_tensor(host[k], device="cuda:0") for k in keys}
for _ in range(reps):
    prob.solve_batch_device(cost, *[dev_in[k] for k in keys])
torch.cuda.synchronize()
print("done", cost, E, prob.last_launch_count())
