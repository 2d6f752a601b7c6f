#!/usr/bin/env python
"""How the solve graphs scale DOWN with the episode count: the strong-scaled sweep (200 episodes over N GPUs) runs 200/N episodes per
launch, so per-GPU time at E = 100 / 50 / 25 against E = 200 is the single-GPU view of the 2 / 4 / 8-GPU efficiency.
Prints, per (cost, E): graph time (CUDA events, median of 7) and the per-kernel-class times of an ungraphed run."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-mmd_b200")):
    sys.path.insert(1, p)
import torch  # noqa: E402
import __graft_entry__ as G  # noqa: E402

G.build()
from mpcmmd_b200 import CEM, scenes  # noqa: E402

args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
dev = torch.device("cuda", 0)
out = {}
for E in [int(a) for a in (sys.argv[1:] or ["200", "100", "50", "25", "13"])]:
    prob = CEM(*args, variant="static", max_episodes=E, device=0)
    host = scenes.static_batch(prob, list(range(E)), "static")
    dev_in = {k: torch.as_tensor(host[k], device=dev) for k in keys}
    for cost in ("cvar", "mmd_opt"):
        for _ in range(3):
            prob.solve_batch_device(cost, *[dev_in[k] for k in keys])
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); prob.solve_batch_device(cost, *[dev_in[k] for k in keys]); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        prof = prob.profile_solve(cost, E)
        out["%s_E%d" % (cost, E)] = {"graph_ms": float(np.median(ts)), "per_episode_us": 1e3 * float(np.median(ts)) / E, "by_kernel": prof["ms"]}
        print(cost, E, "graph %.3f ms" % np.median(ts), "per-episode %.1f us" % (1e3 * np.median(ts) / E), {k: round(v, 3) for k, v in prof["ms"].items()}, flush=True)
    del prob
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "strong_probe.json"), "w"), indent=1)
