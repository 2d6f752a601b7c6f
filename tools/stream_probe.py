#!/usr/bin/env python
"""Does splitting a batch of episodes over G independent handles / CUDA streams help?  (small kernels of one group overlap the inner-CEM
kernel of another; the tail wave of one group is filled by the other.)  Times cvar + mmd_opt of E episodes as
  serial   : one handle, one stream (the round-1 bench step)
  split G  : G mmd_opt handles with E/G episodes each on G streams + 1 cvar handle on its own stream, all concurrent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mpc-mmd_b200")):
    sys.path.insert(1, p)
import torch  # noqa: E402
import __graft_entry__ as G_  # noqa: E402

G_.build()
from mpcmmd_b200 import CEM, scenes  # noqa: E402

args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
dev = torch.device("cuda", 0)


def timed(fn, n=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


for E in [int(a) for a in (sys.argv[1:] or ["200", "25"])]:
    base = CEM(*args, variant="static", max_episodes=E, device=0)
    host = scenes.static_batch(base, list(range(E)), "static")
    dev_in = {k: torch.as_tensor(host[k], device=dev) for k in keys}

    def serial():
        base.solve_batch_device("cvar", *[dev_in[k] for k in keys])
        base.solve_batch_device("mmd_opt", *[dev_in[k] for k in keys])
    t_serial = timed(serial)
    print("E=%d serial (1 handle, 1 stream): %.3f ms" % (E, t_serial), flush=True)
    for Gn in (1, 2, 3, 4):
        bounds = [round(i * E / Gn) for i in range(Gn + 1)]
        groups = [(bounds[i], bounds[i + 1]) for i in range(Gn) if bounds[i + 1] > bounds[i]]
        hs = [CEM(*args, variant="static", max_episodes=b - a, device=0) for a, b in groups]
        hc = CEM(*args, variant="static", max_episodes=E, device=0)
        ins = [{k: dev_in[k][a:b].contiguous() for k in keys} for a, b in groups]
        streams = [torch.cuda.Stream(dev) for _ in range(len(groups) + 1)]

        def split():
            cur = torch.cuda.current_stream(dev)
            for s in streams:
                s.wait_stream(cur)
            for h, i, s in zip(hs, ins, streams):
                with torch.cuda.stream(s):
                    h.solve_batch_device("mmd_opt", *[i[k] for k in keys])
            with torch.cuda.stream(streams[-1]):
                hc.solve_batch_device("cvar", *[dev_in[k] for k in keys])
            for s in streams:
                cur.wait_stream(s)
        t = timed(split)
        print("E=%d split G=%d (+cvar stream): %.3f ms  (%.3fx serial)" % (E, Gn, t, t_serial / t), flush=True)
        del hs, hc
    del base
