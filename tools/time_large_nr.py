"""Times mmd_opt solves with reduced sets larger than 5 (generic inner-CEM kernel).  Usage: python tools/time_large_nr.py [nr ...]"""
import sys, time
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np, torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes
nrs = [int(a) for a in sys.argv[1:]] or [10, 8, 6]
for nr in nrs:
    E = 8
    prob = CEM(nr, 6, 0.1, 60, "gaussian", 0.0, 0.0, variant="dynamic", max_episodes=E)
    b = scenes.static_batch(prob, list(range(E)), "dynamic")
    prob.solve_batch("mmd_opt", **b)
    t = time.perf_counter(); out = prob.solve_batch("mmd_opt", **b); dt = time.perf_counter() - t
    print("nr", nr, "E", E, "ms/solve", round(1e3 * dt / E, 2), prob.profile_solve("mmd_opt", E)["ms"], flush=True)
