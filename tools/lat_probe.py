"""One-episode (batch = 1) mmd_opt solves in a loop, for ncu / latency experiments."""
import sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes
prob = CEM(5, 4, 0.3, 50, "beta", 0.0, 0.0, variant="static", max_episodes=1)
b = scenes.static_batch(prob, [0])
for _ in range(3):
    out = prob.solve_batch("mmd_opt", **b)
print(prob.profile_solve("mmd_opt", 1))
