"""Iteration-0 batch of a real solve through both projection kernels: which 20 rows have the lowest res_norm (cem.py:233)?"""
import os, sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import cem_impl
f32 = np.float32
args = (5, 4, 0.3, 50, "beta", 0.0, 0.0)
out = {}
for tag, env in (("fp32", ""), ("tc", "tc-always")):
    os.environ["MPCMMD_PROJ"] = env
    prob = cem_impl.CEM(*args, variant="static", max_episodes=2)
    zi = prob.tables()[0].reshape(100, 8)
    L = np.sqrt(np.array([20.0] * 4 + [100.0] * 4, f32))
    params = (np.array([15.0] * 4 + [0.0] * 4, f32) + zi * L).astype(f32)
    params[:, :4] = np.clip(params[:, :4], 0.1, 30.0)
    beq_x = np.array([0.0, 5.0, 0.0], f32); beq_y = np.array([1.75, 0.0, 0.0, 0.0], f32)
    out[tag] = prob.stage_project(params, beq_x, beq_y, 15.0, np.zeros((100, 11), f32), np.zeros((100, 11), f32), np.zeros((100, 198), f32))
a, b = out["fp32"], out["tc"]
ra, rb = a["res_norm"], b["res_norm"]
print("res_norm fp32 sorted[:25]", np.sort(ra)[:25])
print("res_norm tc   sorted[:25]", np.sort(rb)[:25])
ka, kb = set(np.argsort(ra, kind="stable")[:20].tolist()), set(np.argsort(rb, kind="stable")[:20].tolist())
print("rows kept by both:", len(ka & kb), "of 20;  feasible-looking rows (res_norm < 1e-3): fp32", int((ra < 1e-3).sum()), "tc", int((rb < 1e-3).sum()))
for k in ("cx", "cy", "acc", "steer", "cost_base", "res_norm"):
    e = np.abs(a[k].astype(np.float64) - b[k]); print(f"{k:9s} max|fp32| {np.abs(a[k]).max():9.4g} max|tc - fp32| {e.max():9.3e}")
