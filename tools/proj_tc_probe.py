"""k_project_tc (MPCMMD_PROJ=tc) against the CPU oracle on seeded inputs: per-output error table and timing."""
import os, sys, time
os.environ["MPCMMD_PROJ"] = "tc-always"
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np
import torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import cem_impl
from oracle import oracle as O
f32 = np.float32
args = (5, 2, 0.1, 30, "gaussian", 0.0, 0.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
prob = cem_impl.CEM(*args, variant="static", max_episodes=max(4, (n + 99) // 100))
ora = O.OracleCEM(*args, variant="static")
rng = np.random.default_rng(5)
hard = os.environ.get("PROBE_HARD", "0") == "1"
params = np.concatenate([rng.uniform(0.1, 45 if hard else 30, (n, 4)), rng.normal(0, 12 if hard else 6, (n, 4))], 1).astype(f32)
beq_x = np.array([0.0, 5.0, 0.3], f32); beq_y = np.array([1.75, 0.2, -0.1, 0.0], f32)
lam_x = rng.normal(0, 0.5, (n, 11)).astype(f32); lam_y = rng.normal(0, 0.5, (n, 11)).astype(f32)
s_lane = np.abs(rng.normal(0, 1, (n, 198))).astype(f32)
lam_x[:10] = 0; lam_y[:10] = 0; s_lane[:10] = 0
got = prob.stage_project(params, beq_x, beq_y, 15.0, lam_x, lam_y, s_lane)
ref = {k: [] for k in got}
for i in range(min(n, 300)):
    lx, ly, sl = lam_x[i].copy(), lam_y[i].copy(), s_lane[i].copy()
    r = ora.project(params[i], beq_x, beq_y, 15.0, lx, ly, sl)
    for k in ("cx", "cy", "res_norm", "acc", "steer", "cost_base"):
        ref[k].append(np.asarray(r[k]))
    ref["lam_x"].append(lx); ref["lam_y"].append(ly); ref["s_lane"].append(sl)
m = min(n, 300)
for k in got:
    R = np.asarray(ref[k], dtype=np.float64); Gt = got[k][:m].astype(np.float64)
    err = np.abs(Gt - R)
    print(f"{k:10s} max|ref| {np.abs(R).max():10.4g}  max abs err {err.max():10.3e}  rel-to-max {err.max() / max(np.abs(R).max(), 1e-30):9.2e}  nan {int(np.isnan(Gt).sum())}", flush=True)
print("sample0 cx got", got["cx"][0][:4], "ref", ref["cx"][0][:4])

import torch
for tag, env in (("tc", "tc-always"), ("fp32", "")):
    os.environ["MPCMMD_PROJ"] = env
    pr = cem_impl.CEM(*args, variant="static", max_episodes=200)
    nn = 20000
    pp = np.tile(params, (nn // n + 1, 1))[:nn]; lx = np.zeros((nn, 11), f32); sl = np.zeros((nn, 198), f32)
    tp, tx, ty, tlx, tly, tsl = (pr._t(v) for v in (pp, beq_x, beq_y, lx, lx.copy(), sl))
    f = dict(device=pr.device, dtype=torch.float32)
    o = [torch.empty(nn, 11, **f), torch.empty(nn, 11, **f), torch.empty(nn, **f), torch.empty(nn, 100, **f), torch.empty(nn, 100, **f), torch.empty(nn, **f)]
    def call():
        pr._lib.mpcmmd_stage_project(pr._h, nn, tp.data_ptr(), tx.data_ptr(), ty.data_ptr(), 15.0, tlx.data_ptr(), tly.data_ptr(), tsl.data_ptr(),
                                     o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), o[3].data_ptr(), o[4].data_ptr(), o[5].data_ptr())
    for _ in range(3): call()
    t0 = time.perf_counter()
    for _ in range(20): call()
    print(tag, "n=20000 wall per call (kernel + memcpy + sync): %.1f us" % ((time.perf_counter() - t0) / 20 * 1e6), flush=True)
