"""Debug dump of k_project_tc's pass-1 products (MPCMMD_PROJ_DEBUG=1|2) against a float64 NumPy emulation."""
import os, sys
os.environ["MPCMMD_PROJ"] = "tc-always"
dbg = int(sys.argv[1]) if len(sys.argv) > 1 else 1
os.environ["MPCMMD_PROJ_DEBUG"] = str(dbg)
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import cem_impl, constants
f32 = np.float32
args = (5, 2, 0.1, 30, "gaussian", 0.0, 0.0)
n = 200
prob = cem_impl.CEM(*args, variant="static", max_episodes=4)
K = constants.build_constants(30)
rng = np.random.default_rng(5)
params = np.concatenate([rng.uniform(0.1, 30, (n, 4)), rng.normal(0, 6, (n, 4))], 1).astype(f32)
beq_x = np.array([0.0, 5.0, 0.3], f32); beq_y = np.array([1.75, 0.2, -0.1, 0.0], f32)
lam_x = rng.normal(0, 0.5, (n, 11)).astype(f32); lam_y = rng.normal(0, 0.5, (n, 11)).astype(f32)
s_lane = np.abs(rng.normal(0, 1, (n, 198))).astype(f32)
got = prob.stage_project(params, beq_x, beq_y, 15.0, lam_x, lam_y, s_lane)
P, Pd, Pdd = (np.asarray(m, np.float64) for m in (K.P, K.Pdot, K.Pddot))
cbx = np.hstack([params[:, :4], np.tile(beq_x, (n, 1))]).astype(np.float64) @ np.asarray(K.Gx, np.float64).T
cby = np.hstack([params[:, 4:], np.tile(beq_y, (n, 1))]).astype(np.float64) @ np.asarray(K.Gy, np.float64).T
xdg, ydg, xddg, yddg = cbx @ Pd.T, cby @ Pd.T, cbx @ Pdd.T, cby @ Pdd.T
def rep(name, g, r):
    e = np.abs(g - r)
    print(f"{name:8s} max|ref| {np.abs(r).max():10.4g} max err {e.max():10.3e} rel {e.max() / np.abs(r).max():9.2e}  worst col {np.unravel_index(e.argmax(), e.shape)}", flush=True)
if dbg == 1:
    rep("xdg", got["acc"], xdg); rep("ydg", got["steer"], ydg)
else:
    rep("xddg", got["acc"], xddg); rep("yddg", got["steer"], yddg)
vmin, vmax, amax = float(prob.v_min), float(prob.v_max), float(prob.a_max)
def polar(gx, gy, lo, hi):
    al = np.unwrap(np.arctan2(gy, gx), axis=1)
    d = np.clip(gx * np.cos(al) + gy * np.sin(al), lo, hi)
    return d * np.cos(al), d * np.sin(al)
bvx, bvy = polar(xdg, ydg, vmin, vmax); bax, bay = polar(xddg, yddg, 0.0, amax)
Wx = bvx @ Pd + bax @ Pdd
ub, lb = float(prob.y_ub), -float(prob.y_lb)
dl = (ub - s_lane[:, :99].astype(np.float64)) - (lb - s_lane[:, 99:].astype(np.float64))
Wy = bvy @ Pd + bay @ Pdd + dl @ P[1:]
Ux = (xdg - bvx) @ Pd + (xddg - bax) @ Pdd
Uy = (ydg - bvy) @ Pd + (yddg - bay) @ Pdd
rep("Wx", got["cx"], Wx); rep("Wy", got["cy"], Wy); rep("Ux", got["lam_x"], Ux); rep("Uy", got["lam_y"], Uy)
print("Wx got", got["cx"][0]); print("Wx ref", Wx[0])
print("Wy got", got["cy"][0]); print("Wy ref", Wy[0])
print("xdg got", got["acc"][0][:12]); print("xdg ref", (xdg if dbg == 1 else xddg)[0][:12])
