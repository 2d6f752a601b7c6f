"""BASELINE configs[3]: the MMD / CVaR risk stage over num_reduced in {5,10,20,40} x num_prime in {20,40,60,80,100} (20 000 samples = 200 episodes x 100,
6 obstacles, gaussian noise): device time of one `mpcmmd_stage_risk` launch (CUDA events around the C-ABI call, inputs resident), the algorithmic
FLOPs of SURVEY section 8(d) (f_roll + f_cost) and the resulting rate.  Prints a markdown table.  Usage: python tools/bench_cfg4.py [cost]"""
import ctypes as C
import sys
sys.path.insert(1, "/root/repo"); sys.path.insert(1, "/root/repo/mpc-mmd_b200")
import numpy as np, torch
import __graft_entry__ as G
G.build()
from mpcmmd_b200 import CEM, scenes, binding as B
cost = sys.argv[1] if len(sys.argv) > 1 else "mmd_random"
N, O = 20000, 6
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
print("| num_reduced | num_prime | ms per launch | algorithmic GFLOP | TFLOP/s | exps per launch |"); print("|---|---|---|---|---|---|")
for nr in (5, 10, 20, 40):
    for npr in (20, 40, 60, 80, 100):
        prob = CEM(nr, O, 0.1, npr, "gaussian", 0.02, 0.01, max_episodes=N // 100)
        f = dict(device=dev, dtype=torch.float32)
        acc = torch.as_tensor(rng.normal(0, 1.0, (N, 100)).astype(np.float32), device=dev); steer = torch.as_tensor(rng.normal(0, 0.05, (N, 100)).astype(np.float32), device=dev)
        st0 = torch.tensor([0.0, 1.75, 5.0, 0.0, 0.0], **f)
        z = [torch.as_tensor(rng.normal(0, 1, nr * npr).astype(np.float32), device=dev) for _ in range(3)]
        keys = torch.zeros(4, device=dev, dtype=torch.int32)
        sc, _ = scenes.static_scene(O, 3)
        xo, yo, _ = prob.cem_helper.compute_obs_trajectories(*sc)
        xo, yo = torch.as_tensor(xo, device=dev), torch.as_tensor(yo, device=dev)
        out = [torch.empty(N, **f), torch.empty(N, **f), torch.zeros(N, nr, **f), torch.zeros(N, **f), torch.zeros(N, 20, **f)]
        def run():
            B.check(prob._lib.mpcmmd_stage_risk(prob._h, B.COST_KINDS[cost], N, acc.data_ptr(), steer.data_ptr(), st0.data_ptr(), z[0].data_ptr(), z[1].data_ptr(),
                                                z[2].data_ptr(), keys.data_ptr(), xo.data_ptr(), yo.data_ptr(), *[o.data_ptr() for o in out]))
        for _ in range(3): run()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        fl = scenes.flops_per_sample(cost, nr, npr, O)["risk"] * N
        exps = (nr * nr + nr) * N if cost == "mmd_random" else 0
        print(f"| {nr} | {npr} | {ms:.3f} | {fl / 1e9:.2f} | {fl / ms / 1e9:.2f} | {exps:.2e} |", flush=True)
        del prob
