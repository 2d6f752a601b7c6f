#!/usr/bin/env python
"""bench.py -- MPC solves/sec of the B200 trajectory-optimizer path on BASELINE.json's cfg2 workload.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] [--warmup W] # the CPU baseline arm (oracle port, all host threads)

One "step" = the 200-episode static sweep of synthetic_static_obs/main_mpc.py:106-128 solved with BOTH cost
functions of configs[1] (`cvar` and `mmd_opt`; beta noise 0.3, num_obs 4, num_prime 50, num_reduced 5) = 400 solves per GPU.
Multi-GPU is weak scaling: every rank solves its own 200 episodes (episode ids rank*200 ...), no data-path
collective; the only exchange is the final NCCL all_gather of the 26-float per-episode record.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "mpc-mmd_b200")):
    if p not in sys.path:
        sys.path.insert(1, p)

WORK = dict(num_reduced=5, num_obs=4, noise_level=0.3, num_prime=50, noise="beta", acc_const_noise=0.0, steer_const_noise=0.0)
COSTS = ("cvar", "mmd_opt")
EPISODES = 200
VARIANT = "static"
CEM_KW = {}
WORKLOAD_NAME = ("configs[1]: synthetic_static_obs, cvar + mmd_opt, beta noise 0.3, num_obs 4, num_prime 50, num_reduced 5, "
                 "200 episodes per GPU (400 solves/step/GPU)")
METRIC = "MPC solves/sec"

# The default (and the driver's) workload is configs[1].  The other BASELINE configs are parity-test cases (tests/); `--workload` lets
# them be timed with the same harness for the numbers quoted in DESIGN.md / profiles/ -- those lines are not the headline.
WORKLOADS = {
    "cfg2": None,
    "cfg3": dict(work=dict(num_reduced=5, num_obs=6, noise_level=0.1, num_prime=60, noise="gaussian", acc_const_noise=0.0, steer_const_noise=0.0),
                 costs=("mmd_opt",), episodes=200, variant="dynamic", kw={},
                 name="configs[2]: synthetic_dynamic_obs, mmd_opt, gaussian noise 0.1, num_obs 6, num_prime 60, num_reduced 5, 200 episodes per GPU"),
    "cfg3b": dict(work=dict(num_reduced=5, num_obs=6, noise_level=0.3, num_prime=60, noise="beta", acc_const_noise=0.0, steer_const_noise=0.0),
                  costs=("mmd_opt",), episodes=200, variant="dynamic", kw={},
                  name="configs[2]: synthetic_dynamic_obs, mmd_opt, beta noise 0.3, num_obs 6, num_prime 60, num_reduced 5, 200 episodes per GPU"),
    "cfg5": dict(work=dict(num_reduced=5, num_obs=32, noise_level=0.1, num_prime=100, noise="gaussian", acc_const_noise=0.0, steer_const_noise=0.0),
                 costs=("cvar",), episodes=25, variant="static", kw=dict(num_batch=16384),
                 name="configs[4] (scaled synthetic): cvar, 16384 CEM samples x 32 obstacles x num_prime 100, 25 episodes per GPU per step "
                      "(the 200-per-GPU sweep is 8 such steps)"),
}


def select_workload(name):
    global WORK, COSTS, EPISODES, VARIANT, CEM_KW, WORKLOAD_NAME
    w = WORKLOADS[name]
    if w is None:
        return
    WORK, COSTS, EPISODES, VARIANT, CEM_KW, WORKLOAD_NAME = w["work"], w["costs"], w["episodes"], w["variant"], w["kw"], w["name"]


def cem_args():
    w = WORK
    return (w["num_reduced"], w["num_obs"], w["noise_level"], w["num_prime"], w["noise"], w["acc_const_noise"], w["steer_const_noise"])


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself needs jax==0.3.23, absent from this image), naive formulation
def cpu_arm_step(ora, O, episodes):
    """solve `episodes` with both costs on the host; returns number of solves"""
    from mpcmmd_b200 import scenes            # host-only scene generators (NumPy), shared with the CUDA arm so both solve the same inputs
    init_state, mean, cov, v_des = scenes.driver_inputs(VARIANT)
    n = 0
    for k in episodes:
        if VARIANT == "dynamic":
            _, idx, xo, yo = scenes.dynamic_scene(WORK["num_obs"], k)
        else:
            sc, idx = scenes.static_scene(WORK["num_obs"], k) if WORK["num_obs"] <= 9 else scenes.scaled_scene(WORK["num_obs"], k)
            xo, yo, _ = ora.compute_obs_trajectories(*sc)
        for cost in COSTS:
            ora.solve(cost, idx, init_state, mean, cov, xo, yo, v_des)
            n += 1
    return n


def make_cpu_arm():
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    ora = O.OracleCEM(*cem_args(), variant=VARIANT, naive=True, **CEM_KW)
    return ora, O, cores


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ora, O, cores = make_cpu_arm()
    REF_EPISODES = [0, 1, 2, 3]          # bounded sample of the sweep per step (about 3.5 s of CPU work per step on 16 threads)
    sample = "episodes 0-3 of the sweep, %s (%d solves per step), oracle port in the reference's naive formulation, %d host threads" % (
        " + ".join(COSTS), len(COSTS) * len(REF_EPISODES), cores)
    for _ in range(args.warmup):
        cpu_arm_step(ora, O, REF_EPISODES)
    t0 = time.perf_counter()
    n = 0
    for _ in range(args.steps):
        n += cpu_arm_step(ora, O, REF_EPISODES)
    dt = time.perf_counter() - t0
    val = n / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD_NAME, "sample_per_step": "%d episodes x %d costs" % (len(REF_EPISODES), len(COSTS))},
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self._stop, self._th = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._th = threading.Thread(target=self._run, daemon=True); self._th.start(); return self

    def __exit__(self, *a):
        self._stop.set(); self._th.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def main():
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout), so everything
    # except the final line is routed to stderr: fd 1 is pointed at fd 2 for the run and the line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--projection", default="exact", choices=["exact", "tc"],
                    help="exact = k_project (bit-exact FP32, the default product path); tc = k_project_tc (tcgen05 kind::tf32 products, 1e-4 stage parity)")
    args = ap.parse_args()
    if args.projection == "tc":
        os.environ["MPCMMD_PROJ"] = "tc"          # read by mpcmmd_create
    else:
        os.environ.pop("MPCMMD_PROJ", None)
    select_workload(args.workload)
    if args.impl == "reference":
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as G
    if local == 0:
        G.build()                      # one builder per node; the other ranks wait, then only load
    if world > 1:
        dist.barrier()
    if local != 0:
        G.build()
    from mpcmmd_b200 import CEM, cem_impl, scenes
    W, K, E = max(args.warmup, 3), args.steps, EPISODES

    prob = CEM(*cem_args(), variant=VARIANT, max_episodes=E, device=local, **CEM_KW)
    eps = list(range(rank * E, rank * E + E))
    host = scenes.static_batch(prob, eps, VARIANT)
    keys = ("idx_mpc", "init_state", "mean_param", "cov_param", "x_obs_traj", "y_obs_traj", "v_des")
    dev_in = {k: torch.as_tensor(host[k], device=dev) for k in keys}
    pinned = {k: torch.as_tensor(host[k]).pin_memory() for k in keys}
    pinned_np = {k: pinned[k].numpy() for k in keys}
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2
    thresholds = {"cvar": 1e-5, "mmd_opt": -prob.ker_wt + 1.0}                    # main_mpc.py:88-97
    HEAVY = COSTS[-1]                                                             # the cost whose risk stage the roofline describes

    def record(out, cost):
        """26-float per-episode record [k, accepted, cost_obs, cost_lane, cx(11), cy(11)] gathered over ranks"""
        rec = torch.cat([torch.as_tensor(eps, device=dev, dtype=torch.float32)[:, None], (out["cost_obs"] <= thresholds[cost]).float()[:, None],
                         out["cost_obs"][:, None], out["cost_lane"][:, None], out["cx"], out["cy"]], 1)
        if world > 1:
            buf = [torch.empty_like(rec) for _ in range(world)]
            dist.all_gather(buf, rec)
            rec = torch.cat(buf, 0)
        return rec

    def step_device():
        recs = {}
        for cost in COSTS:
            out = prob.solve_batch_device(cost, *[dev_in[k] for k in keys])
            recs[cost] = record(out, cost)
        return recs

    def step_host():
        outs = {}
        for cost in COSTS:
            outs[cost] = prob.solve_batch(cost, *[pinned_np[k] for k in keys])
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`)
    for _ in range(W):
        step_device()
    launches_per_step = 0
    for cost in COSTS:            # graphs are cached now; count launches per solve batch
        prob.solve_batch_device(cost, *[dev_in[k] for k in keys]); launches_per_step += prob.last_launch_count()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(K):
            flush.fill_(i & 0xFF)                       # L2 flush between timed steps (outside the event pair)
            ev[i][0].record()
            recs = step_device()
            ev[i][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    ms = sum(a.elapsed_time(b) for a, b in ev)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    solves_per_step = E * len(COSTS) * world
    value = solves_per_step * K / (ms_total * 1e-3)
    accepted = {c: int(recs[c][:, 1].sum().item()) for c in COSTS}

    # ---- end to end through the host-buffer C ABI call (`e2e`): pinned host inputs, H2D + D2H inside the timed region
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        outs = step_host()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = solves_per_step * K / float(e2e_t.item())
    h2d = len(COSTS) * sum(int(pinned_np[k].nbytes) for k in keys)
    d2h = len(COSTS) * sum(int(v.nbytes) for v in outs[HEAVY].values())
    same = all(np.array_equal(outs[c]["cx"], recs[c][rank * E:(rank + 1) * E, 4:15].cpu().numpy()) for c in COSTS)

    # ---- roofline of the dominant kernel (k_risk_opt: rollouts + reduced-set CEM + risk), launch-by-launch CUDA events
    prob.solve_batch_device(HEAVY, *[dev_in[k] for k in keys]); torch.cuda.synchronize()
    prof = prob.profile_solve(HEAVY, E)
    prof_cvar = None
    if "cvar" in COSTS and HEAVY != "cvar":
        prob.solve_batch_device("cvar", *[dev_in[k] for k in keys]); torch.cuda.synchronize()
        prof_cvar = prob.profile_solve("cvar", E)
    fl = scenes.flops_per_sample(HEAVY, WORK["num_reduced"], WORK["num_prime"], WORK["num_obs"])
    flops_per_launch = fl["risk"] * prob.num_batch * E
    avg_launch_s = prof["ms"]["risk"] * 1e-3 / prof["launches"]["risk"]
    peak_tf, sm_count = cem_impl.fp32_peak(local)
    achieved = flops_per_launch / avg_launch_s / 1e12
    # dram bytes per risk-stage launch from the committed ncu --set full captures (profiles/r01_v12_summary.md, r01_v9_summary.md):
    # k_inner_cem_fast 45 + 12 MB, k_rollouts<ROLL_OPT> 19 + 28 MB at the cfg2 shape (k_opt_risk not captured, < 5 MB of controls); null for the
    # other workloads
    traffic = 104.0e6 if (args.workload == "cfg2") else None
    roofline = {"bound": "fp32", "kernel": ("k_rollouts + k_inner_cem_fast<5> + k_opt_risk (mother rollouts, reduced-set inner CEM, MMD risk)" if HEAVY == "mmd_opt"
                                            else "k_rollouts (noisy rollouts + %s risk)" % HEAVY), "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "peak_source": "measured in this run: register-resident FP32 fma micro-kernel (MEASURED_PEAKS.json has no FP32 figure)",
                "flops_per_launch": flops_per_launch, "avg_launch_ms": avg_launch_s * 1e3, "sm_count": sm_count,
                "kernel_share_of_%s_solve" % HEAVY: prof["ms"]["risk"] / prof["ms"]["total"],
                "%s_ms_by_kernel" % HEAVY: prof["ms"]}
    if prof_cvar:
        roofline["cvar_ms_by_kernel"] = prof_cvar["ms"]

    # ---- per-solve latency at batch = 1 (BASELINE.json's second headline)
    lat = {}
    if rank == 0:
        p1 = CEM(*cem_args(), variant=VARIANT, max_episodes=1, device=local, **CEM_KW)
        one = {k: dev_in[k][:1].contiguous() for k in keys}
        for cost in COSTS:
            for _ in range(5):
                p1.solve_batch_device(cost, *[one[k] for k in keys])
            torch.cuda.synchronize()
            ts = []
            for _ in range(30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); p1.solve_batch_device(cost, *[one[k] for k in keys]); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            # the same solve through the reference-facing call (CEM.compute_cem_*: host arrays in, host arrays out, wall clock)
            fn = getattr(p1, "compute_cem_" + cost)
            h1 = [int(host["idx_mpc"][0]), host["init_state"][0], host["mean_param"][0], host["cov_param"][0], host["x_obs_traj"][0], host["y_obs_traj"][0],
                  float(host["v_des"][0])]
            for _ in range(5):
                fn(*h1)
            tw = []
            for _ in range(30):
                t0 = time.perf_counter(); fn(*h1); tw.append(1e3 * (time.perf_counter() - t0))
            lat[cost] = {"p50_ms": float(np.percentile(ts, 50)), "p95_ms": float(np.percentile(ts, 95)),
                         "p50_ms_host_api": float(np.percentile(tw, 50)), "p95_ms_host_api": float(np.percentile(tw, 95)),
                         "ms_by_kernel_ungraphed": p1.profile_solve(cost, 1)["ms"]}
        del p1

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ora, O, cores = make_cpu_arm()
        t0 = time.perf_counter()
        n = cpu_arm_step(ora, O, list(range(12)))
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt, "unit": "solves/s", "cores": cores, "kind": "port",
               "sample": "episodes 0-11 of the sweep x {%s} = %d solves in %.1f s; oracle port (C, reference's naive formulation), %d host threads" % (", ".join(COSTS), n, dt, cores)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_NAME, "episodes_per_gpu": E, "costs": list(COSTS), "num_batch": prob.num_batch,
                           "maxiter_cem": prob.maxiter_cem, "projection": args.projection, "l2": "flushed between timed steps (256 MiB fill)", "parallelism": "episodes sharded %d-way" % world,
                           "accepted": accepted, "wall_s_timed_region": t_wall},
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "matches_device_path": bool(same)},
                "gpu_launches": launches_per_step * K, "roofline": roofline, "cpu_baseline": cpu, "latency_1gpu_batch1": lat, "clocks": clk.summary()}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
