#!/usr/bin/env python
"""bench.py -- the hot path's headline metric on B200: bellman_TRM! DP cell-updates/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n STAGES]

Workload (BASELINE.json configs[3], SURVEY.md 8d "Config 4"): nu = [[0..4]]x3 via product_iterator (K = 125),
n = 100 000 stages, B = 999, dt = 1, beta = 0.5, p = 1, df ~ N(0,1), u_old piecewise constant with n/10 jumps.
A step = one pass of the hot path over one subproblem per GPU: DP (bellman_TRM!) + selection + backtrack
(eval_u_TRM!).  With N GPUs every rank solves its own subproblem (seed + 2*rank; weak scaling, no data-path
collective) and the ranks finish each step with the best-candidate reduction (one 16-byte record per rank,
all_gather over NCCL).

unit of work: one cell-update = one execution of the reference's innermost loop body (HelpFunctions.jl:71-76);
the exact count N = sum_i K*sum_l max(0, B+1-b~_l(i)) is computed on the device from u_old.

`value`  : cell-updates/s, inputs resident in HBM, CUDA-event timed on the launching stream, max over ranks.
`e2e`    : the same through the host-buffer C-ABI call (bb200_solve: H2D df,u_old -> DP -> backtrack -> D2H u).
`roofline`: FP64 pipe (the binding unit for K >= 5): 2 FP64 ops per cell-update over the wavefront
            kernel's own event-timed duration, against the FP64 issue rate measured live on this GPU by the
            library's DADD microbenchmark (MEASURED_PEAKS.json has no FP64 figure).
`cpu_baseline`: the oracle port (oracle/bellman_oracle.c, -O2 -ffp-contract=off), 1 thread like the reference
            (single-threaded Julia), on a bounded sample of the same workload (n_cpu stages; DP cost is linear in n).
--impl reference: the same oracle port with OpenMP over the independent level loop on all host cores.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "bellman_TRM cell-updates/s"
UNIT = "cell-updates/s"
SEED = 20251018


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=100_000, help="stages of the synthetic instance")
    ap.add_argument("--B", type=int, default=999)
    ap.add_argument("--cpu-n", type=int, default=600, help="stages of the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ctas", type=int, default=0)
    ap.add_argument("--jsplit", type=int, default=0)
    ap.add_argument("--variant", type=int, default=0)
    return ap.parse_args()


def workload_desc(n, B):
    return {"workload": f"synthetic bellman_TRM! instance: nt={n}, 3 integer controls x 5 levels (K=125), "
                        f"B={B} (B+1={B + 1} budget states), dt=1, beta=0.5, p=1, df~N(0,1) seed {SEED}+2*rank",
            "n": n, "K": 125, "M": 3, "B": B}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_leg(o, wl, n_cpu, B, threads, steps=1, warmup=0):
    """Times the oracle port on a bounded sample of the workload; returns (updates/s, seconds per step, N)."""
    inst = wl.synthetic(n=n_cpu, B=B, seed=SEED)
    cost = o.jump_cost_table(inst.beta, inst.p, inst.nu, inst.iterator)
    U, Phi = o.alloc_tables(inst.nu, inst.n, inst.B)   # the reference's own Int64 M-tuple table (multi-trust.jl:71-76)
    u = np.zeros_like(inst.u_old)
    N = 0
    for _ in range(warmup):
        o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi, inst.iterator,
                      cost=cost, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        N = o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi, inst.iterator,
                          cost=cost, threads=threads)
        o.eval_u_TRM(u, inst.u_old, U, Phi, inst.B, inst.nu)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return N / dt, dt, N


def run_reference(args):
    """The reference arm: the CPU implementation of the path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    # torchrun pins OMP_NUM_THREADS=1 for every rank; the other ranks have just exited, so rank 0 takes all host cores
    # (set before the OpenMP runtime is loaded with the oracle library)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or "TORCHELASTIC_RUN_ID" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as o
    o.build()
    m = importlib.import_module("mioc_b200")
    wl = importlib.import_module(m.__name__ + ".workloads")
    cores = o.num_threads(True)
    rate, sec, N = cpu_leg(o, wl, args.cpu_n, args.B, True, steps=max(args.steps, 1), warmup=min(args.warmup, 1))
    cfg = workload_desc(args.n, args.B)
    sample = (f"oracle port (C restatement of HelpFunctions.jl:20-124, gcc -O2 -ffp-contract=off, OpenMP over the level "
              f"loop) on the first {args.cpu_n} of {args.n} stages ({N:.3e} cell-updates/step); the reference itself is "
              f"single-threaded Julia and Julia is not installed")
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import __graft_entry__ as g
    g.build()
    m = importlib.import_module("mioc_b200")
    wl = importlib.import_module(m.__name__ + ".workloads")
    d = importlib.import_module(m.__name__ + ".distributed")

    inst = wl.synthetic(n=args.n, B=args.B, seed=SEED + 2 * rank)
    plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, device=local_rank)
    if args.ctas or args.jsplit or args.variant:
        plan.tune(args.ctas, args.jsplit, args.variant)
    stream = torch.cuda.Stream(device=local_rank)
    plan.set_stream(stream.cuda_stream)   # our kernels run on this torch stream, so torch events see them

    # pinned host buffers (the e2e leg's source and destination)
    h_df = torch.from_numpy(inst.df).pin_memory()
    h_uo = torch.from_numpy(inst.u_old).pin_memory()
    h_u = torch.empty_like(h_uo).pin_memory()
    df_np, uo_np, u_np = h_df.numpy(), h_uo.numpy(), h_u.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        plan.bellman_resident(0, 1)
        plan.backtrack_resident(0, inst.B)

    def reduce_best(phi):
        return d.best_candidate(phi, rank, device=torch.device("cuda", local_rank))

    # ---------------- resident leg: `value` ----------------
    plan.upload(0, df_np, uo_np)
    for _ in range(args.warmup):
        step_resident()
        plan.sync()
        reduce_best(0.0)
    n_upd = plan.count_updates()
    launches0 = plan.stats()["launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wave_ms = []
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step_resident()
            phi, _, _ = plan.download(0, None)        # 32-byte optimum record; syncs the stream
            wave_ms.append(plan.stats()["wave_ms"])
            reduce_best(phi)
        ev1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    launches = plan.stats()["launches"] - launches0
    stats = plan.stats()

    # ---------------- e2e leg: host buffers through the C ABI ----------------
    for _ in range(min(args.warmup, 1)):
        plan.solve(df_np, uo_np, u_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        phi, bs, ks = plan.solve(df_np, uo_np, u_np)
        reduce_best(phi)
    barrier()
    e2e_s = time.perf_counter() - t0

    # max over ranks / sums over ranks
    vals = torch.tensor([ms_total, e2e_s, float(n_upd), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total, e2e_s = mx[0].item(), mx[1].item()
        total_upd, launches = sm[2].item(), sm[3].item()
    else:
        total_upd = float(n_upd)

    if rank == 0:
        value = total_upd * args.steps / (ms_total * 1e-3)
        e2e_value = total_upd * args.steps / e2e_s
        # roofline of the dominant kernel (rank 0's wavefront kernel, event-timed by the library on its stream)
        peak_dadd, _ = m.fp64_peak(local_rank, 0, 400.0)
        peak_mix, _ = m.fp64_peak(local_rank, 1, 400.0)
        kms = statistics.mean(wave_ms) if wave_ms and wave_ms[0] > 0 else float("nan")
        achieved = 2.0 * n_upd / (kms * 1e-3) / 1e12
        io = inst.n * inst.M * 8
        alg_bytes = (inst.n - 1) * (inst.B + 1) * 128 * stats["arg_bytes"] + 2 * io + io
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # DRAM traffic of that kernel per launch: measured once per round with ncu on this very command at full size
        # and committed under profiles/ (bench.py itself never runs under a profiler); scaled by n for other sizes
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "wavefront_traffic_r01.json")))
            traffic = tr["dram_bytes_total"] * (inst.n - 1) / (100_000 - 1)
        except (OSError, KeyError, ValueError):
            pass
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak_dadd / 1e12, "unit": "TFLOP/s",
                    "frac": achieved / (peak_dadd / 1e12), "traffic": traffic,
                    "kernel": "bb200::wavefront_kernel", "kernel_ms": kms,
                    "flops_per_unit": 2, "units_per_launch": n_upd,
                    "peak_source": "live DADD issue-rate microbenchmark (bb200_fp64_peak mode 0) on this GPU; "
                                   "MEASURED_PEAKS.json has no FP64 figure",
                    "peak_relaxation_mix_tflops": peak_mix / 1e12,
                    "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (kms * 1e-3) / 1e9,
                            "peak_gbs": hbm_peak, "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import oracle as o
            rate, sec, N = cpu_leg(o, wl, args.cpu_n, args.B, False)
            cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"oracle port, 1 thread (the reference is single-threaded Julia; Julia not installed), first "
                             f"{args.cpu_n} of {args.n} stages with the reference's Int64 tuple table, {sec:.1f} s, {N:.3e} cell-updates"}
        cfg = workload_desc(args.n, args.B)
        cfg.update({"subproblems_per_gpu": 1, "parallelism": f"independent subproblems x{world}",
                    "l2": "every step streams a (n-1)*(B+1)*128-byte argmin table (12.8 GB at full size) through L2, "
                          "far larger than the 126 MB L2; no separate flush",
                    "kernel_path": int(stats["path"]), "ctas": int(stats["ctas"]), "rows_per_cta": int(stats["rows_per_cta"]),
                    "threads_per_cta": int(stats["threads"]), "jsplit": int(stats["jsplit"])})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * io * world,
                        "d2h_bytes_per_step": (io + 32) * world, "ms_per_step": e2e_s * 1e3 / args.steps},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
