"""Per-call latency of one TR inner iteration (bb200_solve: H2D -> DP -> selection -> backtrack -> D2H) on the table
shapes of the reference's examples (BASELINE configs 1-3), next to the CPU oracle port.  No roofline claim: these
shapes are launch/latency bound.  Usage: python tools/example_latency.py"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mioc_b200 as m
from oracle import oracle as o
wl = importlib.import_module(m.__name__ + ".workloads")

cases = [("fishing", 1024), ("vanderpol", 1024), ("doubletank", 1024), ("convolution", 1024), ("heat", 1024),
         ("heat", 8192), ("heat", 16384)]
print(f"{'shape':22s} {'K':>4s} {'B':>5s} {'updates':>10s} | {'gpu solve ms':>12s} {'ctas':>5s} {'graph':>6s} | {'stage-kernel ms':>15s} | {'cpu port ms':>11s} {'speedup':>8s}")
for kind, n in cases:
    inst = wl.example_shaped(kind, n=n, seed=3)
    u = np.zeros_like(inst.u_old)
    row = []
    for flags in (0, 1):
        plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=flags)
        for _ in range(3):
            plan.solve(inst.df, inst.u_old, u)
        reps = 20 if n <= 1024 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            plan.solve(inst.df, inst.u_old, u)
        ms = (time.perf_counter() - t0) / reps * 1e3
        st = plan.stats()
        row.append((ms, int(st["ctas"]), int(st["graph_replays"]), plan.count_updates()))
        plan.close()
    cpu_ms = float("nan")
    if n <= 8192:
        Phi = o.alloc_tables(inst.nu, inst.n, inst.B)[1]
        cost = o.jump_cost_table(inst.beta, inst.p, inst.nu, inst.iterator)
        t0 = time.perf_counter()
        o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, None, Phi, inst.iterator, cost=cost)
        cpu_ms = (time.perf_counter() - t0) * 1e3
    print(f"{kind + ' n=' + str(n):22s} {inst.K:4d} {inst.B:5d} {row[0][3]:10.3e} | {row[0][0]:12.3f} {row[0][1]:5d} {row[0][2]:6d} | "
          f"{row[1][0]:15.3f} | {cpu_ms:11.2f} {cpu_ms / row[0][0]:8.1f}", flush=True)
