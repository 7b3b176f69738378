"""Throughput of the DP on level sets the pipelined kernel does not take (jump-cost table larger than shared memory,
K > 150; uint16 argmin, K > 255): these shapes run one launch per stage -- with the pruned scan (kernel_stage_pruned.cu)
unless the plan was created with BB200_FLAG_STAGE_KERNELS.  With K^2 candidates per cell a stage is long enough for the
launch overhead not to matter.  Usage: python tools/wide_k_probe.py [plain]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mioc_b200 as m
wl = importlib.import_module(m.__name__ + ".workloads")

print(f"{'shape':28s} {'K':>4s} {'path':>5s} {'arg bytes':>9s} {'updates':>10s} | {'DP ms':>9s} {'us/stage':>9s} {'T upd/s':>8s}")
for levels, M, n in ((5, 3, 2000), (6, 3, 2000), (7, 3, 1000), (16, 2, 1000)):
    inst = wl.synthetic(n=n, B=999, seed=11, levels=levels, M=M)
    plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=(1 if 'plain' in sys.argv and inst.K > 150 else 0))
    plan.upload(0, inst.df, inst.u_old)
    for _ in range(2):
        plan.bellman_resident(0, 1); plan.sync()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        plan.bellman_resident(0, 1)
    plan.sync()
    ms = (time.perf_counter() - t0) / reps * 1e3
    st = plan.stats()
    N = plan.count_updates()
    print(f"{'synthetic ' + str(levels) + '^' + str(M) + ' n=' + str(n):28s} {inst.K:4d} {int(st['path']):5d} {int(st['arg_bytes']):9d} {N:10.3e} | {ms:9.2f} "
          f"{ms * 1e3 / (n - 1):9.2f} {N / ms / 1e9:8.3f}", flush=True)
    plan.close()
