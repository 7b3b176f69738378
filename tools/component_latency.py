import importlib, os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import mioc_b200 as m
wl = importlib.import_module(m.__name__ + ".workloads")
for kind, n in (("fishing", 1024), ("heat", 1024)):
    inst = wl.example_shaped(kind, n=n, seed=3)
    plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
    u = np.zeros_like(inst.u_old)
    plan.upload(0, inst.df, inst.u_old)
    for _ in range(3):
        plan.bellman_resident(0, 1); plan.sync()
    t0 = time.perf_counter()
    for _ in range(20):
        plan.bellman_resident(0, 1); plan.sync()
    t_dp = (time.perf_counter() - t0) / 20 * 1e3
    st = plan.stats()
    t0 = time.perf_counter()
    for _ in range(20):
        plan.backtrack_resident(0, inst.B); plan.sync()
    t_bt = (time.perf_counter() - t0) / 20 * 1e3
    st2 = plan.stats()
    t0 = time.perf_counter()
    for _ in range(20):
        plan.solve(inst.df, inst.u_old, u)
    t_solve = (time.perf_counter() - t0) / 20 * 1e3
    print(kind, n, f"dp host {t_dp:.3f} ms (events: all {st['dp_ms']:.3f}, main kernel {st['wave_ms']:.3f}) | backtrack host {t_bt:.3f} ms (events {st2['backtrack_ms']:.3f}) | solve {t_solve:.3f} ms path={int(st['path'])}")
