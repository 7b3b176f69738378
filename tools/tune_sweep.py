"""Sweeps the wavefront kernel's geometry (tile variant, j-split, CTA count) on config-4-shaped input and
prints the kernel time and cell-updates/s of each combination.  Usage: python tools/tune_sweep.py [n] [quick]"""
import importlib
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mioc_b200 as m  # noqa: E402

wl = importlib.import_module(m.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
inst = wl.synthetic(n=n, B=999, seed=20251018)
plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
plan.upload(0, inst.df, inst.u_old)
rows = []
variants = {1: "7x4", 2: "8x4", 3: "4x4", 4: "7x2", 5: "8x2", 6: "8x1", 7: "4x1"}
combos = list(itertools.product([1, 2, 3, 4, 5, 6], [2, 4, 6, 7, 8], [0]))
for v, js, ctas in combos:
    try:
        plan.tune(ctas, js, v)
    except m.BellmanB200Error:
        continue
    best = 1e30
    for _ in range(2):
        plan.bellman_resident(0, 1)
        plan.sync()
        best = min(best, plan.stats()["wave_ms"])
    st = plan.stats()
    N = plan.count_updates()
    rows.append((N / (best * 1e-3) / 1e12, variants[v], js, int(st["ctas"]), int(st["rows_per_cta"]), int(st["threads"]), best))
    print(f"variant {variants[v]} js={js} ctas={int(st['ctas'])} rows={int(st['rows_per_cta'])} thr={int(st['threads'])}: "
          f"{best:8.3f} ms  {rows[-1][0]:.3f} T upd/s  ({best * 1e3 / (n - 1):.2f} us/stage)", flush=True)
rows.sort(reverse=True)
print("best:", rows[:5])
