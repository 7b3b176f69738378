"""Sweeps the wavefront kernel's geometry (tile variant, j-split, scatter warps) and prints the kernel time of each
combination next to the automatic choice.  Usage: python tools/tune_sweep.py [n] [kind] [full|auto]
kind: synthetic (config 4 shape, default) or an example shape of workloads.example_shaped (heat, ...)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import mioc_b200 as m  # noqa: E402

wl = importlib.import_module(m.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
kind = sys.argv[2] if len(sys.argv) > 2 else "synthetic"
full = len(sys.argv) > 3 and sys.argv[3] == "full"
inst = wl.synthetic(n=n, B=999, seed=20251018) if kind == "synthetic" else wl.example_shaped(kind, n=n, seed=3)
plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=4)
plan.upload(0, inst.df, inst.u_old)
TILES = ["7+0x2", "8+0x2", "6+0x2", "5+0x2", "4+0x2", "3+0x2", "2+0x2", "1+0x2", "8+0x1", "4+0x1", "2+0x1", "1+0x1",
         "4+3x2", "4+4x2", "3+3x2", "3+2x2", "2+2x2", "2+1x2", "1+1x2", "4+4x1", "2+2x1", "1+1x1", "4+3x1", "3+3x1"]


def run(ctas, js, code):
    try:
        plan.tune(ctas, js, code)
    except m.BellmanB200Error:
        return None
    best = 1e30
    for _ in range(2):
        plan.bellman_resident(0, 1)
        plan.sync()
        best = min(best, plan.stats()["wave_ms"])
    return best, plan.stats()


def show(tag, res):
    best, st = res
    N = plan.count_updates()
    print(f"{tag:10s} tile {TILES[int(st['variant']) - 1]:6s} js={int(st['jsplit'])} ns={int(st['scatter_warps'])} ctas={int(st['ctas'])} "
          f"rows={int(st['rows_per_cta'])} thr={int(st['threads'])}: {best:8.3f} ms  {N / best / 1e9:.3f} T upd/s  "
          f"({best * 1e3 / (n - 1):.2f} us/stage)", flush=True)
    return N / best / 1e9


auto = run(0, 0, 0)
show("AUTO", auto)
if len(sys.argv) > 3 and sys.argv[3] == "auto":
    sys.exit(0)
rows = []
js_list = [1, 2, 3, 4, 6, 8] if full else [1, 2, 4]
ns_list = [10, 1, 2, 4, 6, 8] if full else [10, 2, 4, 6]
for v in range(1, len(TILES) + 1):
    two = "+0" not in TILES[v - 1]
    for js in js_list:
        for ns in ns_list:
            if two and ns == 10:
                continue
            res = run(0, js, 100 * ns + v)
            if res is None:
                continue
            rows.append((show("", res), TILES[v - 1], js, ns % 10))
rows.sort(reverse=True)
print("best:", rows[:8])
print("auto: %.3f T upd/s = %.1f%% of the best" % (plan.count_updates() / auto[0] / 1e9, 100 * plan.count_updates() / auto[0] / 1e9 / rows[0][0]))
