"""One small config-4-shaped DP on the production geometry (pruned tiles, two-zone slices, compile-time level count),
checked against the oracle: the command tools/gpu_r2_call31.sh runs under compute-sanitizer."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mioc_b200 as m
from oracle import oracle as o
wl = importlib.import_module(m.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
inst = wl.synthetic(n=n, B=999, seed=3, tie_heavy=True)
plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=4)
plan.bellman(inst.df, inst.u_old)
st = plan.stats()
print("geometry:", {k: st[k] for k in ("path", "variant", "ctas", "rows_per_cta", "ctas_full_rows", "rows_per_cta_top", "threads", "prune_block")})
U, Phi = o.alloc_tables(inst.nu, inst.n, inst.B)
cost = o.jump_cost_table(inst.beta, inst.p, inst.nu, inst.iterator)
o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi, inst.iterator, cost=cost)
ua = np.zeros_like(inst.u_old); ub = np.zeros_like(inst.u_old)
ok = True
for Bn in (999, 300, 0):
    va = plan.eval_u(ua, Bn)
    info = {}
    o.eval_u_TRM(ub, inst.u_old, U, Phi, Bn, inst.nu, info=info)
    same = np.array_equal(ua, ub)
    print("radius", Bn, "value", va, info.get("phi_star"), "trajectory equal", same)
    ok = ok and same
plan.close()
print("OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
