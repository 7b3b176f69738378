"""Compares the pruned scan's executed-candidate count with a numpy emulation of the same decisions (debug aid)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mioc_b200 as m
from oracle import oracle as o
wl = importlib.import_module(m.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 25
blk = 8 if variant == 26 else 4
inst = wl.synthetic(n=n, B=999, seed=20251018)
plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=4)
plan.tune(variant=variant)
plan.bellman(inst.df, inst.u_old)
st = plan.stats()
N = plan.count_updates()
print(f"kernel: executed {st['executed_updates']:.4g} of N {N:.4g} = {st['executed_updates'] / N:.3f}  (prune block {st['prune_block']})")
# numpy emulation of the whole DP with the kernel's decisions
K = len(inst.iterator); B1 = inst.B + 1
lv = o.level_values(inst.nu, inst.iterator).astype(float)
cost = plan.cost
nbr = (K + blk - 1) // blk
cminq = np.stack([cost[q * blk:min(K, (q + 1) * blk)].min(axis=0) for q in range(nbr)])
def stage_cost(i):
    s = np.zeros(K)
    for mm in range(3):
        s = s + (inst.dt * inst.df[i, mm]) * lv[:, mm]
    bt = np.abs(lv - inst.u_old[i][None, :]).sum(1).astype(int)
    return s, bt
s, bt = stage_cost(n - 1)
P = np.full((B1, K), np.inf)
for l in range(K):
    if bt[l] < B1:
        P[bt[l], l] = s[l]
ex_total = 0.0; full_total = 0.0
for i in range(n - 2, -1, -1):
    s, bt = stage_cost(i)
    v = (s[None, :, None] + cost.T[None, :, :]) + P[:, None, :]
    best = v.min(axis=2)
    pminq = np.stack([np.fmin.reduce(P[:, q * blk:min(K, (q + 1) * blk)], axis=1) for q in range(nbr)])
    vblk = np.stack([np.fmin.reduce(v[:, :, q * blk:min(K, (q + 1) * blk)], axis=2) for q in range(nbr)])
    LB = (s[None, None, :] + cminq[:, None, :]) + pminq[:, :, None]
    ex = 0; tot = 0
    for b0 in range(0, B1, 7):
        for (ra, rb) in ((0, 4), (4, 7)):
            rows = slice(b0 + ra, min(B1, b0 + rb))
            if rows.start >= B1:
                continue
            seed = int(np.argmin(pminq[:, rows].min(axis=1)))
            for l0 in range(0, K, 32):
                ls = slice(l0, min(K, l0 + 32))
                UB = vblk[seed][rows, ls]
                need = ~(LB[:, rows, ls] > UB[None])
                ex += need.any(axis=(1, 2)).sum() * (rb - ra) * (ls.stop - ls.start) * blk
                tot += nbr * (rb - ra) * (ls.stop - ls.start) * blk
    ex_total += ex; full_total += tot
    if i % 20 == 0:
        print(f"  stage {i}: emulated executed fraction {ex / tot:.3f}, finite P {np.isfinite(P).mean():.3f}")
    Pn = np.full_like(P, np.inf)
    for l in range(K):
        if bt[l] < B1:
            Pn[bt[l]:, l] = best[:B1 - bt[l], l]
    P = Pn
print(f"emulation: executed {ex_total:.4g} of full {full_total:.4g} = {ex_total / full_total:.3f}; vs N: {ex_total / N:.3f}")
