"""Offline statistics for the design of the pruned scan's bound tests (CPU only, numpy).

Runs the config-4-shaped DP stage by stage and counts, per warp tile (rows_per_warp rows x 32 levels, blocks of 4
successors), how many blocks survive
  exact   : the exact test  LB[q][r][l] > UB[r][l]  (union over the tile's cells) -- what has to be scanned
  super   : blocks inside super-blocks (16 successors) that survive the coarse test merged over the tile's rows (round-2 kernel)
  rowlane : blocks that survive a per-row test merged over the tile's 32 levels:
            Cw[q] + pm[q][r] > max_l (UB[r][l] - s_l)      with Cw[q] = min over the tile's levels of cmin[q][l]
usage: python tools/prune_stats.py [n] [rows_per_warp]
"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mioc_b200 as m
from oracle import oracle as o
wl = importlib.import_module(m.__name__ + ".workloads")

n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
RW = int(sys.argv[2]) if len(sys.argv) > 2 else 2
blk = 4
inst = wl.synthetic(n=n, B=999, seed=20251018)
K = len(inst.iterator); B1 = inst.B + 1
lv = o.level_values(inst.nu, inst.iterator).astype(float)
cost = np.array([[o.jump_cost(lv[j], lv[l], inst.beta, inst.p) for l in range(K)] for j in range(K)]) if hasattr(o, "jump_cost") else None
if cost is None:
    d = np.abs(lv[:, None, :] - lv[None, :, :]).sum(2)
    cost = inst.beta * d
Kr = 128
nbr = Kr // blk
cpad = np.full((Kr, K), np.inf); cpad[:K] = cost
cminq = cpad.reshape(nbr, blk, K).min(axis=1)           # [q][l]
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
acc = {}
for i in range(n - 2, -1, -1):
    s, bt = stage_cost(i)
    v = (s[None, :, None] + cost.T[None, :, :]) + P[:, None, :]      # [b', l, j]
    best = v.min(axis=2)
    if i % 25 == 0 and np.isfinite(P).mean() > 0.9:
        Ppad = np.full((B1, Kr), 1e300); Ppad[:, :K] = P
        pm = Ppad.reshape(B1, nbr, blk).min(axis=2)                  # [b'][q]
        vpad = np.full((B1, K, Kr), np.inf); vpad[:, :, :K] = v
        vblk = vpad.reshape(B1, K, nbr, blk).min(axis=3)             # [b'][l][q]
        seed = pm.argmin(axis=1)                                     # [b']
        UB = np.minimum(vblk[np.arange(B1), :, seed], v[:, np.arange(K), np.arange(K)])   # [b'][l]
        LB = (s[None, :, None] + cminq.T[None, :, :]) + pm[:, None, :]                   # [b'][l][q]
        need = ~(LB > UB[:, :, None])                                                  # [b'][l][q]
        jstar = np.argmin(np.where(np.isnan(Ppad), np.inf, Ppad), axis=1)                # [b']  the row's smallest value
        UB1 = np.minimum(v[np.arange(B1), :, np.minimum(jstar, K - 1)], v[:, np.arange(K), np.arange(K)])
        need1 = ~(LB > UB1[:, :, None])
        needp = ~(LB > best[:, :, None])                                               # a perfect upper bound: the minimum itself
        cnt = dict(exact=0, exact1=0, perfect=0, super=0, rowlane=0, rowlane_any=0, tiles=0, exact_cells=0.0)
        for b0 in range(0, B1 - RW + 1, RW * 5):                      # a sample of row groups
            rows = slice(b0, b0 + RW)
            for l0 in range(0, K, 32):
                ls = slice(l0, min(K, l0 + 32))
                nd = need[rows, ls, :]
                ex = nd.any(axis=(0, 1))
                cnt["exact"] += ex.sum()
                cnt["exact1"] += need1[rows, ls, :].any(axis=(0, 1)).sum()
                cnt["perfect"] += needp[rows, ls, :].any(axis=(0, 1)).sum()
                # super-block coarse test of the round-2 kernel: rows merged, ubmax
                pms = pm[rows].reshape(RW, nbr // 4, 4).min(axis=(0, 2))                 # [Q]
                cms = cminq.reshape(nbr // 4, 4, K).min(axis=1)[:, ls]                  # [Q][l]
                ubmax = UB[rows, ls].max(axis=0)                                        # [l]
                lbs = (s[None, ls] + cms) + pms[:, None]
                sneed = (~(lbs > ubmax[None, :])).any(axis=1)
                cnt["super"] += 4 * sneed.sum()
                # per-row test merged over the tile's levels
                Cw = cminq[:, ls].min(axis=1)                                           # [q]
                U = (UB[rows, ls] - s[None, ls]).max(axis=1)                            # [r]
                rl = ~((Cw[None, :] + pm[rows]) > U[:, None])                           # [r][q]
                cnt["rowlane"] += rl.any(axis=0).sum()
                cnt["rowlane_any"] += rl.sum() / RW
                cnt["tiles"] += 1
        t = cnt["tiles"]
        print(f"stage {i:6d}: per tile of {nbr} blocks: exact {cnt['exact']/t:5.2f}  exact(UB from j* and self only) {cnt['exact1']/t:5.2f}  perfect-UB {cnt['perfect']/t:5.2f}  super(r2 coarse) {cnt['super']/t:5.2f}  "
              f"rowlane(union rows) {cnt['rowlane']/t:5.2f}  rowlane(per row) {cnt['rowlane_any']/t:5.2f}", flush=True)
    Pn = np.full_like(P, np.inf)
    for l in range(K):
        if bt[l] < B1:
            Pn[bt[l]:, l] = best[:B1 - bt[l], l]
    P = Pn
