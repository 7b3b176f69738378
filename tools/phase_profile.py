"""Prints the in-kernel cycle breakdown of the wavefront kernel (bb200_profile) on config-4-shaped input.
Usage: python tools/phase_profile.py [n] [ctas jsplit variant] [kind]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mioc_b200 as m
wl = importlib.import_module(m.__name__ + ".workloads")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
tune = [int(x) for x in sys.argv[2:5]] if len(sys.argv) > 4 else [0, 0, 0]
kind = sys.argv[5] if len(sys.argv) > 5 else "synthetic"
inst = wl.synthetic(n=n, B=999, seed=20251018) if kind == "synthetic" else wl.example_shaped(kind, n=n, seed=3)
plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=4)
plan.tune(*tune)
plan.upload(0, inst.df, inst.u_old)
plan.bellman_resident(0, 1); plan.sync()
plan.profile(True)
plan.bellman_resident(0, 1); plan.sync()
st = plan.stats()
prof = plan.profile(True, fetch=True)
G = int(st["ctas"])
print(f"{kind} n={n} K={inst.K} B={inst.B} variant={int(st['variant'])} ctas={G} rows={int(st['rows_per_cta'])} threads={int(st['threads'])} "
      f"js={int(st['jsplit'])} ns={int(st['scatter_warps'])} wave_ms={st['wave_ms']:.3f} "
      f"us/stage={st['wave_ms']*1e3/(n-1):.2f}  T upd/s={plan.count_updates()/st['wave_ms']/1e9:.3f}")
names_c = ["waitA", "phaseB", "handover", "waitB"]
names_s = ["wait_scan", "wait_halo_ring", "phaseC"]
names_x = ["C_combine", "C_store"]
names_m = ["trips", "idle", "pred_polls", "succ_polls"]
def line(row):
    stg = max(row[4], 1)
    return ("compute: " + " ".join(f"{nm}={row[k]/stg:7.0f}" for k, nm in enumerate(names_c)) +
            " | scatter: " + " ".join(f"{nm}={row[5+k]/stg:7.0f}" for k, nm in enumerate(names_s)) +
            " (" + " ".join(f"{nm}={row[14+k]/stg:6.0f}" for k, nm in enumerate(names_x)) + ")" +
            f" ring_wait={row[13]/stg:6.0f}" + " | comm/stage: " + " ".join(f"{nm}={row[8+k]/stg:6.2f}" for k, nm in enumerate(names_m)))
if st.get("prune_block", 0):
    print("pruned tiles: the scatter fields hold the sub-phases of the scan (warp 0): wait_scan = block minima + seed, wait_halo_ring = upper "
          "bounds, phaseC = masks, C_combine = scan of the surviving blocks, C_store = blocks scanned; "
          f"executed fraction of the launch = {st['executed_updates'] / plan.count_updates():.3f}")
for gsel in sorted({0, 1, 2, G // 2, G - 2, G - 1}):
    if 0 <= gsel < G:
        print(f"cta {gsel:3d} " + line(prof[gsel]))
print("avg     " + line(prof[:G].mean(axis=0)))
if os.environ.get("PROFILE_DUMP"):
    np.save(os.environ["PROFILE_DUMP"], prof[:G])
