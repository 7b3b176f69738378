"""Writes tests/golden/*.json|npz.

kat.json   : the two hand-derived known-answer tests of SURVEY.md section 8c (KAT-1, KAT-2).  The expected
             values are typed in from the survey's derivation, NOT produced by running code.
seeded.npz : outputs of the CPU oracle (oracle/bellman_oracle.c) on a few seeded instances, so that the GPU
             tests also compare against committed vectors.  The reference itself (Julia) cannot run in the
             build container, so these are oracle-generated, not reference-generated: PARITY UNPINNED.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

INF = "inf"

KAT = {
    "KAT-1": {
        "nu": [[0, 1]], "iterator": "product", "n": 2, "dt": 1.0, "beta": 0.5, "p": 1, "B": 1,
        "df": [[1.0], [-1.0]],            # Julia [1.0 -1.0] (M x n) in (n, M) memory order
        "u_old": [[0.0], [0.0]],
        # Phi[b, l, slot] given as slot -> list over admissible k -> list over b
        "phi_slot1": [[0.0, -0.5], [INF, 1.5]],
        "phi_slot2": [[0.0, INF], [INF, -1.0]],
        # written U cells: [stage i, b (0-based), k (0-based admissible index)] -> 1-based tuple
        "U": [[1, 0, 0, [1]], [1, 1, 0, [2]], [1, 1, 1, [1]]],
        "select": {"1": {"b": 1, "k": 0, "u": [[0.0], [1.0]], "phi": -0.5}},
    },
    "KAT-2": {
        "nu": [[0, 1], [0, 1], [0, 1]], "iterator": ["bounded_sum", 1, 1], "n": 3, "dt": 0.5, "beta": 0.25,
        "p": 1, "B": 2,
        "df": [[0.5, -0.5, -0.5], [-0.5, 0.5, 0.5], [0.5, -0.5, -0.5]],
        "u_old": [[1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 0.0, 0.0]],
        "phi_slot1": [[0.25, INF, 0.25], [INF, INF, 0.25], [INF, INF, 0.25]],
        "phi_slot2": [[0.0, INF, 0.0], [INF, INF, 1.0], [INF, INF, 1.0]],
        # the exact tie between successors k=2 and k=3 at U[:,3,(2,1,1),2] goes to the earlier one, (1,2,1)
        "U": [[2, 2, 0, [1, 2, 1]]],
        "U_all_others_point_to": [2, 1, 1],
        "select": {"2": {"b": 0, "k": 0, "u": "u_old", "phi": 0.25},
                   "1": {"b": 0, "k": 0, "u": "u_old", "phi": 0.25},
                   "0": {"b": 0, "k": 0, "u": "u_old", "phi": 0.25}},
    },
}


def main():
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(KAT, f, indent=1)

    from oracle import oracle as o
    import mioc_b200 as m
    from importlib import import_module
    wl = import_module(m.__name__ + ".workloads")
    cases = {
        "synthetic_small": wl.synthetic(n=40, B=30, seed=11, levels=3, M=2),
        "synthetic_ties": wl.synthetic(n=40, B=30, seed=12, levels=3, M=2, tie_heavy=True),
        "fishing": wl.example_shaped("fishing", n=64, seed=3),
        "heat_ties": wl.example_shaped("heat", n=48, seed=4, tie_heavy=True),
        "convolution": wl.example_shaped("convolution", n=256, seed=5),
    }
    out = {}
    for name, inst in cases.items():
        U, Phi = o.alloc_tables(inst.nu, inst.n, inst.B)
        cost = o.jump_cost_table(inst.beta, inst.p, inst.nu, inst.iterator)
        nupd = o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi,
                             inst.iterator, cost=cost)
        radii = sorted({inst.B, inst.B // 2, inst.B // 4, 0}, reverse=True)
        us = []
        for r in radii:
            u = np.zeros_like(inst.u_old)
            o.eval_u_TRM(u, inst.u_old, U, Phi, r, inst.nu)
            us.append(u)
        out[name + "/df"] = inst.df
        out[name + "/u_old"] = inst.u_old
        out[name + "/cost"] = cost
        out[name + "/Phi"] = Phi
        out[name + "/radii"] = np.array(radii, dtype=np.int64)
        out[name + "/u"] = np.stack(us)
        out[name + "/n_updates"] = np.array([nupd], dtype=np.int64)
        out[name + "/meta"] = np.array(json.dumps({
            "nu": inst.nu, "iterator": [list(t) for t in inst.iterator], "n": inst.n, "B": inst.B,
            "dt": inst.dt, "beta": inst.beta, "p": "inf" if inst.p == float("inf") else inst.p}))
    np.savez_compressed(os.path.join(HERE, "seeded.npz"), **out)
    print("wrote kat.json and seeded.npz:", {k: v.shape for k, v in out.items() if k.endswith("/Phi")})


if __name__ == "__main__":
    main()
