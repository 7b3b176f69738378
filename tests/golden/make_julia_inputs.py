"""Writes tests/golden/julia_in/<case>.txt: the inputs of every known-answer test and committed golden instance
in a dependency-free text format that tools/make_julia_golden.jl reads under Julia 1.10 (no JSON.jl needed).
Floats are written as hexadecimal IEEE-754 bit patterns, so both sides see identical bits.

    python tests/golden/make_julia_inputs.py          (run from the repo root; CPU only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def hexwords(a):
    return " ".join(f"{w:016x}" for w in np.ascontiguousarray(a, dtype=np.float64).view(np.uint64).ravel())


def write_case(path, name, nu, iterator_kind, n, B, dt, beta, p, df, u_old, radii):
    M = len(nu)
    with open(path, "w") as f:
        f.write(f"name {name}\nM {M}\nn {n}\nB {B}\n")
        f.write(f"dt {hexwords([dt])}\nbeta {hexwords([beta])}\n")
        if isinstance(p, float) and np.isinf(p):
            f.write("p Inf\n")                       # Julia: p = Inf (Float64), multi-trust.jl:183,186,189
        else:
            f.write(f"p {int(p)}\n")                 # Julia: p::Int (TRM_parameters default p = 1; heat p = 2)
        f.write("nu " + " ; ".join(" ".join(str(int(v)) for v in row) for row in nu) + "\n")
        f.write("iterator " + " ".join(str(x) for x in iterator_kind) + "\n")
        f.write("radii " + " ".join(str(int(r)) for r in radii) + "\n")
        f.write("df " + hexwords(df) + "\n")          # (n, M) C-order == Julia M x n column-major
        f.write("u_old " + hexwords(u_old) + "\n")


def main():
    from helpers import load_kats, load_seeded
    from oracle import oracle as o
    out = os.path.join(HERE, "julia_in")
    os.makedirs(out, exist_ok=True)
    for name, kat in load_kats().items():
        kind = ["product"] if kat["iterator"] == "product" else list(kat["iterator"])
        write_case(os.path.join(out, name + ".txt"), name, kat["nu"], kind, kat["n"], kat["B"], kat["dt"], kat["beta"],
                   kat["p"], np.array(kat["df"], dtype=np.float64), np.array(kat["u_old"], dtype=np.float64),
                   list(range(kat["B"], -1, -1)))
    for name, g in load_seeded().items():
        meta = g["meta"]
        it = [tuple(t) for t in meta["iterator"]]
        if it == [tuple(t) for t in o.product_iterator(meta["nu"])]:
            kind = ["product"]
        elif it == [tuple(t) for t in o.bounded_sum_iterator(meta["nu"], 1, 1)]:
            kind = ["bounded_sum", 1, 1]
        else:
            raise SystemExit(f"{name}: iterator is neither product nor bounded_sum(1,1)")
        write_case(os.path.join(out, name + ".txt"), name, meta["nu"], kind, meta["n"], meta["B"], meta["dt"],
                   meta["beta"], meta["p"], g["df"], g["u_old"], [int(r) for r in g["radii"]])
    print("wrote", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
