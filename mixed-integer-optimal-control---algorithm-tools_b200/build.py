"""Builds libbellman_b200.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbellman_b200.so")
SOURCES = ["bb200_api.cu", "bb200_multi.cu", "kernels_common.cu", "kernel_wavefront.cu", "kernel_stage_pruned.cu", "kernels_microbench.cu"]
HEADERS = ["bb200_internal.cuh", "kernels.cuh", "pruned_scan.cuh", os.path.join("..", "..", "include", "bellman_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # B200 only; no other arch, no PTX fallback
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # belt and braces: the kernels use explicit _rn intrinsics on the value path
    "-Xcompiler", "-fPIC", "-shared",
    "-ldl",  # NCCL is bound at run time (bb200_multi.cu), not linked
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
