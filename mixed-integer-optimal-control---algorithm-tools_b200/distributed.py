"""Multi-GPU sharding of independent subproblems.

One DP is sequential in time, so one subproblem lives on one GPU; the box is partitioned over INDEPENDENT
subproblems (multi-start x0; a sweep over trial radii costs no extra DP, multi-trust.jl:109-110).  There is no
data-path collective.  The only exchange is the final best-candidate reduction: every rank contributes one
16-byte (value, global subproblem index) record, gathered with torch.distributed.all_gather (NCCL over
NVLink on the GPU box, gloo in the CPU tests) and reduced locally with the deterministic lexicographic
rule "smallest value, then smallest index" (NCCL has no MINLOC), so every rank agrees on the winner.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def shard(S: int, rank: int, world: int):
    """Global subproblem indices owned by `rank`: s with s % world == rank (equal cost -> static)."""
    return list(range(rank, S, world))


def local_best(values, global_indices):
    """Deterministic (value, index) minimum of one rank's candidates via the C ABI."""
    lib = _lib.load()
    v = np.ascontiguousarray(values, dtype=np.float64).ravel()
    i = np.ascontiguousarray(global_indices, dtype=np.int64).ravel()
    bv, bi = ctypes.c_double(), ctypes.c_int64()
    _lib.check(lib.bb200_best_candidate(_lib.f64p(v), _lib.i64p(i), v.shape[0], ctypes.byref(bv), ctypes.byref(bi)))
    return bv.value, bi.value


def best_candidate(value: float, global_index: int, device=None, group=None):
    """All ranks call this with their local best; returns the global (value, index) on every rank."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value), int(global_index)
    world = dist.get_world_size(group)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    rec = torch.empty(2, dtype=torch.float64, device=dev)
    rec[0] = float(value)
    rec[1] = float(global_index)            # exact for indices < 2^53
    out = torch.empty(2 * world, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(out, rec, group=group)
    out = out.cpu().numpy().reshape(world, 2)
    return local_best(out[:, 0], out[:, 1].astype(np.int64))


def fetch_winner_control(u_local, owner_rank: int, n: int, M: int, device=None, group=None):
    """Broadcast the winning control trajectory (n*M*8 bytes) from the rank that owns it."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return np.asarray(u_local)
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    buf = torch.empty((n, M), dtype=torch.float64, device=dev)
    if dist.get_rank(group) == owner_rank:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(u_local, dtype=np.float64)))
    dist.broadcast(buf, src=owner_rank, group=group)
    return buf.cpu().numpy()
