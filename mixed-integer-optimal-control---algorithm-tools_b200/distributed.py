"""Multi-GPU sharding of independent subproblems for the one-process-per-GPU launch (torchrun).

One DP is sequential in time, so one subproblem lives on one GPU; the box is partitioned over INDEPENDENT
subproblems (multi-start x0; a sweep over trial radii costs no extra DP, multi-trust.jl:109-110).  There is no
data-path collective.  The only exchange is the final best-candidate reduction: every rank contributes one
16-byte (value, global subproblem index) record and all ranks reduce the gathered records with the deterministic
rule of the selection (Julia findmin order, then smallest index; NCCL has no MINLOC).

The collective itself lives behind the C ABI (bb200_comm_*: ncclAllGather / ncclBroadcast over NVLink, NCCL bound
at run time); this module is the thin caller: torch.distributed is only the side channel that ships the 128-byte NCCL
unique id from rank 0 to the other ranks.  Without GPUs (the gloo tests on CPU) the same records are gathered with
torch.distributed.all_gather and reduced by the same C function.  A single process that owns several GPUs does not
need this module at all: api.MultiPlan drives them with one call.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

_comm = None


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


def init_comm(local_device: int, group=None):
    """Creates this rank's NCCL communicator through the C ABI (call once per process after init_process_group with
    the nccl backend).  Returns the api.Comm, or None when there is nothing to communicate with."""
    global _comm
    import torch.distributed as dist
    from .api import Comm

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    _comm = Comm(local_device, world, rank, box[0])
    return _comm


def close_comm():
    global _comm
    if _comm is not None:
        _comm.close()
        _comm = None


def best_candidate(value: float, global_index: int, device=None, group=None):
    """All ranks call this with their local best; returns the global (value, index) on every rank."""
    if _comm is not None:
        bv, bi, _ = _comm.best_candidate(value, global_index)
        return bv, bi
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
    if _comm is not None:
        buf = np.ascontiguousarray(u_local, dtype=np.float64).copy() if _comm.rank == owner_rank else np.zeros((n, M))
        return _comm.broadcast(owner_rank, buf)
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
