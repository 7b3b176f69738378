"""Host-side mirror of the reference's operator interface for the hot path.

    bellman_TRM(df, u_old, B, beta, p, dt, nu, U, Phi, iterator)    HelpFunctions.jl:20
    eval_u_TRM(u, u_old, U, Phi, B, nu)                             HelpFunctions.jl:98

Same names, argument order and meaning as the reference; arrays use the reference's memory layout, which
in C-order numpy reads  df/u_old/u: (n, M),  Phi: (2, LM, .., L1, B+1),  U: (n-1, LM, .., L1, B+1, M).
Everything is forwarded to the C ABI (include/bellman_b200.h); nothing is computed in Python.

`TRMPlan` is the object the Julia glue keeps per `TRM` run (keyed on objectid(U)); the two free functions
keep one plan per `U` array the same way, so a transliterated TRM loop runs unchanged.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np

from . import _lib
from .iterators import flatten, jump_cost_table


def _as_io(a, n, M, name):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if a.shape != (n, M):
        raise ValueError(f"{name} must have shape (n, M) = ({n}, {M}); got {a.shape}")
    return a


class TRMPlan:
    """Device-resident state of one TRM run: tables, value rows, packed argmin table (multi-trust.jl:69-77)."""

    def __init__(self, nu, iterator, n, B, beta, p, dt, *, device=0, batch=1, flags=0, cost=None):
        self.lib = _lib.load()
        for v in nu:
            for x in v:
                if float(x) != int(x):      # the reference's nu is Vector{Vector{Int64}} (multi-trust.jl:64)
                    raise ValueError(f"control levels must be integers (the reference's nu is Int64); got {x!r}")
        self.nu = [[int(x) for x in v] for v in nu]
        self.level_values, self.grid_offset, self.grid_dims = flatten(self.nu, iterator)
        self.n, self.M, self.K, self.B = int(n), len(self.nu), int(self.level_values.shape[0]), int(B)
        self.dt = float(dt)
        self.beta, self.p = float(beta), p
        self.batch = int(batch)
        if cost is None:
            cost = jump_cost_table(beta, p, self.level_values)
        self.cost = np.ascontiguousarray(cost, dtype=np.float64)
        if self.cost.shape != (self.K, self.K):
            raise ValueError("jump cost table must be K x K")
        handle = _lib.c_plan_p()
        _lib.check(self.lib.bb200_plan_create(
            int(device), self.n, self.M, self.K, self.B, _lib.i64p(self.grid_dims),
            _lib.i32p(np.ascontiguousarray(self.level_values)), _lib.i64p(self.grid_offset),
            _lib.f64p(self.cost), self.dt, self.batch, int(flags), ctypes.byref(handle)))
        self._h = handle
        self._fin = weakref.finalize(self, self.lib.bb200_plan_destroy, handle)

    # ---- reference-shaped entry points ---------------------------------------------------------
    def bellman(self, df, u_old):
        """bellman_TRM! for budget B (HelpFunctions.jl:20-83)."""
        df = _as_io(df, self.n, self.M, "df")
        u_old = _as_io(u_old, self.n, self.M, "u_old")
        _lib.check(self.lib.bb200_bellman(self._h, _lib.f64p(df), _lib.f64p(u_old)))

    def eval_u(self, u, B_new=None):
        """eval_u_TRM! (HelpFunctions.jl:98-124); returns (phi_star, b_star, k_star)."""
        if u.dtype != np.float64 or not u.flags.c_contiguous or u.shape != (self.n, self.M):
            raise ValueError("u must be a C-contiguous float64 (n, M) array")
        ps, bs, ks = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.bb200_select_and_backtrack(
            self._h, self.B if B_new is None else int(B_new), _lib.f64p(u), ctypes.byref(ps),
            ctypes.byref(bs), ctypes.byref(ks)))
        return ps.value, bs.value, ks.value

    def solve(self, df, u_old, u, B_new=None):
        """One TR inner iteration: DP + selection + backtrack in one call."""
        df = _as_io(df, self.n, self.M, "df")
        u_old = _as_io(u_old, self.n, self.M, "u_old")
        ps, bs, ks = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.bb200_solve(
            self._h, _lib.f64p(df), _lib.f64p(u_old), self.B if B_new is None else int(B_new),
            _lib.f64p(u), ctypes.byref(ps), ctypes.byref(bs), ctypes.byref(ks)))
        return ps.value, bs.value, ks.value

    def solve_batched(self, df_all, u_old_all, radii, want_u=True, *, first=0, stride=1, out=None, strict=True):
        """S independent subproblems, each with len(radii) selections from its one table.

        first/stride restrict the call to the shard {first, first+stride, ..} (indices stay global; the other rows of
        the outputs are left untouched).  Returns (u_out, phi, b_star, k_star); with strict=False per-entry failures
        (InexactError of one subproblem, a stale selection) do not raise and a fifth value, status (S, R), is returned."""
        df_all = np.ascontiguousarray(df_all, dtype=np.float64)
        u_old_all = np.ascontiguousarray(u_old_all, dtype=np.float64)
        S = df_all.shape[0]
        if df_all.shape != (S, self.n, self.M) or u_old_all.shape != df_all.shape:
            raise ValueError("df_all/u_old_all must have shape (S, n, M)")
        radii = np.ascontiguousarray(radii, dtype=np.int64)
        R = radii.shape[0]
        if out is None:
            u_out = np.zeros((S, R, self.n, self.M), dtype=np.float64) if want_u else None
            phi = np.full((S, R), np.nan, dtype=np.float64)
            bs = np.full((S, R), -1, dtype=np.int64)
            ks = np.full((S, R), -1, dtype=np.int64)
            status = np.zeros((S, R), dtype=np.int32)
        else:
            u_out, phi, bs, ks, status = out
        rc = self.lib.bb200_solve_batched_shard(
            self._h, S, int(first), int(stride), _lib.f64p(df_all), _lib.f64p(u_old_all), R, _lib.i64p(radii),
            _lib.f64p(u_out), _lib.f64p(phi), _lib.i64p(bs), _lib.i64p(ks), _lib.i32p(status))
        if strict or rc not in (_lib.ERR_INEXACT, _lib.ERR_STALE):
            _lib.check(rc)
        return (u_out, phi, bs, ks) if strict else (u_out, phi, bs, ks, status)

    # ---- resident interface --------------------------------------------------------------------
    def upload(self, slot, df, u_old):
        df = _as_io(df, self.n, self.M, "df")
        u_old = _as_io(u_old, self.n, self.M, "u_old")
        _lib.check(self.lib.bb200_upload(self._h, slot, _lib.f64p(df), _lib.f64p(u_old)))
        _lib.check(self.lib.bb200_sync(self._h))  # the numpy temporaries may die after return

    def upload_device(self, slot, d_df_ptr, d_u_old_ptr):
        _lib.check(self.lib.bb200_upload_device(self._h, slot, ctypes.c_void_p(d_df_ptr),
                                                ctypes.c_void_p(d_u_old_ptr)))

    def bellman_resident(self, slot0=0, count=1):
        _lib.check(self.lib.bb200_bellman_resident(self._h, slot0, count))

    def backtrack_resident(self, slot=0, B_new=None):
        _lib.check(self.lib.bb200_backtrack_resident(self._h, slot, self.B if B_new is None else int(B_new)))

    def download(self, slot, u):
        ps, bs, ks = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.bb200_download(self._h, slot, _lib.f64p(u), ctypes.byref(ps), ctypes.byref(bs),
                                           ctypes.byref(ks)))
        return ps.value, bs.value, ks.value

    def sync(self):
        _lib.check(self.lib.bb200_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        _lib.check(self.lib.bb200_plan_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def tune(self, ctas=0, jsplit=0, variant=0):
        _lib.check(self.lib.bb200_plan_tune(self._h, ctas, jsplit, variant))

    # ---- inspection ----------------------------------------------------------------------------
    def export_phi(self, slot=0):
        """Reference-shaped value table (2, LM, .., L1, B+1); +Inf in inadmissible grid cells."""
        out = np.empty((2, *reversed([int(d) for d in self.grid_dims]), self.B + 1), dtype=np.float64)
        _lib.check(self.lib.bb200_export_phi(self._h, slot, _lib.f64p(out)))
        return out

    def export_argmin(self, i0=1, i1=None, fill=0, slot=0):
        """Reference-shaped argmin table for stages [i0, i1): (i1-i0, LM, .., L1, B+1, M)."""
        i1 = self.n if i1 is None else i1
        out = np.empty((i1 - i0, *reversed([int(d) for d in self.grid_dims]), self.B + 1, self.M), dtype=np.int64)
        _lib.check(self.lib.bb200_export_argmin(self._h, slot, i0, i1, _lib.i64p(out), fill))
        return out

    def count_updates(self, slot=0):
        v = ctypes.c_int64()
        _lib.check(self.lib.bb200_count_updates(self._h, slot, ctypes.byref(v)))
        return v.value

    def pred_integral(self, slot=0):
        v = ctypes.c_double()
        _lib.check(self.lib.bb200_pred_integral(self._h, slot, ctypes.byref(v)))
        return v.value

    def tv(self, p, slot=0):
        v = ctypes.c_double()
        _lib.check(self.lib.bb200_tv(self._h, slot, float(p), ctypes.byref(v)))
        return v.value

    def stats(self):
        out = np.zeros(22, dtype=np.float64)
        _lib.check(self.lib.bb200_stats(self._h, _lib.f64p(out), 22))
        keys = ("dp_ms", "backtrack_ms", "launches", "path", "ctas", "rows_per_cta", "arg_bytes",
                "device_bytes", "threads", "jsplit", "wave_ms", "graph_replays", "variant", "scatter_warps",
                "batch_ms", "batch_waves", "batch_syncs", "executed_updates", "prune_block", "prune_switched_off",
                "ctas_full_rows", "rows_per_cta_top")
        return dict(zip(keys, out.tolist()))

    def profile(self, enable=True, fetch=False, max_ctas=148):
        """Switch the wavefront kernel's cycle counters on/off; fetch=True returns the last launch's (ctas, 16) array."""
        out = np.zeros((max_ctas, 16), dtype=np.int64) if fetch else None
        _lib.check(self.lib.bb200_profile(self._h, int(enable), _lib.i64p(out), max_ctas if fetch else 0))
        return out

    def close(self):
        self._fin()


def fp64_peak(device=0, mode=0, target_ms=300.0):
    """FP64-pipe issue rate of `device` in lane-operations per second (roofline denominator)."""
    lib = _lib.load()
    ops, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.bb200_fp64_peak(int(device), int(mode), float(target_ms), ctypes.byref(ops), ctypes.byref(ms)))
    return ops.value, ms.value


# ---- drop-in free functions (one plan per U array, like the Julia glue's objectid(U) key) -----------
_plans: dict = {}


class _Entry:
    """Plan + what it was built for + the trajectory the last bellman_TRM already produced."""
    __slots__ = ("plan", "sig", "u_cache", "cache_ok")

    def __init__(self, plan, sig):
        self.plan, self.sig = plan, sig
        self.u_cache = np.empty((plan.n, plan.M), dtype=np.float64)
        self.cache_ok = False


def _signature(n, M, B, beta, p, dt, nu, it):
    """Everything that is baked into a plan at creation.  The reference's bellman_TRM! is stateless, so a caller may
    reuse U/Phi with another beta, p, dt, nu or iterator (beta-continuation, parameter sweeps): the plan is then
    rebuilt instead of silently answering for the old parameters."""
    return (int(n), int(M), int(B), float(beta), (type(p).__name__, float(p)), float(dt),
            tuple(tuple(int(x) for x in v) for v in nu), tuple(tuple(int(x) for x in t) for t in it))


def _entry_for(U, Phi):
    key = id(U) if U is not None else id(Phi)
    return key, _plans.get(key)


def bellman_TRM(df, u_old, B, beta, p, dt, nu, U, Phi, iterator, *, device=0, write_back=False):
    """Drop-in for bellman_TRM! (HelpFunctions.jl:20).  The device keeps its own packed tables; the
    caller's U/Phi are only filled when write_back=True (parity/debug).

    The reference always calls eval_u_TRM!(.., B, ..) right after this (multi-trust.jl:112-113), so the whole inner
    iteration -- H2D, DP, selection and backtrack for budget B, D2H -- goes out as ONE bb200_solve (a CUDA-graph
    replay, one synchronisation); the trajectory is kept and handed out by the eval_u_TRM that follows."""
    u_old_a = np.asarray(u_old)
    n, M = u_old_a.shape
    it = [tuple(t) for t in iterator]
    sig = _signature(n, M, B, beta, p, dt, nu, it)
    key, ent = _entry_for(U, Phi)
    if ent is None or ent.sig != sig:
        if ent is not None:
            ent.plan.close()
        ent = _Entry(TRMPlan(nu, it, n, B, beta, p, dt, device=device), sig)
        if key not in _plans:
            anchor = U if U is not None else Phi
            try:
                weakref.finalize(anchor, _plans.pop, key, None)
            except TypeError:
                pass
        _plans[key] = ent
    plan = ent.plan
    ent.cache_ok = False
    try:
        plan.solve(df, u_old, ent.u_cache, B)
        ent.cache_ok = True
    except _lib.StaleCellError:
        pass    # no feasible trajectory for budget B: the reference fails in eval_u_TRM!, not here; the DP is resident
    if write_back:
        if Phi is not None:
            Phi[...] = plan.export_phi()
        if U is not None and n > 1:
            U[...] = plan.export_argmin(1, n, fill=0)
    return None


def eval_u_TRM(u, u_old, U, Phi, B, nu, *, info=None):
    """Drop-in for eval_u_TRM! (HelpFunctions.jl:98); B may be any budget <= the table's."""
    _, ent = _entry_for(U, Phi)
    if ent is None:
        raise _lib.BellmanB200Error(_lib.ERR_STATE, "eval_u_TRM called before bellman_TRM for these tables")
    plan = ent.plan
    if int(B) == plan.B and ent.cache_ok and info is None:
        if u.shape != ent.u_cache.shape:
            raise ValueError("u must have shape (n, M)")
        u[...] = ent.u_cache       # produced by the bb200_solve of the bellman_TRM just before
        return None
    ps, bs, ks = plan.eval_u(u, B)
    if info is not None:
        info.update(phi_star=ps, b_star=bs, k_star=ks, g_star=int(plan.grid_offset[ks]) if ks >= 0 else -1)
    return None


# ---- multi-GPU behind the C ABI ---------------------------------------------------------
def nccl_version():
    """NCCL version the library bound at run time (0: NCCL not available)."""
    return int(_lib.load().bb200_nccl_version())


class MultiPlan:
    """All GPUs of this process behind one handle: one plan and one host thread per device; subproblem s runs on
    device s mod G; the best-candidate reduction is an ncclAllGather of 16-byte records inside the library."""

    def __init__(self, devices, nu, iterator, n, B, beta, p, dt, *, batch_per_device=1, flags=0, cost=None):
        self.lib = _lib.load()
        self.nu = [[int(x) for x in v] for v in nu]
        self.level_values, self.grid_offset, self.grid_dims = flatten(self.nu, iterator)
        self.n, self.M, self.K, self.B = int(n), len(self.nu), int(self.level_values.shape[0]), int(B)
        self.devices = np.ascontiguousarray(devices, dtype=np.int32)
        if cost is None:
            cost = jump_cost_table(beta, p, self.level_values)
        self.cost = np.ascontiguousarray(cost, dtype=np.float64)
        handle = _lib.c_plan_p()
        _lib.check(self.lib.bb200_multi_create(
            _lib.i32p(self.devices), int(self.devices.shape[0]), self.n, self.M, self.K, self.B,
            _lib.i64p(self.grid_dims), _lib.i32p(np.ascontiguousarray(self.level_values)), _lib.i64p(self.grid_offset),
            _lib.f64p(self.cost), float(dt), int(batch_per_device), int(flags), ctypes.byref(handle)))
        self._h = handle
        self._fin = weakref.finalize(self, self.lib.bb200_multi_destroy, handle)

    def solve_batched(self, df_all, u_old_all, radii, want_u=True, want_best_u=True, strict=True):
        """Returns dict(u, phi, b_star, k_star, status, best_value, best_subproblem, best_radius, u_best)."""
        df_all = np.ascontiguousarray(df_all, dtype=np.float64)
        u_old_all = np.ascontiguousarray(u_old_all, dtype=np.float64)
        S = df_all.shape[0]
        if df_all.shape != (S, self.n, self.M) or u_old_all.shape != df_all.shape:
            raise ValueError("df_all/u_old_all must have shape (S, n, M)")
        radii = np.ascontiguousarray(radii, dtype=np.int64)
        R = radii.shape[0]
        u_out = np.zeros((S, R, self.n, self.M), dtype=np.float64) if want_u else None
        phi = np.full((S, R), np.nan)
        bs = np.full((S, R), -1, dtype=np.int64)
        ks = np.full((S, R), -1, dtype=np.int64)
        status = np.zeros((S, R), dtype=np.int32)
        u_best = np.zeros((self.n, self.M)) if want_best_u else None
        bv, bsub, brad = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int32()
        rc = self.lib.bb200_multi_solve_batched(
            self._h, S, _lib.f64p(df_all), _lib.f64p(u_old_all), R, _lib.i64p(radii), _lib.f64p(u_out), _lib.f64p(phi),
            _lib.i64p(bs), _lib.i64p(ks), _lib.i32p(status), ctypes.byref(bv), ctypes.byref(bsub), ctypes.byref(brad),
            _lib.f64p(u_best))
        if strict or rc not in (_lib.ERR_INEXACT, _lib.ERR_STALE):
            _lib.check(rc)
        return dict(u=u_out, phi=phi, b_star=bs, k_star=ks, status=status, best_value=bv.value,
                    best_subproblem=bsub.value, best_radius=brad.value, u_best=u_best)

    def stats(self):
        out = np.zeros(2 + len(self.devices), dtype=np.float64)
        _lib.check(self.lib.bb200_multi_stats(self._h, _lib.f64p(out), out.shape[0]))
        return dict(devices=int(out[0]), wall_ms=out[1], device_ms=out[2:].tolist())

    def close(self):
        self._fin()


class Comm:
    """One rank of an NCCL communicator created through the C ABI (one process per GPU).  The 128-byte unique id is
    drawn by rank 0 with `Comm.unique_id()` and shipped to the other ranks by the host program."""

    @staticmethod
    def unique_id() -> bytes:
        buf = ctypes.create_string_buffer(128)
        _lib.check(_lib.load().bb200_comm_unique_id(ctypes.cast(buf, ctypes.c_void_p)))
        return buf.raw

    def __init__(self, device, nranks, rank, uid: bytes):
        self.lib = _lib.load()
        if len(uid) != 128:
            raise ValueError("the NCCL unique id is 128 bytes")
        buf = ctypes.create_string_buffer(uid, 128)
        handle = _lib.c_plan_p()
        _lib.check(self.lib.bb200_comm_create(int(device), int(nranks), int(rank), ctypes.cast(buf, ctypes.c_void_p),
                                              ctypes.byref(handle)))
        self._h = handle
        self.rank, self.nranks = int(rank), int(nranks)
        self._fin = weakref.finalize(self, self.lib.bb200_comm_destroy, handle)

    def best_candidate(self, value, index):
        bv, bi, own = ctypes.c_double(), ctypes.c_int64(), ctypes.c_int32()
        _lib.check(self.lib.bb200_comm_best_candidate(self._h, float(value), int(index), ctypes.byref(bv), ctypes.byref(bi),
                                                      ctypes.byref(own)))
        return bv.value, bi.value, own.value

    def broadcast(self, root, array):
        if array.dtype != np.float64 or not array.flags.c_contiguous:
            raise ValueError("broadcast needs a C-contiguous float64 array")
        _lib.check(self.lib.bb200_comm_broadcast(self._h, int(root), _lib.f64p(array), int(array.size)))
        return array

    def close(self):
        self._fin()
