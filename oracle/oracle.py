"""Python front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see bellman_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED (no Julia here; see bellman_oracle.c header).

Array conventions: every array has exactly the reference's memory layout.  Julia is
column-major, numpy here is C-order, so shapes appear reversed:

    Julia  df, u_old, u :: Float64[M, n]              numpy (n, M)
    Julia  Phi :: Float64[B+1, L1..LM, 2]             numpy (2, LM, .., L1, B+1)
    Julia  U   :: Int64[M, B+1, L1..LM, n-1]          numpy (n-1, LM, .., L1, B+1, M)

Reference lines restated here (relative to /root/reference):
    julia_opt/AdmissibleIterators.jl:9-18   product_iterator
    julia_opt/AdmissibleIterators.jl:26-34  bounded_sum_iterator
    julia_opt/AdmissibleIterators.jl:41-49  check_sum
    multi-trust.jl:69-77                    table allocation (alloc_tables)
"""
from __future__ import annotations

import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I64P = ctypes.POINTER(ctypes.c_int64)
_F64P = ctypes.POINTER(ctypes.c_double)
_I16P = ctypes.POINTER(ctypes.c_int16)

ORACLE_OK, ORACLE_ERR_INEXACT, ORACLE_ERR_ARG, ORACLE_ERR_STALE = 0, 1, 2, 3


class InexactError(ValueError):
    """Julia's InexactError: u_old holds a non-integer (HelpFunctions.jl:37,57)."""


def build(force: bool = False) -> None:
    """Compile oracle/liboracle.so and liboracle_omp.so with the committed Makefile."""
    need = force or not all(
        os.path.exists(os.path.join(_HERE, f)) for f in ("liboracle.so", "liboracle_omp.so"))
    if not need:
        src = os.path.getmtime(os.path.join(_HERE, "bellman_oracle.c"))
        need = any(os.path.getmtime(os.path.join(_HERE, f)) < src
                   for f in ("liboracle.so", "liboracle_omp.so"))
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B"])


_libs: dict = {}


def _lib(threads: bool = False):
    key = "omp" if threads else "st"
    if key not in _libs:
        build()
        lib = ctypes.CDLL(os.path.join(_HERE, "liboracle_omp.so" if threads else "liboracle.so"))
        lib.oracle_bellman_trm.restype = ctypes.c_int
        lib.oracle_bellman_trm.argtypes = [
            _F64P, _F64P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
            _I64P, _I64P, _I64P, ctypes.c_int64, _F64P, _I64P, _F64P, _I16P, _I64P]
        lib.oracle_eval_u_trm.restype = ctypes.c_int
        lib.oracle_eval_u_trm.argtypes = [
            _F64P, _F64P, _I64P, _F64P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_int64, _I64P, _I64P, _I64P, _I64P, _F64P]
        lib.oracle_jump_cost_table.restype = None
        lib.oracle_jump_cost_table.argtypes = [
            ctypes.c_double, ctypes.c_double, ctypes.c_int, _I64P, _I64P, _I64P,
            ctypes.c_int64, ctypes.c_int64, _F64P]
        lib.oracle_tv_p.restype = ctypes.c_double
        lib.oracle_tv_p.argtypes = [_F64P, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                    ctypes.c_int]
        lib.oracle_pred_integral.restype = ctypes.c_double
        lib.oracle_pred_integral.argtypes = [_F64P, _F64P, _F64P, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.c_double]
        lib.oracle_num_threads.restype = ctypes.c_int
        _libs[key] = lib
    return _libs[key]


def num_threads(threads: bool = True) -> int:
    return int(_lib(threads).oracle_num_threads())


# --------------------------------------------------------------------------------------
# AdmissibleIterators.jl
# --------------------------------------------------------------------------------------
def product_iterator(nu):
    """AdmissibleIterators.jl:9-18 -- all 1-based index tuples, FIRST index fastest."""
    ranges = [range(1, len(v) + 1) for v in nu]
    # itertools.product varies the LAST index fastest; reverse in, reverse out.
    return [tuple(reversed(t)) for t in itertools.product(*reversed(ranges))]


def check_sum(l, nu, lb, ub):
    """AdmissibleIterators.jl:41-49."""
    val = 0
    for i in range(len(nu)):
        val += nu[i][l[i] - 1]
    return lb <= val <= ub


def bounded_sum_iterator(nu, lower_bound, upper_bound):
    """AdmissibleIterators.jl:26-34 -- product order filtered by check_sum."""
    return [l for l in product_iterator(nu) if check_sum(l, nu, lower_bound, upper_bound)]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _flat_nu(nu):
    vals = np.array([v for row in nu for v in row], dtype=np.int64)
    off = np.zeros(len(nu) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in nu])
    return vals, off


def _p_args(p):
    """Julia `p` is an Int (1, 2, ..) or a Float64 (Inf)."""
    if isinstance(p, (int, np.integer)) and not isinstance(p, bool):
        return float(p), 1
    return float(p), 0


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


def level_values(nu, iterator):
    """nu_k[m] = nu[m][l_k[m]] for every admissible tuple; int64 (K, M)."""
    return np.array([[nu[m][l[m] - 1] for m in range(len(nu))] for l in iterator], dtype=np.int64)


def grid_offsets(nu, iterator):
    """0-based column-major offset of each admissible tuple in the L1 x .. x LM grid."""
    out = []
    for l in iterator:
        g, stride = 0, 1
        for m in range(len(nu)):
            g += (l[m] - 1) * stride
            stride *= len(nu[m])
        out.append(g)
    return np.array(out, dtype=np.int64)


def alloc_tables(nu, n, B):
    """multi-trust.jl:69-77: U = zeros(Int64, M, B+1, L.., n-1); Phi = zeros(B+1, L.., 2)."""
    M = len(nu)
    dims = [len(v) for v in nu]
    U = np.zeros((max(n - 1, 0), *reversed(dims), B + 1, M), dtype=np.int64)
    Phi = np.zeros((2, *reversed(dims), B + 1), dtype=np.float64)
    return U, Phi


def jump_cost_table(beta, p, nu, iterator, threads=False):
    """cost[j, l] = beta * (sum_m |nu_j[m]-nu_l[m]|^p)^(1/p)   (HelpFunctions.jl:63-67)."""
    vals, off = _flat_nu(nu)
    it = np.ascontiguousarray(np.array(iterator, dtype=np.int64).reshape(len(iterator), len(nu)))
    K = it.shape[0]
    cost = np.empty((K, K), dtype=np.float64)
    pv, pint = _p_args(p)
    _lib(threads).oracle_jump_cost_table(beta, pv, pint, _ptr(vals, _I64P), _ptr(off, _I64P),
                                         _ptr(it, _I64P), K, len(nu), _ptr(cost, _F64P))
    return cost


def bellman_TRM(df, u_old, B, beta, p, dt, nu, U, Phi, iterator, *, cost=None, argk=None,
                threads=False):
    """HelpFunctions.jl:20 signature.  Returns the number of innermost-loop executions.

    `U` may be None (skip the 8*M-byte-per-cell reference table); `argk` is an optional
    int16 (n-1, K, B+1) array receiving the compact winner table (-1 = not written).
    """
    df = np.ascontiguousarray(df, dtype=np.float64)
    u_old = np.ascontiguousarray(u_old, dtype=np.float64)
    n, M = u_old.shape
    assert df.shape == (n, M)
    vals, off = _flat_nu(nu)
    it = np.ascontiguousarray(np.array(iterator, dtype=np.int64).reshape(len(iterator), M))
    K = it.shape[0]
    if cost is None:
        cost = jump_cost_table(beta, p, nu, iterator)
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    assert cost.shape == (K, K)
    assert Phi.dtype == np.float64 and Phi.flags.c_contiguous
    assert Phi.shape == (2, *reversed([len(v) for v in nu]), B + 1)
    if U is not None:
        assert U.dtype == np.int64 and U.flags.c_contiguous
        assert U.shape == (n - 1, *reversed([len(v) for v in nu]), B + 1, M)
    if argk is not None:
        assert argk.dtype == np.int16 and argk.flags.c_contiguous and argk.shape == (n - 1, K, B + 1)
    nupd = ctypes.c_int64(0)
    rc = _lib(threads).oracle_bellman_trm(
        _ptr(df, _F64P), _ptr(u_old, _F64P), M, n, B, dt, _ptr(vals, _I64P), _ptr(off, _I64P),
        _ptr(it, _I64P), K, _ptr(cost, _F64P), _ptr(U, _I64P), _ptr(Phi, _F64P),
        _ptr(argk, _I16P), ctypes.byref(nupd))
    if rc == ORACLE_ERR_INEXACT:
        raise InexactError("u_old is not integer valued")
    if rc != ORACLE_OK:
        raise ValueError(f"oracle_bellman_trm failed rc={rc}")
    return nupd.value


def eval_u_TRM(u, u_old, U, Phi, B, nu, *, table_B=None, info=None):
    """HelpFunctions.jl:98 signature; `B` may be any budget <= the table's (S10).

    `info`, if a dict, receives b_star, g_star (0-based grid offset) and phi_star.
    """
    u_old = np.ascontiguousarray(u_old, dtype=np.float64)
    n, M = u_old.shape
    assert u.dtype == np.float64 and u.flags.c_contiguous and u.shape == (n, M)
    tb = Phi.shape[-1] - 1 if table_B is None else table_B
    vals, off = _flat_nu(nu)
    bs, gs, ps = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_double()
    rc = _lib().oracle_eval_u_trm(_ptr(u, _F64P), _ptr(u_old, _F64P), _ptr(U, _I64P),
                                  _ptr(Phi, _F64P), M, n, tb, B, _ptr(vals, _I64P),
                                  _ptr(off, _I64P), ctypes.byref(bs), ctypes.byref(gs),
                                  ctypes.byref(ps))
    if info is not None:
        info.update(b_star=bs.value, g_star=gs.value, phi_star=ps.value)
    if rc == ORACLE_ERR_STALE:
        raise IndexError("backtrack visited a cell the DP never wrote")
    if rc != ORACLE_OK:
        raise ValueError(f"oracle_eval_u_trm failed rc={rc}")


def TV_p(u, p):
    """HelpFunctions.jl:251-268 on a Float64 (n, M) control."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    pv, pint = _p_args(p)
    if not (pv > 0):
        raise ValueError("Only positive integer valued `p` are accepted!")
    return float(_lib().oracle_tv_p(_ptr(u, _F64P), u.shape[1], u.shape[0], pv, pint))


def pred_integral(df, u_old, u, dt):
    """multi-trust.jl:117-121."""
    df, u_old, u = (np.ascontiguousarray(a, dtype=np.float64) for a in (df, u_old, u))
    return float(_lib().oracle_pred_integral(_ptr(df, _F64P), _ptr(u_old, _F64P), _ptr(u, _F64P),
                                             u.shape[1], u.shape[0], dt))


def count_updates(u_old, B, nu, iterator):
    """Exact number of innermost-loop executions, N = sum_i K * sum_l max(0, B+1-b~_l(i))."""
    u_old = np.asarray(u_old, dtype=np.float64)
    lv = level_values(nu, iterator).astype(np.float64)          # (K, M)
    K = lv.shape[0]
    total = 0
    step = 4096
    for i0 in range(0, u_old.shape[0] - 1, step):
        blk = u_old[i0:min(i0 + step, u_old.shape[0] - 1)]      # stages 1..n-1
        bt = np.abs(lv[None, :, :] - blk[:, None, :]).sum(axis=2).astype(np.int64)
        total += int(K * np.maximum(0, B + 1 - bt).sum())
    return total


# --------------------------------------------------------------------------------------
# Brute force (tiny instances only): independent check of what the DP is supposed to compute.
# --------------------------------------------------------------------------------------
def brute_force(df, u_old, B, dt, nu, iterator, cost):
    """min over all admissible trajectories with sum|u-u_old|_1 <= B of
    sum_i dt*df_i . nu_i + sum_{i<n} cost[k_{i+1}, k_i].  Different summation order than the DP:
    compare with a tolerance."""
    df = np.asarray(df, dtype=np.float64)
    u_old = np.asarray(u_old, dtype=np.float64)
    n, M = u_old.shape
    lv = level_values(nu, iterator)
    K = lv.shape[0]
    best = np.inf
    best_traj = None
    for traj in itertools.product(range(K), repeat=n):
        used = sum(int(abs(lv[traj[i]] - u_old[i]).sum()) for i in range(n))
        if used > B:
            continue
        val = sum(float(dt * df[i] @ lv[traj[i]]) for i in range(n))
        val += sum(cost[traj[i + 1], traj[i]] for i in range(n - 1))
        if val < best - 1e-12:
            best, best_traj = val, traj
    return best, best_traj
