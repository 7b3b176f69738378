"""ctypes binding of libbellman_b200.so -- the same entry points the Julia glue reaches via ccall.

Fails loudly when the CUDA extension is missing: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BELLMAN_B200_LIB", os.path.join(HERE, "libbellman_b200.so"))

OK, ERR_ARG, ERR_CUDA, ERR_INEXACT, ERR_STALE, ERR_STATE, ERR_NOMEM = range(7)
FLAG_STAGE_KERNELS, FLAG_NO_GRAPH, FLAG_FORCE_WAVEFRONT = 1, 2, 4

c_plan_p = ctypes.c_void_p
_F64P = ctypes.POINTER(ctypes.c_double)
_I64P = ctypes.POINTER(ctypes.c_int64)
_I32P = ctypes.POINTER(ctypes.c_int32)

# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header.
SIGNATURES = {
    "bb200_last_error": (ctypes.c_char_p, []),
    "bb200_version": (ctypes.c_int, []),
    "bb200_device_count": (ctypes.c_int, []),
    "bb200_plan_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                         ctypes.c_int64, _I64P, _I32P, _I64P, _F64P, ctypes.c_double,
                                         ctypes.c_int32, ctypes.c_uint32, ctypes.POINTER(c_plan_p)]),
    "bb200_plan_destroy": (ctypes.c_int, [c_plan_p]),
    "bb200_plan_set_stream": (ctypes.c_int, [c_plan_p, ctypes.c_void_p]),
    "bb200_plan_tune": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32]),
    "bb200_wave_geometry": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                           ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _I64P, ctypes.c_int32]),
    "bb200_bellman": (ctypes.c_int, [c_plan_p, _F64P, _F64P]),
    "bb200_select_and_backtrack": (ctypes.c_int, [c_plan_p, ctypes.c_int64, _F64P, _F64P, _I64P, _I64P]),
    "bb200_solve": (ctypes.c_int, [c_plan_p, _F64P, _F64P, ctypes.c_int64, _F64P, _F64P, _I64P, _I64P]),
    "bb200_solve_batched": (ctypes.c_int, [c_plan_p, ctypes.c_int64, _F64P, _F64P, ctypes.c_int32, _I64P,
                                           _F64P, _F64P, _I64P, _I64P, _I32P]),
    "bb200_solve_batched_shard": (ctypes.c_int, [c_plan_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _F64P, _F64P,
                                                 ctypes.c_int32, _I64P, _F64P, _F64P, _I64P, _I64P, _I32P]),
    "bb200_best_candidate": (ctypes.c_int, [_F64P, _I64P, ctypes.c_int64, _F64P, _I64P]),
    "bb200_nccl_version": (ctypes.c_int, []),
    "bb200_multi_create": (ctypes.c_int, [_I32P, ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_int64, _I64P, _I32P, _I64P, _F64P, ctypes.c_double,
                                          ctypes.c_int32, ctypes.c_uint32, ctypes.POINTER(c_plan_p)]),
    "bb200_multi_destroy": (ctypes.c_int, [c_plan_p]),
    "bb200_multi_solve_batched": (ctypes.c_int, [c_plan_p, ctypes.c_int64, _F64P, _F64P, ctypes.c_int32, _I64P, _F64P,
                                                 _F64P, _I64P, _I64P, _I32P, _F64P, _I64P, _I32P, _F64P]),
    "bb200_multi_stats": (ctypes.c_int, [c_plan_p, _F64P, ctypes.c_int32]),
    "bb200_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "bb200_comm_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                         ctypes.POINTER(c_plan_p)]),
    "bb200_comm_destroy": (ctypes.c_int, [c_plan_p]),
    "bb200_comm_best_candidate": (ctypes.c_int, [c_plan_p, ctypes.c_double, ctypes.c_int64, _F64P, _I64P, _I32P]),
    "bb200_comm_broadcast": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _F64P, ctypes.c_int64]),
    "bb200_upload": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _F64P, _F64P]),
    "bb200_upload_device": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]),
    "bb200_bellman_resident": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_int32]),
    "bb200_backtrack_resident": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_int64]),
    "bb200_download": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _F64P, _F64P, _I64P, _I64P]),
    "bb200_sync": (ctypes.c_int, [c_plan_p]),
    "bb200_export_phi": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _F64P]),
    "bb200_export_argmin": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64, _I64P,
                                           ctypes.c_int64]),
    "bb200_count_updates": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _I64P]),
    "bb200_pred_integral": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _F64P]),
    "bb200_tv": (ctypes.c_int, [c_plan_p, ctypes.c_int32, ctypes.c_double, _F64P]),
    "bb200_stats": (ctypes.c_int, [c_plan_p, _F64P, ctypes.c_int32]),
    "bb200_profile": (ctypes.c_int, [c_plan_p, ctypes.c_int32, _I64P, ctypes.c_int32]),
    "bb200_fp64_peak": (ctypes.c_int, [ctypes.c_int, ctypes.c_int32, ctypes.c_double, _F64P, _F64P]),
}


class BellmanB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[bb200 error {code}] {msg}")
        self.code = code


class InexactError(BellmanB200Error, ValueError):
    """u_old is not integer valued (Julia's InexactError at HelpFunctions.jl:37,57)."""


class StaleCellError(BellmanB200Error, IndexError):
    """Selection/backtrack reached a cell the DP never wrote."""


_lib = None


def _prefer_bundled_nccl():
    """bb200_multi.cu binds NCCL at run time (dlopen).  In a Python process that also imports torch, both must end up
    on the SAME NCCL: the dynamic loader reuses an already loaded libnccl.so.2 by SONAME, and torch's CUDA library needs
    the (newer) NCCL it ships with.  So, unless the user chose one, point the library at the pip-bundled NCCL when it
    exists; without torch (Julia, plain C) the system libnccl.so.2 is used."""
    if os.environ.get("BELLMAN_B200_NCCL"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["BELLMAN_B200_NCCL"] = cand
                return
    except (ImportError, ValueError, AttributeError):
        pass


def load():
    """Load the shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
        _prefer_bundled_nccl()
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc == OK:
        return
    msg = load().bb200_last_error().decode("utf-8", "replace")
    if rc == ERR_INEXACT:
        raise InexactError(rc, msg)
    if rc == ERR_STALE:
        raise StaleCellError(rc, msg)
    raise BellmanB200Error(rc, msg)


def f64p(a):
    return a.ctypes.data_as(_F64P) if a is not None else _F64P()


def i64p(a):
    return a.ctypes.data_as(_I64P) if a is not None else _I64P()


def i32p(a):
    return a.ctypes.data_as(_I32P) if a is not None else _I32P()
