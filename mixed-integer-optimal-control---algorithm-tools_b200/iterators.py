"""Host-side mirror of julia_opt/AdmissibleIterators.jl plus the flattening into device tables.

The enumeration itself stays on the host (in Julia for the real integration) so that any user-supplied
iterator keeps working; what the device needs is the flattened result:
    level_values int32[K][M]   nu[m][l_k[m]]
    grid_offset  int64[K]      0-based column-major offset of tuple k in the L1 x .. x LM grid
    jump_cost    double[K][K]  beta * (sum_m |nu_j[m]-nu_l[m]|^p)^(1/p)     (HelpFunctions.jl:63-67)
"""
from __future__ import annotations

import itertools
import math

import numpy as np


def product_iterator(nu):
    """AdmissibleIterators.jl:9-18: every 1-based index tuple, first index fastest."""
    ranges = [range(1, len(v) + 1) for v in nu]
    return [tuple(reversed(t)) for t in itertools.product(*reversed(ranges))]


def check_sum(l, nu, nx, lb, ub):
    """AdmissibleIterators.jl:41-49."""
    val = 0
    for i in range(nx):
        val += nu[i][l[i] - 1]
    return val >= lb and val <= ub


def bounded_sum_iterator(nu, lower_bound, upper_bound):
    """AdmissibleIterators.jl:26-34: product order filtered by the sum of the selected values."""
    nx = len(nu)
    return [l for l in product_iterator(nu) if check_sum(l, nu, nx, lower_bound, upper_bound)]


def flatten(nu, iterator):
    """(level_values int32 (K, M), grid_offset int64 (K,), grid_dims int64 (M,)) of an iterator."""
    tuples = [tuple(int(x) for x in l) for l in iterator]
    M = len(nu)
    dims = np.array([len(v) for v in nu], dtype=np.int64)
    lv = np.empty((len(tuples), M), dtype=np.int32)
    goff = np.empty(len(tuples), dtype=np.int64)
    for k, l in enumerate(tuples):
        if len(l) != M:
            raise ValueError("iterator tuple length differs from the number of controls")
        g, stride = 0, 1
        for m in range(M):
            if not 1 <= l[m] <= len(nu[m]):
                raise IndexError(f"index {l[m]} outside nu[{m}]")
            lv[k, m] = nu[m][l[m] - 1]
            g += (l[m] - 1) * stride
            stride *= len(nu[m])
        goff[k] = g
    return lv, goff, dims


def jump_cost_table(beta, p, level_values):
    """cost[j, l] = beta * (sum_m |nu_j[m]-nu_l[m]|^p)^(1/p), evaluated like HelpFunctions.jl:63-67:
    integer power when p is an int, Float64 power otherwise (p = inf), accumulated into a Float64
    starting from 0., then ^(1/p) as a Float64 power.  libm pow stands in for Julia's here; the Julia
    glue evaluates the table with Julia's own `^`."""
    lv = np.asarray(level_values, dtype=np.int64)
    K, M = lv.shape
    p_is_int = isinstance(p, (int, np.integer)) and not isinstance(p, bool)
    if not (p > 0):
        raise ValueError("Only positive integer valued `p` are accepted!")
    cost = np.empty((K, K), dtype=np.float64)
    inv = 1.0 / p
    for j in range(K):
        for l in range(K):
            tv = 0.0
            for m in range(M):
                d = abs(int(lv[j, m]) - int(lv[l, m]))
                tv += float(d ** int(p)) if p_is_int else _fpow(float(d), float(p))
            cost[j, l] = beta * _fpow(tv, inv)
    return cost


def _fpow(x: float, y: float) -> float:
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf
