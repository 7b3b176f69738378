"""Second, independent restatement of the hot path (TEST INFRASTRUCTURE ONLY): a literal pure-Python
transliteration of the reference's Julia, statement by statement, with 1-based column-major arrays.

    bellman_TRM!   /root/reference/HelpFunctions.jl:20-83
    eval_u_TRM!    /root/reference/HelpFunctions.jl:98-124
    product_iterator / bounded_sum_iterator / check_sum   /root/reference/julia_opt/AdmissibleIterators.jl:9-49

Purpose: `oracle/bellman_oracle.c` and the CUDA kernels were written from the same
reading of the Julia; a shared misreading would be invisible.  This file was written separately from the
Julia source alone -- it shares no code, layout or loop structure with the C oracle (it keeps the reference's
own `U[:, b, l..., i]` tuple table, `Phi[b, l..., slot]` indexing and recomputes the jump cost inside the
loops exactly where the reference does) -- and `tests/test_transliteration.py` cross-checks the two on
hundreds of random instances, the KATs and the committed golden vectors.  Pure-Python loops: small cases only.

PARITY UNPINNED: like the C oracle this has never been diffed against real Julia output (no Julia in the
image; `tools/make_julia_golden.jl` produces the dump that pins both wherever Julia 1.10 exists).

Julia semantics that are spelled out here because Python differs:
  * arrays are 1-based and column-major: `JArray` below; `A[b, l..., s]` splats the index tuple `l`
  * `Δt * ∇f[m,i] * numl` is the n-ary `*`, a left fold: (Δt * ∇f) * numl, each product rounded; `numl::Int64`
    is promoted to Float64                                                   (HelpFunctions.jl:35,56)
  * `convert(Int64, x::Float64)` throws InexactError unless x is integer valued      (:37,57,121)
  * `abs(a - b)^p`: Int64^Int64 is an exact integer power; with `p::Float64` (p = Inf) the base is
    promoted and `^` is the floating-point power                                     (:65)
  * `temp_val_2^(1/p)`: `1/p` is Float64 for every p; Float64^Float64                 (:67)
  * `argmin` = first index of `findmin`, column-major, ordered by `isless` with NaN first (Julia 1.10
    `_rf_findmin` / `isgreater`)                                                     (:106)
"""
from __future__ import annotations

import math

import numpy as np


class InexactError(ValueError):
    pass


class JArray:
    """Minimal Julia-style dense array: 1-based, column-major, scalar indexing and `A[:, rest...]` slices."""

    def __init__(self, dtype, *dims, fill=0):
        self.dims = tuple(int(d) for d in dims)
        self.a = np.full(self.dims, fill, dtype=dtype, order="F")

    @classmethod
    def wrap(cls, arr_f):
        out = cls.__new__(cls)
        out.a = arr_f
        out.dims = arr_f.shape
        return out

    def _ix(self, idx):
        assert len(idx) == len(self.dims), (idx, self.dims)
        for k, (i, d) in enumerate(zip(idx, self.dims)):
            if not (1 <= i <= d):
                raise IndexError(f"BoundsError: index {idx} dims {self.dims}")
        return tuple(i - 1 for i in idx)

    def __getitem__(self, idx):
        return self.a[self._ix(idx)]

    def __setitem__(self, idx, v):
        self.a[self._ix(idx)] = v

    def column(self, rest):
        """A[:, rest...] (a copy, like Julia's slicing)."""
        return [self.a[(m,) + self._ix((1,) + tuple(rest))[1:]] for m in range(self.dims[0])]

    def set_column(self, rest, values):
        """A[:, rest...] .= values"""
        tail = self._ix((1,) + tuple(rest))[1:]
        for m in range(self.dims[0]):
            self.a[(m,) + tail] = values[m]


def convert_Int64(x):
    """convert(Int64, x::Float64)"""
    if isinstance(x, (int, np.integer)):
        return int(x)
    if math.isnan(x) or math.isinf(x) or x != math.floor(x) or abs(x) >= 2.0 ** 63:
        raise InexactError(f"InexactError: Int64({x})")
    return int(x)


def jl_pow(x, y):
    """x^y for the operand types that occur at HelpFunctions.jl:65,67."""
    if isinstance(x, int) and isinstance(y, int):
        return x ** y                               # Int64^Int64 (y >= 0 here)
    x, y = float(x), float(y)
    if y == 0.0:
        return 1.0                                  # also Inf^0.0 and 0.0^0.0
    try:
        return math.pow(x, y)
    except OverflowError:
        return math.inf
    except ValueError:                              # 0.0^negative etc. never occurs here
        return math.nan


# ---- AdmissibleIterators.jl ----------------------------------------------------------------------------
def product_iterator(nu):
    """:9-18  Iterators.product(range_vec...): the FIRST range varies fastest."""
    nx = len(nu)
    sizes = [len(nu[i]) for i in range(nx)]
    out = []                                       # explicit odometer, first digit fastest
    l = [1] * nx
    total = 1
    for s in sizes:
        total *= s
    for _ in range(total):
        out.append(tuple(l))
        for i in range(nx):
            l[i] += 1
            if l[i] <= sizes[i]:
                break
            l[i] = 1
    return out


def check_sum(l, nu, nx, lb, ub):
    """:41-49"""
    val = 0
    for i in range(1, nx + 1):
        val += nu[i - 1][l[i - 1] - 1]
    return val >= lb and val <= ub


def bounded_sum_iterator(nu, lower_bound, upper_bound):
    """:26-34"""
    nx = len(nu)
    prod_iterator = product_iterator(nu)
    return [l for l in prod_iterator if check_sum(l, nu, nx, lower_bound, upper_bound)]


# ---- HelpFunctions.jl:20-83 ----------------------------------------------------------------------------
def bellman_TRM(grad_f, u_old, B, beta, p, dt, nu, U, Phi, iterator):
    """grad_f, u_old: JArray Float64[M, n];  U: JArray Int64[M, B+1, L1..LM, n-1];  Phi: JArray Float64[B+1, L1..LM, 2]."""
    M, n = u_old.dims
    grid = Phi.dims[1:-1]

    def fill_inf(slot):                            # Phi[Inds, slot] .= Inf
        Phi.a[..., slot - 1] = math.inf

    fill_inf((n + 1) % 2 + 1)                                                              # :27
    for l in iterator:                                                                     # :29
        b = 0
        temp_val_1 = 0.0
        for m in range(1, M + 1):
            numl = nu[m - 1][l[m - 1] - 1]
            temp_val_1 += (dt * float(grad_f[m, n])) * float(numl)                         # :35
            b += convert_Int64(abs(numl - float(u_old[m, n])))                             # :37
        if b <= B:                                                                         # :40
            Phi[(b + 1,) + tuple(l) + ((n + 1) % 2 + 1,)] = temp_val_1                     # :41
    for i in range(n - 1, 0, -1):                                                          # :45
        fill_inf((i + 1) % 2 + 1)                                                          # :47
        for l in iterator:                                                                 # :49
            temp_val_1 = 0.0
            bt = 0
            for m in range(1, M + 1):
                numl = nu[m - 1][l[m - 1] - 1]
                temp_val_1 += (dt * float(grad_f[m, i])) * float(numl)                     # :56
                bt += convert_Int64(abs(numl - float(u_old[m, i])))                        # :57
            for j in iterator:                                                             # :60
                temp_val_2 = 0.0
                for m in range(1, M + 1):
                    temp_val_2 += jl_pow(abs(nu[m - 1][j[m - 1] - 1] - nu[m - 1][l[m - 1] - 1]), p)   # :65
                temp_val_2 = temp_val_1 + beta * jl_pow(temp_val_2, 1 / p)   # :67
                for b in range(0, B - bt + 1):                                             # :69
                    val = temp_val_2 + float(Phi[(b + 1,) + tuple(j) + (i % 2 + 1,)])      # :71
                    tgt = (b + bt + 1,) + tuple(l)
                    if float(Phi[tgt + ((i + 1) % 2 + 1,)]) > val:                         # :73
                        U.set_column(tgt + (i,), j)                                        # :74
                        Phi[tgt + ((i + 1) % 2 + 1,)] = val                                # :75
    return None


# ---- HelpFunctions.jl:98-124 ---------------------------------------------------------------------------
def _isless(a, b):
    """Base.isless(::Float64, ::Float64): NaN is larger than everything, -0.0 < +0.0."""
    if math.isnan(a):
        return False
    if math.isnan(b):
        return True
    if a < b:
        return True
    if a == b:
        return math.copysign(1.0, a) < 0 and not math.copysign(1.0, b) < 0
    return False


def _isgreater(x, y):
    """Base.isgreater (Julia 1.10): isunordered(x) || isunordered(y) ? isless(x, y) : isless(y, x)."""
    if math.isnan(x) or math.isnan(y):
        return _isless(x, y)
    return _isless(y, x)


def eval_u_TRM(u, u_old, U, Phi, B, nu):
    M, n = u_old.dims
    grid = Phi.dims[1:-1]
    # index = argmin(@view Phi[1:B+1, Inds, 1]): iterate the view in column-major order                 :106
    fm, im = None, None
    cells = 1
    for d in grid:
        cells *= d
    l = [1] * len(grid)
    for _ in range(cells):
        for b1 in range(1, B + 2):
            fx = float(Phi[(b1,) + tuple(l) + (1,)])
            if fm is None or _isgreater(fm, fx):
                fm, im = fx, (b1,) + tuple(l)
        for k in range(len(grid)):
            l[k] += 1
            if l[k] <= grid[k]:
                break
            l[k] = 1
    index = im
    l = [index[m] for m in range(1, M + 1)]                                                # :108
    b = index[0] - 1                                                                       # :109
    for m in range(1, M + 1):
        u[m, 1] = nu[m - 1][l[m - 1] - 1]                                                  # :111
    for i in range(1, n):                                                                  # :115
        l = [int(x) for x in U.column((b + 1,) + tuple(l) + (i,))]                         # :116
        for m in range(1, M + 1):
            u[m, i + 1] = nu[m - 1][l[m - 1] - 1]                                          # :118
        nrm = 0.0
        for m in range(1, M + 1):
            nrm += abs(float(u[m, i]) - float(u_old[m, i]))                                # norm(.,1)
        b = convert_Int64(b - nrm)                                                         # :121
    return fm, index


# ---- convenience: run on numpy inputs in the oracle's array conventions ---------------------------------
def run(df, u_old, B, beta, p, dt, nu, iterator, radii=None, U_init=0):
    """df, u_old: numpy (n, M) (== Julia M x n).  Returns Phi and U as numpy arrays in the oracle's
    conventions (reversed shapes) and, per radius, (u (n, M), phi_star, index)."""
    n, M = u_old.shape
    dims = [len(v) for v in nu]
    g = JArray.wrap(np.asfortranarray(np.asarray(df, dtype=np.float64).T))
    uo = JArray.wrap(np.asfortranarray(np.asarray(u_old, dtype=np.float64).T))
    U = JArray(np.int64, M, B + 1, *dims, max(n - 1, 0), fill=U_init)       # multi-trust.jl:71-76
    Phi = JArray(np.float64, B + 1, *dims, 2, fill=0.0)                      # multi-trust.jl:77
    bellman_TRM(g, uo, B, beta, p, dt, nu, U, Phi, iterator)
    outs = []
    for Bn in ([B] if radii is None else radii):
        u = JArray(np.float64, M, n, fill=0.0)
        try:
            fm, index = eval_u_TRM(u, uo, U, Phi, Bn, nu)
            outs.append((np.ascontiguousarray(u.a.T), fm, index))
        except (IndexError, InexactError) as e:     # the reference would read a stale / zero U cell here
            outs.append((None, None, e))
    return np.ascontiguousarray(Phi.a.T), np.ascontiguousarray(U.a.T), outs
