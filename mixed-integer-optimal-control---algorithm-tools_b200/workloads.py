"""Synthetic subproblem instances of the shapes BASELINE.json names.

Host-side input generation only (numpy); used by bench.py, smoke() and the tests.  `random_start` mimics the
reference's rand_func_int (HelpFunctions.jl:204-225): a piecewise-constant admissible control with a given
number of jump times -- with numpy's generator instead of Julia's MersenneTwister/StatsBase.sample, so the
trajectories are realistic in shape, not identical to what Julia would draw.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .iterators import bounded_sum_iterator, flatten, product_iterator


@dataclass
class Instance:
    name: str
    nu: list
    iterator: list
    n: int
    B: int
    dt: float
    beta: float
    p: object
    df: np.ndarray = field(repr=False)      # (n, M)
    u_old: np.ndarray = field(repr=False)   # (n, M)

    @property
    def M(self):
        return len(self.nu)

    @property
    def K(self):
        return len(self.iterator)


def random_start(nu, iterator, n, jumps, rng):
    """Piecewise-constant random admissible control, `jumps` switching times in 2..n."""
    lv, _, _ = flatten(nu, iterator)
    jumps = min(jumps, max(n - 1, 0))
    t = np.sort(rng.choice(np.arange(2, n + 1), size=jumps, replace=False)) if jumps > 0 else np.array([], dtype=int)
    seg = np.searchsorted(t, np.arange(1, n + 1), side="right")        # segment index of every stage
    picks = rng.integers(0, lv.shape[0], size=jumps + 1)
    return lv[picks[seg]].astype(np.float64)


def synthetic(n=100_000, B=999, seed=20251018, *, levels=5, M=3, dt=1.0, beta=0.5, p=1, tie_heavy=False):
    """BASELINE config 4: nu = [[0..levels-1]] x M via product_iterator (K = levels^M, first index fastest),
    df ~ N(0,1) from default_rng(seed), u_old piecewise constant with n/10 jumps from default_rng(seed+1).
    tie_heavy quantises df to multiples of 2^-2 and uses beta = 2^-2 so that exact ties are frequent."""
    nu = [list(range(levels)) for _ in range(M)]
    it = product_iterator(nu)
    df = np.random.default_rng(seed).standard_normal((n, M))
    if tie_heavy:
        df = np.round(df * 4.0) / 4.0
        beta = 0.25
    u_old = random_start(nu, it, n, n // 10, np.random.default_rng(seed + 1))
    return Instance(f"synthetic_n{n}_K{len(it)}_B{B}" + ("_ties" if tie_heavy else ""), nu, it, n, B, dt,
                    beta, p, np.ascontiguousarray(df), np.ascontiguousarray(u_old))


def example_shaped(kind, n=1024, seed=7, tie_heavy=False):
    """Instances with the table shapes and TRM parameters of the reference's examples (multi-trust.jl:181-195)
    and synthetic df (the examples' own df needs their ODE/PDE models; see oracle/trm_harness.py for the ODE ones)."""
    rng = np.random.default_rng(seed)
    if kind == "fishing":          # example_fishing.jl:17-24, main: beta=1e-4, D0=2, p=Inf
        nu = [[0, 1]] * 3; it = bounded_sum_iterator(nu, 1, 1); T = 12.0; beta, d0, p = 1e-4, 2.0, float("inf")
    elif kind == "vanderpol":      # example_vanderpol.jl:15-22, main: beta=0.1, D0=1, p=Inf
        nu = [[0, 1]] * 3; it = bounded_sum_iterator(nu, 1, 1); T = 20.0; beta, d0, p = 0.1, 1.0, float("inf")
    elif kind == "doubletank":     # example_doubletank.jl:15-22, main: beta=1e-5, D0=2, p=Inf
        nu = [[0, 1]] * 3; it = bounded_sum_iterator(nu, 1, 1); T = 10.0; beta, d0, p = 1e-5, 2.0, float("inf")
    elif kind == "convolution":    # example_convolution.jl:19-25, main: beta=1e-4, D0=.125, p=1
        nu = [[-2, -1, 0, 1, 2]]; it = product_iterator(nu); T = 2.0; beta, d0, p = 1e-4, 0.125, 1
    elif kind == "heat":           # example_heat.jl:37-44, main: beta=1e-3, D0=2, p=2
        nu = [[0, 1, 2, 3, 4, 5]] * 2; it = product_iterator(nu); T = 10.0; beta, d0, p = 1e-3, 2.0, 2
    else:
        raise ValueError(kind)
    dt = T / n
    B = int(np.floor(d0 / dt))
    M = len(nu)
    df = rng.standard_normal((n, M))
    if tie_heavy:
        df = np.round(df * 4.0) / 4.0
        beta = 0.25
    u_old = random_start(nu, it, n, n // 10, rng)
    return Instance(f"{kind}_n{n}" + ("_ties" if tie_heavy else ""), [list(v) for v in nu], it, n, B, dt, beta, p,
                    np.ascontiguousarray(df), np.ascontiguousarray(u_old))
