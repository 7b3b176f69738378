"""Restatement of the reference's outer trust-region loop and ODE examples -- TEST INFRASTRUCTURE ONLY.

Used by tests/ to check the TRM OBJECTIVE HISTORY: the same loop is run twice, once with the oracle's
bellman_TRM / eval_u_TRM and once with the device path's, on the same start control; because every DP result
must be bit-identical, the two runs must produce identical log tables (Iter, k, radius, J, pred, ared, step) and
identical final controls.  The product never imports this module.

PARITY UNPINNED against Julia (no Julia in the build container): the ODE right-hand sides follow the reference
line by line, but the floating-point history is only compared oracle-vs-device, not against a Julia run.

Reference lines restated (relative to /root/reference):
    multi-trust.jl:53-170                      TRM
    multi-trust.jl:181-189                     main()'s parameters for fishing / doubletank / vanderpol
    julia_opt/AbstractObjective.jl:81-102      eval_f! / eval_df! lazy contract
    julia_opt/ODEObjective.jl:125-150          eval_f_helper (forward Euler + trapezoidal rule)
    julia_opt/ODEObjective.jl:153-184          eval_df_helper (adjoint Euler + gradient)
    julia_opt/example_fishing.jl:56-92, example_vanderpol.jl:48-81, example_doubletank.jl:48-82   F, Fy, Fu, G, Gy
Arrays use the reference's memory layout in C order: x, df: (nt, nx); state, adjoint: (nt, ny).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import oracle as o


# --------------------------------------------------------------------------------------------------------
# ODE objectives
# --------------------------------------------------------------------------------------------------------
class ODEObjective:
    """AbstractODEObjective + AbstractObjectiveLazy (eval counters, df_valid flag)."""

    T0 = 0.0
    T1 = 1.0
    V = [[0, 1], [0, 1], [0, 1]]
    state0 = np.zeros(2)

    def __init__(self, nt):
        self.nt = int(nt)
        self.nx = len(self.V)
        self.ny = len(self.state0)
        self.tau = (self.T1 - self.T0) / self.nt
        self.iterator = o.bounded_sum_iterator(self.V, 1, 1)
        self.x = np.zeros((self.nt, self.nx))
        self.df = np.zeros((self.nt, self.nx))
        self.state = np.zeros((self.nt, self.ny))
        self.adjoint = np.zeros((self.nt, self.ny))
        self.f = 0.0
        self.df_valid = False
        self.f_evals = 0
        self.df_evals = 0

    # to be provided: F(y, x) -> (ny,), Fy(y, x) -> (ny, ny), Fu(y, x) -> (ny, nx), G(y, x), Gy(y, x) -> (ny,)
    def eval_f(self):                                   # eval_f! (AbstractObjective.jl:81-91)
        self.f_evals += 1
        x = self.x
        state = self.state0.copy()
        fval = 0.5 * self.G(self.state0, x[0])          # ODEObjective.jl:130
        for i in range(self.nt):                        # :133
            state = state + self.tau * self.F(state, x[i])
            self.state[i] = state
            if i < self.nt - 1:
                fval += self.G(state, x[i + 1])
            else:
                fval += 0.5 * self.G(state, x[self.nt - 1])
        fval *= self.tau
        self.f = fval
        self.df_valid = False
        return fval

    def eval_df(self):                                  # eval_df! (AbstractObjective.jl:94-102)
        if self.df_valid:
            return
        self.df_evals += 1
        nt, tau = self.nt, self.tau
        Gy = self.Gy(self.state[nt - 1], self.x[nt - 1])
        self.adjoint[nt - 1] = -0.5 * tau * Gy          # ODEObjective.jl:166-167
        for i in range(nt - 1, 0, -1):                  # :169-173 (1-based i = nt-1 .. 1)
            y, xi = self.state[i - 1], self.x[i]
            Gy = self.Gy(y, xi)
            Fy = self.Fy(y, xi)
            self.adjoint[i - 1] = self.adjoint[i] + tau * (Fy.T @ self.adjoint[i] - Gy)
        self.df[:] = 0.0
        for i in range(1, nt + 1):                      # :177-183
            y = self.state0 if i == 1 else self.state[i - 2]
            Fu = self.Fu(y, self.x[i - 1])
            self.df[i - 1] -= Fu.T @ self.adjoint[i - 1]
        self.df_valid = True


class LVMObj(ODEObjective):                             # example_fishing.jl
    T1 = 12.0
    state0 = np.array([0.5, 0.7])
    v1 = np.array([0.2, 0.4, 0.01])
    v2 = np.array([0.1, 0.2, 0.1])

    def F(self, y, x):
        return np.array([y[0] * (1.0 - 1.0 * y[1] - 1.0 * np.sum(x * self.v1)),
                         y[1] * (-1.0 + 1.0 * y[0] - 1.0 * np.sum(x * self.v2))])

    def Fy(self, y, x):
        return np.array([[1.0 - y[1] - np.sum(x * self.v1), -y[0]],
                         [y[1], -1.0 + y[0] - np.sum(x * self.v2)]])

    def Fu(self, y, x):
        return np.array([y[0] * -1.0 * self.v1, y[1] * -1.0 * self.v2])

    def G(self, y, x):
        return 0.5 * (y[0] - 1.0) ** 2 + 0.5 * (y[1] - 1.0) ** 2

    def Gy(self, y, x):
        return np.array([y[0] - 1.0, y[1] - 1.0])


class VPOObj(ODEObjective):                             # example_vanderpol.jl
    T1 = 20.0
    state0 = np.array([1.0, 0.0])
    c = np.array([-1.0, 0.75, -2.0])

    def F(self, y, x):
        return np.array([y[1], (1 - y[0] ** 2) * y[1] * float(self.c @ x) - y[0]])

    def Fy(self, y, x):
        cx = float(self.c @ x)
        return np.array([[0.0, 1.0], [-2 * y[0] * y[1] * cx - 1, (1 - y[0] ** 2) * cx]])

    def Fu(self, y, x):
        return np.array([np.zeros(3), self.c * (1 - y[0] ** 2) * y[1]])

    def G(self, y, x):
        return y[0] ** 2 + y[1] ** 2

    def Gy(self, y, x):
        return np.array([2 * y[0], 2 * y[1]])


class DTMObj(ODEObjective):                             # example_doubletank.jl
    T1 = 10.0
    state0 = np.array([2.0, 2.0])
    c = np.array([1.0, 0.5, 2.0])
    k1, k2 = 2.0, 3.0

    def F(self, y, x):
        return np.array([float(self.c @ x) - math.sqrt(y[0]), math.sqrt(y[0]) - math.sqrt(y[1])])

    def Fy(self, y, x):
        return np.array([[-1 / (2 * math.sqrt(y[0])), 0.0], [1 / (2 * math.sqrt(y[0])), -1 / (2 * math.sqrt(y[1]))]])

    def Fu(self, y, x):
        return np.array([self.c, np.zeros(3)])

    def G(self, y, x):
        return self.k1 * (y[1] - self.k2) ** 2

    def Gy(self, y, x):
        return np.array([0.0, 2 * self.k1 * (y[1] - self.k2)])


@dataclass
class TRMParameters:                                    # multi-trust.jl:26-34
    beta: float = 0.001
    p: object = 1
    delta0: float = 1.0
    sigma: float = 0.5
    kmax: int = 40
    maxiter: int = 1000


MAIN = {                                                # multi-trust.jl:181-189
    "fishing": (LVMObj, TRMParameters(beta=0.0001, delta0=2.0, p=float("inf"))),
    "doubletank": (DTMObj, TRMParameters(beta=0.00001, delta0=2.0, p=float("inf"))),
    "vanderpol": (VPOObj, TRMParameters(beta=0.1, delta0=1.0, p=float("inf"))),
}


def start_control(obj, seed=0, jumps=None):
    """A fixed admissible piecewise-constant x0 (stands in for rand_func, HelpFunctions.jl:136-225)."""
    rng = np.random.default_rng(seed)
    lv = o.level_values(obj.V, obj.iterator)
    n = obj.nt
    jumps = n // 10 if jumps is None else jumps
    t = np.sort(rng.choice(np.arange(2, n + 1), size=jumps, replace=False))
    seg = np.searchsorted(t, np.arange(1, n + 1), side="right")
    return lv[rng.integers(0, lv.shape[0], size=jumps + 1)[seg]].astype(np.float64)


@dataclass
class History:
    rows: list = field(default_factory=list)            # (iter, k, radius, J, pred, ared, step)
    J: float = math.nan
    u: np.ndarray = None
    dp_calls: int = 0
    backtracks: int = 0


def TRM(obj, par, x0, bellman, eval_u, max_outer=None):
    """multi-trust.jl:53-170 with pluggable bellman_TRM!/eval_u_TRM! (reference signatures).

    bellman(df, u_old, B, beta, p, dt, nu, U, Phi, iterator); eval_u(u, u_old, U, Phi, B, nu)."""
    n, dt, it, nu = obj.nt, obj.tau, obj.iterator, obj.V
    beta, p = par.beta, par.p
    u = obj.x
    u[:] = x0
    u_old = u.copy()
    B = int(math.floor(par.delta0 / dt))                # :69
    U, Phi = o.alloc_tables(nu, n, B)                   # :71-77
    h = History()
    J = math.inf
    iter_ = 1
    stop = False
    J_old = obj.eval_f()                                # :83
    h.rows.append((0, 0, par.delta0, J_old + beta * o.TV_p(u, p), 0.0, 0.0, "Initial Value"))
    maxiter = par.maxiter if max_outer is None else min(par.maxiter, max_outer)
    while not stop and iter_ <= maxiter:                # :92
        dk = par.delta0
        k = 1
        ared, pred = 0.0, 1.0
        halved = False
        TV_old = o.TV_p(u, p)
        obj.eval_df()
        df = obj.df
        while ared < par.sigma * pred and k <= par.kmax:   # :105
            if halved:
                B_new = int(math.floor(dk / dt))
                eval_u(u, u_old, U, Phi, B_new, nu)
                h.backtracks += 1
            else:
                bellman(df, u_old, B, beta, p, dt, nu, U, Phi, it)
                eval_u(u, u_old, U, Phi, B, nu)
                h.dp_calls += 1
                h.backtracks += 1
            int_val = o.pred_integral(df, u_old, u, dt)   # :117-121
            TV_new = o.TV_p(u, p)
            J_new = obj.eval_f()
            pred = int_val + beta * (TV_old - TV_new)
            ared = J_old - J_new + beta * (TV_old - TV_new)
            if pred <= 0:
                J = J_old
                stop = True
                h.rows.append((iter_, k, dk, J + beta * TV_old, pred, ared, "optimal solution found"))
                break
            elif ared < par.sigma * pred:
                h.rows.append((iter_, k, dk, J_old + beta * TV_old, pred, ared, "bad step, halved"))
                dk = dk / 2
                halved = True
            else:
                u_old[:] = u
                J_old = J_new
                TV_old = TV_new
                J = J_new
                h.rows.append((iter_, k, dk, J + beta * TV_new, pred, ared, "good step"))
            k += 1
        iter_ += 1
    h.J = J + beta * o.TV_p(u, p)
    h.u = u.copy()
    return h


# --------------------------------------------------------------------------------------------------------
# Multi-start TRM in lock-step: the executable mirror of julia/MultiStartTRM.jl
# --------------------------------------------------------------------------------------------------------
def radius_ladder(par, dt):
    """The trial budgets one outer iteration can ask for, in the order the inner loop visits them:
    k = 1: B = floor(D0/dt) (multi-trust.jl:69); after the (k-1)-th halving: floor((D0/2^(k-1))/dt) (:109).
    Returns (distinct budgets in visiting order, index into them for k = 1..kmax)."""
    radii, index = [], []
    dk = par.delta0
    for _ in range(int(par.kmax)):
        b = int(math.floor(dk / dt))
        if not radii or radii[-1] != b:
            radii.append(b)
        index.append(len(radii) - 1)
        dk = dk / 2
    return radii, index


def oracle_solve_batched(nu, iterator, beta, p, dt, B):
    """A solve_batched callable backed by the CPU oracle (what the device's bb200_solve_batched computes)."""
    def solve(df_all, u_old_all, radii):
        A, n, M = df_all.shape
        u_all = np.zeros((A, len(radii), n, M))
        status = np.zeros((A, len(radii)), dtype=np.int32)
        for a in range(A):
            U, Phi = o.alloc_tables(nu, n, B)
            o.bellman_TRM(df_all[a], u_old_all[a], B, beta, p, dt, nu, U, Phi, iterator)
            for r, Bn in enumerate(radii):
                try:
                    o.eval_u_TRM(u_all[a, r], u_old_all[a], U, Phi, Bn, nu)
                except IndexError:
                    status[a, r] = 4
        return u_all, status
    return solve


def TRM_multistart(objs, par, x0s, solve_batched, max_outer=None, max_radii=16):
    """S trust-region loops (multi-trust.jl:92-163) advanced in lock-step.

    Per outer iteration the starts that are still running share ONE batched DP call: their gradients and current
    controls go in, and because a sweep over trial radii costs no extra DP (the reference re-runs only eval_u_TRM!
    after a rejected step, multi-trust.jl:109-110), the trajectories for the whole halving ladder come back at once.
    Each start then walks its own inner loop (:105-159) over its precomputed trajectories.  Every DP result is what the
    single-start loop would have computed, so the per-start histories equal S independent TRM runs.

    solve_batched(df_all (A, n, M), u_old_all (A, n, M), radii) -> (u_all (A, R, n, M), status (A, R))
    Returns a list of History, one per start."""
    S = len(objs)
    dt, nu = objs[0].tau, objs[0].V
    beta, p = par.beta, par.p
    radii, kidx = radius_ladder(par, dt)
    hs = [History() for _ in range(S)]
    us, u_olds, J_olds, Js = [], [], [], [math.inf] * S
    for s, obj in enumerate(objs):
        u = obj.x
        u[:] = x0s[s]
        us.append(u)
        u_olds.append(u.copy())
        J_olds.append(obj.eval_f())
        hs[s].rows.append((0, 0, par.delta0, J_olds[s] + beta * o.TV_p(u, p), 0.0, 0.0, "Initial Value"))
    stop = [False] * S
    iter_ = 1
    maxiter = par.maxiter if max_outer is None else min(par.maxiter, max_outer)
    while not all(stop) and iter_ <= maxiter:
        act = [s for s in range(S) if not stop[s]]
        TV_olds = {s: o.TV_p(us[s], p) for s in act}
        for s in act:
            objs[s].eval_df()
        df_all = np.stack([objs[s].df for s in act])
        uo_all = np.stack([u_olds[s] for s in act])
        u_all, status = solve_batched(df_all, uo_all, radii[:max_radii])   # the first max_radii rungs of the ladder
        for a, s in enumerate(act):
            obj, h = objs[s], hs[s]
            h.dp_calls += 1
            dk, k = par.delta0, 1
            ared, pred = 0.0, 1.0
            TV_old = TV_olds[s]
            lo, u_row, st_row = 0, u_all[a], status[a]     # radii[lo : lo + max_radii] are at hand for this start
            while ared < par.sigma * pred and k <= par.kmax:
                r = kidx[k - 1]
                if not (lo <= r < lo + max_radii):         # deeper than that (only when B >= 2^max_radii): one more DP
                    lo = r
                    u_more, st_more = solve_batched(df_all[a:a + 1], uo_all[a:a + 1], radii[lo:lo + max_radii])
                    u_row, st_row = u_more[0], st_more[0]
                if st_row[r - lo] != 0:
                    raise IndexError("backtrack visited a cell the DP never wrote")
                us[s][:] = u_row[r - lo]
                h.backtracks += 1
                int_val = o.pred_integral(obj.df, u_olds[s], us[s], dt)
                TV_new = o.TV_p(us[s], p)
                J_new = obj.eval_f()
                pred = int_val + beta * (TV_old - TV_new)
                ared = J_olds[s] - J_new + beta * (TV_old - TV_new)
                if pred <= 0:
                    Js[s] = J_olds[s]
                    stop[s] = True
                    h.rows.append((iter_, k, dk, Js[s] + beta * TV_old, pred, ared, "optimal solution found"))
                    break
                elif ared < par.sigma * pred:
                    h.rows.append((iter_, k, dk, J_olds[s] + beta * TV_old, pred, ared, "bad step, halved"))
                    dk = dk / 2
                else:
                    u_olds[s][:] = us[s]
                    J_olds[s] = J_new
                    TV_old = TV_new
                    Js[s] = J_new
                    h.rows.append((iter_, k, dk, Js[s] + beta * TV_new, pred, ared, "good step"))
                k += 1
        iter_ += 1
    for s in range(S):
        hs[s].J = Js[s] + beta * o.TV_p(us[s], p)
        hs[s].u = us[s].copy()
    return hs
