"""TRM objective-history parity (BASELINE configs 1-2): the reference's outer loop (restated in
oracle/trm_harness.py) is run with the oracle's DP and with the device's DP on the same start control; the log
tables (Iter, k, radius, J, pred, ared, step) and the final controls must be IDENTICAL, because every DP result is."""
import numpy as np
import pytest

from oracle import trm_harness as th


def test_harness_runs_on_cpu_and_descends(oracle):
    cls, par = th.MAIN["fishing"]
    obj = cls(128)
    x0 = th.start_control(obj, seed=1)
    h = th.TRM(obj, par, x0, oracle.bellman_TRM, oracle.eval_u_TRM, max_outer=4)
    assert h.dp_calls >= 1 and len(h.rows) >= 2
    goods = [r for r in h.rows if r[6] == "good step"]
    assert all(r[5] >= par.sigma * r[4] > 0 for r in goods)          # accepted steps satisfy ared >= sigma*pred
    js = [h.rows[0][3]] + [r[3] for r in goods]
    assert all(b <= a for a, b in zip(js, js[1:]))                    # J + beta*TV never increases
    # B' = floor(radius/dt) shrinks on every bad step of one outer iteration
    assert obj.f_evals >= 1 + len(h.rows) - 1 and obj.df_evals >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("problem", ["fishing", "vanderpol", "doubletank"])
def test_objective_history_identical(gpu_lib, oracle, problem):
    cls, par = th.MAIN[problem]
    n = 1024                                                          # main()'s default discretisation
    x0 = th.start_control(cls(n), seed=2)
    obj_ref, obj_dev = cls(n), cls(n)
    h_ref = th.TRM(obj_ref, par, x0, oracle.bellman_TRM, oracle.eval_u_TRM, max_outer=5)
    h_dev = th.TRM(obj_dev, par, x0, gpu_lib.bellman_TRM, gpu_lib.eval_u_TRM, max_outer=5)
    assert h_ref.rows == h_dev.rows
    assert h_ref.J == h_dev.J
    np.testing.assert_array_equal(h_ref.u, h_dev.u)
    assert (h_ref.dp_calls, h_ref.backtracks) == (h_dev.dp_calls, h_dev.backtracks)
    assert h_ref.dp_calls >= 1


def _multistart_case(problem, S, n, seed0):
    cls, par = th.MAIN[problem]
    x0s = [th.start_control(cls(n), seed=seed0 + s) for s in range(S)]
    return cls, par, x0s


def test_multistart_lockstep_equals_independent_runs(oracle):
    """S TRM loops in lock-step with one batched DP per outer iteration (radius ladder precomputed) give,
    start by start, the history of S independent single-start runs.  CPU: the batched solver is the oracle."""
    cls, par, x0s = _multistart_case("fishing", 3, 96, 11)
    ref = [th.TRM(cls(96), par, x0, oracle.bellman_TRM, oracle.eval_u_TRM, max_outer=3) for x0 in x0s]
    objs = [cls(96) for _ in x0s]
    B = int(np.floor(par.delta0 / objs[0].tau))
    solver = th.oracle_solve_batched(objs[0].V, objs[0].iterator, par.beta, par.p, objs[0].tau, B)
    got = th.TRM_multistart(objs, par, x0s, solver, max_outer=3)
    for h_ref, h in zip(ref, got):
        assert h_ref.rows == h.rows and h_ref.J == h.J
        np.testing.assert_array_equal(h_ref.u, h.u)
    # a ladder deeper than the resident radii (max_radii = 2) only costs extra DP calls, never a different result
    objs2 = [cls(96) for _ in x0s]
    got2 = th.TRM_multistart(objs2, par, x0s, solver, max_outer=3, max_radii=2)
    for h_ref, h in zip(ref, got2):
        assert h_ref.rows == h.rows
        np.testing.assert_array_equal(h_ref.u, h.u)


def test_radius_ladder_matches_the_inner_loop(oracle):
    cls, par = th.MAIN["fishing"]
    radii, kidx = th.radius_ladder(par, cls(1024).tau)
    assert radii[0] == int(np.floor(par.delta0 / cls(1024).tau)) and radii[-1] == 0
    assert all(a > b for a, b in zip(radii, radii[1:])) and len(kidx) == par.kmax and kidx[0] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("problem", ["fishing", "doubletank"])
def test_multistart_on_device_equals_independent_oracle_runs(gpu_lib, oracle, problem):
    """The same through bb200_solve_batched (one persistent DP launch + one selection + one backtrack launch per wave)."""
    S, n = 5, 256
    cls, par, x0s = _multistart_case(problem, S, n, 31)
    ref = [th.TRM(cls(n), par, x0, oracle.bellman_TRM, oracle.eval_u_TRM, max_outer=4) for x0 in x0s]
    objs = [cls(n) for _ in x0s]
    B = int(np.floor(par.delta0 / objs[0].tau))
    plan = gpu_lib.TRMPlan(objs[0].V, objs[0].iterator, n, B, par.beta, par.p, objs[0].tau, batch=3)

    def solver(df_all, u_old_all, radii):
        u_all, _, _, _, status = plan.solve_batched(df_all, u_old_all, radii, strict=False)
        return u_all, status

    got = th.TRM_multistart(objs, par, x0s, solver, max_outer=4)
    for h_ref, h in zip(ref, got):
        assert h_ref.rows == h.rows and h_ref.J == h.J
        np.testing.assert_array_equal(h_ref.u, h.u)
    plan.close()
