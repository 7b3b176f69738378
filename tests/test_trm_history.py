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
