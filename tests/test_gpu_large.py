"""Bit-parity at the sizes BASELINE names: config 4 against the oracle at n = 2 000 (value table,
the WHOLE reference-layout argmin table, trajectories of five radii), the heat-shaped instance at the refined n = 8 192
against the oracle, and the full-size n = 100 000 run of the pipelined kernel against the one-launch-per-stage kernels.
Needs the GPU box's memory (the reference's Int64 tuple table is 6 - 8 GB at these sizes)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def compare_with_oracle(m, o, inst, radii, tune=None, flags=0):
    plan = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=flags)
    if tune:
        plan.tune(**tune)
    plan.bellman(inst.df, inst.u_old)
    U, Phi = o.alloc_tables(inst.nu, inst.n, inst.B)
    n_upd = o.bellman_TRM(inst.df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi, inst.iterator,
                          cost=plan.cost, threads=True)
    assert plan.count_updates() == n_upd
    assert np.array_equal(plan.export_phi().view(np.int64), Phi.view(np.int64)), "value table differs"
    step = 256                                                   # the argmin table in chunks of stages (bounded host memory)
    for i0 in range(1, inst.n, step):
        i1 = min(inst.n, i0 + step)
        assert np.array_equal(plan.export_argmin(i0, i1, fill=0), U[i0 - 1:i1 - 1]), f"argmin table differs in stages [{i0}, {i1})"
    for Bn in radii:
        u, ur = np.zeros_like(inst.u_old), np.zeros_like(inst.u_old)
        info = {}
        ps, bs, ks = plan.eval_u(u, Bn)
        o.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu, info=info)
        assert np.array_equal(u, ur), f"trajectory differs at B'={Bn}"
        assert (ps, bs, int(plan.grid_offset[ks])) == (info["phi_star"], info["b_star"], info["g_star"])
    st = plan.stats()
    plan.close()
    return st


@pytest.mark.parametrize("tune", [None, dict(variant=25)], ids=["auto", "pruned"])
@pytest.mark.parametrize("tie", [False, True], ids=["random", "ties"])
def test_config4_n2000_against_oracle(gpu_lib, oracle, tie, tune):
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=2000, B=999, seed=20251018, tie_heavy=tie)
    st = compare_with_oracle(gpu_lib, oracle, inst, [999, 499, 249, 124, 0], tune=tune)
    assert st["path"] == 1 and st["ctas"] >= 100


@pytest.mark.parametrize("n", [8192])
def test_heat_shaped_refined_against_oracle(gpu_lib, oracle, n):
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.example_shaped("heat", n=n, seed=3)
    st = compare_with_oracle(gpu_lib, oracle, inst, [inst.B, inst.B // 2, inst.B // 4, 7, 0])
    assert st["path"] == 1


def test_full_size_pipelined_against_per_stage_kernels(gpu_lib):
    """The n = 100 000 instance of the headline number: the persistent flag-synchronised pipeline against the plain
    one-launch-per-stage kernels (no inter-CTA synchronisation at all) -- exit state and trajectories, bit for bit."""
    m = gpu_lib
    wl = importlib.import_module(m.__name__ + ".workloads")
    inst = wl.synthetic(n=100_000, B=999, seed=20251018)
    fast = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
    ref = m.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt, flags=1)
    fast.bellman(inst.df, inst.u_old)
    ref.bellman(inst.df, inst.u_old)
    assert fast.stats()["path"] == 1 and ref.stats()["path"] == 0
    assert fast.count_updates() == ref.count_updates()
    assert np.array_equal(fast.export_phi().view(np.int64), ref.export_phi().view(np.int64))
    for i0 in (1, 50_000, 99_900):                               # argmin table: three windows of 64 stages
        assert np.array_equal(fast.export_argmin(i0, i0 + 64), ref.export_argmin(i0, i0 + 64))
    ua, ub = np.zeros_like(inst.u_old), np.zeros_like(inst.u_old)
    for Bn in (999, 499, 249, 124, 0):
        assert fast.eval_u(ua, Bn) == ref.eval_u(ub, Bn)
        assert np.array_equal(ua, ub)
    fast.close(); ref.close()
