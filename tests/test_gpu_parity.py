"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI, must be
BIT-EXACT against the CPU oracle -- same value table (both slots), same argmin on every cell the DP wrote,
same control trajectory for every trial radius -- on the known-answer tests, the committed golden vectors,
tie-heavy and random instances of every example shape, and through size-independent properties at the
full BASELINE size."""
import importlib

import numpy as np
import pytest

from helpers import inf_list, kat_iterator, load_kats, load_seeded, objective_of, phi_admissible, random_instance

pytestmark = pytest.mark.gpu

# 0: the plan picks (single-CTA small-problem kernel or pipelined wavefront kernel); 1: one launch per stage;
# 4: pipelined wavefront kernel even for small problems
PATHS = [pytest.param(0, id="auto"), pytest.param(1, id="stage-kernels"), pytest.param(4, id="wavefront")]


def oracle_tables(o, nu, it, n, B, df, u_old, beta, p, dt, cost):
    U, Phi = o.alloc_tables(nu, n, B)
    n_upd = o.bellman_TRM(df, u_old, B, beta, p, dt, nu, U, Phi, it, cost=cost, threads=True)
    return U, Phi, n_upd


def check_against_oracle(m, o, nu, it, n, B, df, u_old, beta, p, dt, flags, radii=None, tune=None):
    plan = m.TRMPlan(nu, it, n, B, beta, p, dt, flags=flags)
    if tune:
        plan.tune(**tune)
    plan.bellman(df, u_old)
    U, Phi, n_upd = oracle_tables(o, nu, it, n, B, df, u_old, beta, p, dt, plan.cost)
    got_phi = plan.export_phi()
    if n == 1:
        # the reference never touches slot 2 for n == 1 (it keeps the caller's zeros); the device reports +Inf
        np.testing.assert_array_equal(got_phi[0], Phi[0])
    else:
        np.testing.assert_array_equal(got_phi, Phi)                  # bit-exact incl. +Inf cells
    assert plan.count_updates() == n_upd
    if n > 1:
        # every cell the reference wrote must hold the same index tuple; unwritten cells are `fill`
        ref = U.copy()
        got = plan.export_argmin(1, n, fill=0)
        np.testing.assert_array_equal(got, ref)
    for Bn in (radii if radii is not None else sorted({B, B // 2, B // 3, 1 if B >= 1 else 0, 0}, reverse=True)):
        u = np.zeros((n, len(nu)))
        u_ref = np.zeros((n, len(nu)))
        info = {}
        ps, bs, ks = plan.eval_u(u, Bn)
        o.eval_u_TRM(u_ref, u_old, U, Phi, Bn, nu, info=info)
        np.testing.assert_array_equal(u, u_ref)
        assert (ps, bs, int(plan.grid_offset[ks])) == (info["phi_star"], info["b_star"], info["g_star"])
    st = plan.stats()
    path = int(st["path"])
    assert (path == 0) if flags & 1 else ((path == 1) if (flags & 4 or tune) else path in (1, 2)), "the requested kernel path did not run"
    plan.close()
    return st


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("name", ["KAT-1", "KAT-2"])
def test_known_answers(gpu_lib, oracle, name, flags):
    m, o = gpu_lib, oracle
    kat = load_kats()[name]
    it = kat_iterator(o, kat)
    df, u_old = np.array(kat["df"]), np.array(kat["u_old"])
    plan = m.TRMPlan(kat["nu"], it, kat["n"], kat["B"], kat["beta"], kat["p"], kat["dt"], flags=flags)
    plan.bellman(df, u_old)
    pa = phi_admissible(o, plan.export_phi(), kat["nu"], it)
    np.testing.assert_array_equal(pa[0], inf_list(kat["phi_slot1"]))
    np.testing.assert_array_equal(pa[1], inf_list(kat["phi_slot2"]))
    g = o.grid_offsets(kat["nu"], it)
    Ur = plan.export_argmin(1, kat["n"], fill=0).reshape(kat["n"] - 1, -1, kat["B"] + 1, len(kat["nu"]))
    for i, b, k, tup in kat["U"]:
        assert Ur[i - 1, g[k], b].tolist() == tup
    for Bn, exp in kat["select"].items():
        u = np.zeros_like(u_old)
        ps, bs, ks = plan.eval_u(u, int(Bn))
        assert (ps, bs, ks) == (exp["phi"], exp["b"], exp["k"])
        np.testing.assert_array_equal(u, u_old if exp["u"] == "u_old" else np.array(exp["u"]))
    plan.close()


@pytest.mark.parametrize("flags", PATHS)
def test_committed_golden_vectors(gpu_lib, flags):
    m = gpu_lib
    for name, c in load_seeded().items():
        meta = c["meta"]
        plan = m.TRMPlan(meta["nu"], meta["iterator"], meta["n"], meta["B"], meta["beta"], meta["p"], meta["dt"],
                         flags=flags, cost=c["cost"])
        plan.bellman(c["df"], c["u_old"])
        np.testing.assert_array_equal(plan.export_phi(), c["Phi"], err_msg=name)
        assert plan.count_updates() == int(c["n_updates"][0])
        for r, u_exp in zip(c["radii"], c["u"]):
            u = np.zeros_like(c["u_old"])
            plan.eval_u(u, int(r))
            np.testing.assert_array_equal(u, u_exp, err_msg=f"{name} B'={r}")
        plan.close()


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("tie_heavy", [False, True])
def test_random_small_instances(gpu_lib, oracle, flags, tie_heavy):
    rng = np.random.default_rng(2024 + int(tie_heavy))
    for _ in range(40):
        inst = random_instance(rng, oracle, tie_heavy=tie_heavy, n_max=12, B_max=14)
        check_against_oracle(gpu_lib, oracle, inst["nu"], inst["it"], inst["n"], inst["B"], inst["df"], inst["u_old"],
                             inst["beta"], inst["p"], inst["dt"], flags,
                             radii=list(range(inst["B"], -1, -1)))


@pytest.mark.parametrize("flags", PATHS)
@pytest.mark.parametrize("kind,n,tie", [("fishing", 1024, False), ("vanderpol", 1024, True), ("doubletank", 1024, False),
                                        ("convolution", 1024, True), ("heat", 256, True), ("heat", 1024, False)])
def test_example_shapes(gpu_lib, oracle, kind, n, tie, flags):
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.example_shaped(kind, n=n, seed=31, tie_heavy=tie)
    check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, inst.beta,
                         inst.p, inst.dt, flags)


@pytest.mark.parametrize("tie", [False, True])
def test_synthetic_config4_shape_vs_oracle(gpu_lib, oracle, tie):
    """BASELINE config 4 (K=125, B=999) at an n the oracle finishes in seconds, bit for bit."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=160, B=999, seed=20251018, tie_heavy=tie)
    st = check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, inst.beta,
                              inst.p, inst.dt, 0, radii=[999, 499, 249, 124, 0])
    assert st["ctas"] >= 100  # the pipelined path really spreads over the GPU


@pytest.mark.parametrize("tune", [dict(ctas=8, jsplit=1), dict(ctas=20, jsplit=2), dict(variant=1), dict(ctas=20, variant=2),
                                  dict(variant=3), dict(variant=4, jsplit=3), dict(ctas=148, jsplit=6, variant=5),
                                  dict(variant=6), dict(variant=7, jsplit=1), dict(variant=8), dict(ctas=9, variant=9),
                                  dict(variant=10), dict(variant=11), dict(variant=12, jsplit=2), dict(variant=113),
                                  dict(variant=213, jsplit=3), dict(variant=413), dict(ctas=16, variant=214),
                                  dict(variant=115), dict(variant=316), dict(variant=617), dict(variant=218),
                                  dict(variant=119), dict(ctas=30, variant=220), dict(variant=321),
                                  dict(variant=822, jsplit=2),
                                  # pruned (branch-and-bound) scan: 4+4 rows fit the 212 budget rows on 27 CTAs etc.
                                  dict(variant=25), dict(variant=26), dict(variant=27), dict(variant=28),
                                  dict(variant=29), dict(variant=30), dict(variant=31), dict(variant=32),
                                  # two-zone slices (the upper CTAs own one row group less): 16 x 8 + 14 x 6, 26 x 8 + 1 x 4,
                                  # 6 x 4 + 94 x 2 and 106 x 2 rows = the 212 budget rows
                                  dict(ctas=30, variant=28), dict(ctas=27, variant=27), dict(ctas=100, variant=29),
                                  dict(ctas=106, variant=29)])
def test_wavefront_geometries(gpu_lib, oracle, tune):
    """Every tile variant / CTA count / j-split / scatter-warp count (variant + 100 * NS) of the pipelined kernel
    gives identical bits."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=70, B=211, seed=77, levels=4, M=3, tie_heavy=True)   # K = 64
    check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, inst.beta,
                         inst.p, inst.dt, 0, tune=tune)


@pytest.mark.parametrize("levels,M,B,tune", [(9, 2, 150, None), (9, 2, 150, dict(variant=413, jsplit=4)),   # K = 81 -> Kp = 96
                                           (33, 1, 97, None), (33, 1, 97, dict(variant=1)),                # K = 33 -> Kp = 64
                                           (7, 1, 40, dict(variant=215)), (10, 2, 333, None),              # K = 7, K = 100
                                           (10, 2, 333, dict(ctas=148, variant=614, jsplit=2)),
                                           (9, 2, 150, dict(variant=25)), (33, 1, 97, dict(variant=27)),   # pruned scan
                                           (7, 1, 40, dict(variant=29)), (10, 2, 333, dict(variant=26)),
                                           (5, 3, 999, dict(variant=25)), (5, 3, 999, dict(variant=27)),   # K = 125
                                           (5, 3, 999, dict(variant=28)), (5, 3, 999, dict(ctas=140, variant=28))])  # two-zone slices
def test_partially_filled_level_blocks(gpu_lib, oracle, levels, M, B, tune):
    """Level counts that are not multiples of 32 / 64: padded lanes, a half-filled last work unit in phase C, odd
    successor ranges in phase B -- all bits must still match."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=45, B=B, seed=levels * 100 + M, levels=levels, M=M, tie_heavy=True)
    check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, inst.beta,
                         inst.p, inst.dt, 4, tune=tune)


def test_large_budget_use_spans_many_slices(gpu_lib, oracle):
    """u_old far from most levels: pushes cross several CTA slices (halo depth D > 1)."""
    nu = [[0, 5, 10, 15]] * 2
    it = oracle.product_iterator(nu)
    rng = np.random.default_rng(3)
    n, B = 50, 120
    lv = oracle.level_values(nu, it)
    u_old = lv[rng.integers(0, len(it), size=n)].astype(np.float64)
    df = np.round(rng.standard_normal((n, 2)) * 4) / 4
    check_against_oracle(gpu_lib, oracle, nu, it, n, B, df, u_old, 0.25, 1, 0.5, 0, tune=dict(ctas=60))


def test_n_equals_one_and_two(gpu_lib, oracle):
    nu = [[0, 1, 2]]
    it = oracle.product_iterator(nu)
    for n in (1, 2, 3):
        for flags in (0, 1, 4):
            df = np.arange(1, n + 1, dtype=np.float64).reshape(n, 1) * -0.5
            u_old = np.ones((n, 1))
            plan = gpu_lib.TRMPlan(nu, it, n, 2, 0.25, 1, 1.0, flags=flags)
            plan.bellman(df, u_old)
            U, Phi, _ = oracle_tables(oracle, nu, it, n, 2, df, u_old, 0.25, 1, 1.0, plan.cost)
            got = plan.export_phi()
            # for n == 1 the reference never touches slot 2 (zeros from the allocation); compare slot 1 only
            np.testing.assert_array_equal(got[0], Phi[0])
            if n > 1:
                np.testing.assert_array_equal(got[1], Phi[1])
            u, u_ref = np.zeros((n, 1)), np.zeros((n, 1))
            plan.eval_u(u, 2)
            oracle.eval_u_TRM(u_ref, u_old, U, Phi, 2, nu)
            np.testing.assert_array_equal(u, u_ref)
            plan.close()


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_shortest_horizons_on_the_pruned_production_geometry(gpu_lib, oracle, n):
    """n = 1 .. 5 on the config-4 shape with the pruned tiles forced (148 CTAs, two-zone slices, three value-row buffers, no
    'scanned' hand-over): the halo / cost cursors of the comm warp and the barrier phases start and end within a few steps;
    three subproblems walked by ONE launch exercise the hand-over from one subproblem to the next."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    insts = [wl.synthetic(n=n, B=999, seed=40 + s, tie_heavy=(s == 1)) for s in range(3)]
    base = insts[0]
    plan = gpu_lib.TRMPlan(base.nu, base.iterator, n, base.B, base.beta, base.p, base.dt, flags=4, batch=3)
    plan.tune(variant=28)
    for s, inst in enumerate(insts):
        plan.upload(s, inst.df, inst.u_old)
    plan.bellman_resident(0, 3)
    plan.sync()
    st = plan.stats()
    assert int(st["path"]) == 1 and int(st["prune_block"]) == 4 and int(st["ctas"]) == 148
    for s, inst in enumerate(insts):
        U, Phi, n_upd = oracle_tables(oracle, inst.nu, inst.iterator, n, inst.B, inst.df, inst.u_old, base.beta, base.p, base.dt,
                                      plan.cost)
        got = plan.export_phi(s)
        np.testing.assert_array_equal(got[0], Phi[0])
        if n > 1:
            np.testing.assert_array_equal(got[1], Phi[1])
            np.testing.assert_array_equal(plan.export_argmin(1, n, fill=0, slot=s), U)
        for Bn in (999, 5, 0):
            u, ur = np.zeros((n, 3)), np.zeros((n, 3))
            try:
                plan.backtrack_resident(s, Bn); plan.sync()
                plan.download(s, u)
            except gpu_lib.StaleCellError:
                with pytest.raises(IndexError):
                    oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu)
                continue
            oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu)
            np.testing.assert_array_equal(u, ur)
    plan.close()


def test_error_behaviour(gpu_lib):
    m = gpu_lib
    nu = [[0, 1]]
    it = m.product_iterator(nu)
    plan = m.TRMPlan(nu, it, 4, 2, 0.5, 1, 1.0)
    with pytest.raises(m.BellmanB200Error):          # backtrack before any DP
        plan.eval_u(np.zeros((4, 1)))
    with pytest.raises(m.InexactError):              # Julia: InexactError at HelpFunctions.jl:37,57
        plan.bellman(np.zeros((4, 1)), np.array([[0.0], [0.5], [1.0], [0.0]]))
    with pytest.raises(m.InexactError):
        plan.bellman(np.zeros((4, 1)), np.array([[0.0], [np.nan], [1.0], [0.0]]))
    plan.bellman(np.zeros((4, 1)), np.zeros((4, 1)))
    with pytest.raises(m.BellmanB200Error):          # B_new > B
        plan.eval_u(np.zeros((4, 1)), 3)
    with pytest.raises(ValueError):
        plan.bellman(np.zeros((5, 1)), np.zeros((5, 1)))
    # inadmissible start (no level within the budget): the reference would read stale U
    with pytest.raises(m.StaleCellError):
        plan.bellman(np.zeros((4, 1)), np.full((4, 1), 9.0))
        plan.eval_u(np.zeros((4, 1)), 2)
    plan.close()


def test_drop_in_functions_and_write_back(gpu_lib, oracle):
    """bellman_TRM / eval_u_TRM with the reference's signatures, including filling the caller's U and Phi."""
    m, o = gpu_lib, oracle
    rng = np.random.default_rng(8)
    inst = random_instance(rng, o, tie_heavy=True, K_choice=36, n_max=9, B_max=9)
    inst["n"] = 9
    lv = o.level_values(inst["nu"], inst["it"])
    inst["u_old"] = lv[rng.integers(0, 36, size=9)].astype(np.float64)
    inst["df"] = np.round(rng.standard_normal((9, 2)) * 4) / 4
    U, Phi = o.alloc_tables(inst["nu"], 9, inst["B"])
    Ur, Phir = o.alloc_tables(inst["nu"], 9, inst["B"])
    m.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], U, Phi,
                  inst["it"], write_back=True)
    o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], Ur, Phir,
                  inst["it"])
    np.testing.assert_array_equal(Phi, Phir)
    np.testing.assert_array_equal(U, Ur)
    u, ur = np.zeros((9, 2)), np.zeros((9, 2))
    m.eval_u_TRM(u, inst["u_old"], U, Phi, inst["B"], inst["nu"])
    o.eval_u_TRM(ur, inst["u_old"], Ur, Phir, inst["B"], inst["nu"])
    np.testing.assert_array_equal(u, ur)


def test_repeated_iterations_reuse_the_plan(gpu_lib, oracle):
    """multi-trust.jl:105-114 call protocol: one DP per outer iteration, several shrinking radii, same tables."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.example_shaped("heat", n=128, seed=5, tie_heavy=True)
    plan = gpu_lib.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
    rng = np.random.default_rng(0)
    u_old = inst.u_old.copy()
    for it in range(4):
        df = np.round(rng.standard_normal(inst.df.shape) * 4) / 4
        U, Phi, _ = oracle_tables(oracle, inst.nu, inst.iterator, inst.n, inst.B, df, u_old, inst.beta, inst.p, inst.dt,
                                  plan.cost)
        u_first = np.zeros_like(u_old)
        ps, bs, ks = plan.solve(df, u_old, u_first)
        for Bn in (inst.B, inst.B // 2, inst.B // 4):
            u, ur = np.zeros_like(u_old), np.zeros_like(u_old)
            plan.eval_u(u, Bn)
            oracle.eval_u_TRM(ur, u_old, U, Phi, Bn, inst.nu)
            np.testing.assert_array_equal(u, ur)
            if Bn == inst.B:
                np.testing.assert_array_equal(u_first, ur)
        u_old = ur.copy()
    plan.close()


def test_batched_subproblems_and_radius_sweep(gpu_lib, oracle):
    """BASELINE config 5 in miniature: S subproblems x 4 radii, waves of `batch` slots."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    S, n, B = 7, 60, 99
    insts = [wl.synthetic(n=n, B=B, seed=20251018 + 2 * s, levels=3, M=3, tie_heavy=(s % 2 == 0)) for s in range(S)]
    base = insts[0]
    plan = gpu_lib.TRMPlan(base.nu, base.iterator, n, B, 0.25, 1, base.dt, batch=3)
    radii = [99, 49, 24, 12]
    u_all, phi, bs, ks = plan.solve_batched(np.stack([i.df for i in insts]), np.stack([i.u_old for i in insts]), radii)
    for s, inst in enumerate(insts):
        U, Phi, _ = oracle_tables(oracle, inst.nu, inst.iterator, n, B, inst.df, inst.u_old, 0.25, 1, inst.dt, plan.cost)
        for r, Bn in enumerate(radii):
            ur = np.zeros((n, 3))
            info = {}
            oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu, info=info)
            np.testing.assert_array_equal(u_all[s, r], ur)
            assert phi[s, r] == info["phi_star"] and bs[s, r] == info["b_star"]
    plan.close()


def test_next_rows_pred_integral_and_tv(gpu_lib, oracle):
    """SURVEY 8f N1/N3 on the resident arrays."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.example_shaped("heat", n=300, seed=9)
    plan = gpu_lib.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
    u = np.zeros_like(inst.u_old)
    plan.solve(inst.df, inst.u_old, u)
    assert plan.pred_integral() == oracle.pred_integral(inst.df, inst.u_old, u, inst.dt)
    assert plan.tv(1) == oracle.TV_p(u, 1)
    assert plan.tv(float("inf")) == oracle.TV_p(u, float("inf"))
    assert abs(plan.tv(2) - oracle.TV_p(u, 2)) <= 1e-12 * max(1.0, oracle.TV_p(u, 2))
    plan.close()


def test_full_size_properties(gpu_lib, oracle):
    """BASELINE config 4 at full width (K=125, B=999) and a long horizon the oracle cannot reach in seconds:
    size-independent properties -- used budget equals the selected row, optimum is monotone in the radius,
    B'=0 returns u_old, and the objective re-evaluated along the returned trajectory equals the table value."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=6000, B=999, seed=20251018)
    plan = gpu_lib.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, inst.beta, inst.p, inst.dt)
    plan.bellman(inst.df, inst.u_old)
    assert plan.count_updates() == oracle.count_updates(inst.u_old, inst.B, inst.nu, inst.iterator)
    d = dict(nu=inst.nu, it=inst.iterator, dt=inst.dt, df=inst.df)
    prev = None
    for Bn in (999, 499, 249, 124, 10, 0):
        u = np.zeros_like(inst.u_old)
        ps, bs, ks = plan.eval_u(u, Bn)
        used = int(np.abs(u - inst.u_old).sum())
        assert used == bs <= Bn
        if prev is not None:
            assert ps >= prev
        prev = ps
        if Bn == 0:
            np.testing.assert_array_equal(u, inst.u_old)
        obj = objective_of(oracle, u, d, plan.cost)
        assert abs(obj - ps) <= 1e-9 * max(1.0, abs(obj))
    # the two kernel paths agree bit for bit at this size as well
    plan2 = gpu_lib.TRMPlan(inst.nu, inst.iterator, 600, inst.B, inst.beta, inst.p, inst.dt, flags=1)
    plan3 = gpu_lib.TRMPlan(inst.nu, inst.iterator, 600, inst.B, inst.beta, inst.p, inst.dt)
    plan2.bellman(inst.df[:600], inst.u_old[:600]); plan3.bellman(inst.df[:600], inst.u_old[:600])
    np.testing.assert_array_equal(plan2.export_phi(), plan3.export_phi())
    u2, u3 = np.zeros((600, 3)), np.zeros((600, 3))
    assert plan2.eval_u(u2, 700) == plan3.eval_u(u3, 700)
    np.testing.assert_array_equal(u2, u3)
    for p_ in (plan, plan2, plan3):
        p_.close()


def test_uint16_argmin_and_large_level_sets(gpu_lib, oracle):
    """K > 255 needs the uint16 argmin table; its jump-cost table does not fit in shared memory, so the plan
    falls back to the per-stage kernels on its own (still CUDA, still bit-exact)."""
    nu = [[0, 1, 2, 3, 4, 5, 6]] * 3                       # K = 343
    it = oracle.product_iterator(nu)
    rng = np.random.default_rng(12)
    n, B = 6, 40
    lv = oracle.level_values(nu, it)
    u_old = lv[rng.integers(0, len(it), size=n)].astype(np.float64)
    df = np.round(rng.standard_normal((n, 3)) * 4) / 4
    plan = gpu_lib.TRMPlan(nu, it, n, B, 0.25, 1, 0.5)
    plan.bellman(df, u_old)
    st = plan.stats()
    assert int(st["arg_bytes"]) == 2 and int(st["path"]) == 0
    U, Phi, n_upd = oracle_tables(oracle, nu, it, n, B, df, u_old, 0.25, 1, 0.5, plan.cost)
    np.testing.assert_array_equal(plan.export_phi(), Phi)
    np.testing.assert_array_equal(plan.export_argmin(1, n, fill=0), U)
    for Bn in (40, 13, 0):
        u, ur = np.zeros((n, 3)), np.zeros((n, 3))
        plan.eval_u(u, Bn)
        oracle.eval_u_TRM(ur, u_old, U, Phi, Bn, nu)
        np.testing.assert_array_equal(u, ur)
    plan.close()


@pytest.mark.parametrize("levels,M,n,B", [(6, 3, 50, 211), (16, 2, 40, 97), (7, 3, 24, 60), (15, 2, 30, 333)])
@pytest.mark.parametrize("kind", ["random", "ties", "far", "flat"])
def test_wide_level_sets_run_the_pruned_per_stage_kernels(gpu_lib, oracle, levels, M, n, B, kind):
    """K = 216, 256, 343 (uint16 argmin), 225: the jump-cost table does not fit in shared memory, so the DP runs one launch
    per stage -- with the same branch-and-bound scan as the pipelined kernel, the successor axis in segments of 32 blocks
    (kernel_stage_pruned.cu).  Bits must match the oracle; data on which the bound test drops nothing ("flat") sends the plan
    back to the plain per-stage kernels, which must give the same bits again."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=n, B=B, seed=levels * 10 + M, levels=levels, M=M, tie_heavy=(kind == "ties"))
    beta = {"ties": 0.25, "random": 0.5, "flat": 0.0, "far": 50.0}[kind]
    df = inst.df if kind != "flat" else np.zeros_like(inst.df)
    plan = gpu_lib.TRMPlan(inst.nu, inst.iterator, n, B, beta, inst.p, inst.dt)
    U, Phi, n_upd = oracle_tables(oracle, inst.nu, inst.iterator, n, B, df, inst.u_old, beta, inst.p, inst.dt, plan.cost)
    for rep in range(2):                                   # the second DP may run on the plain kernels (adaptive switch)
        plan.bellman(df, inst.u_old)
        st = plan.stats()
        assert int(st["path"]) == 0 and int(st["arg_bytes"]) == (2 if inst.K > 255 else 1)
        if rep == 0:   # the first DP ran the pruned kernels (and may have switched the plan back to the plain ones afterwards)
            assert int(st["prune_block"]) == 4 or st["prune_switched_off"] >= 1
        np.testing.assert_array_equal(plan.export_phi(), Phi)
        np.testing.assert_array_equal(plan.export_argmin(1, n, fill=0), U)
        assert plan.count_updates() == n_upd
        for Bn in (B, B // 3, 0):
            u, ur = np.zeros((n, M)), np.zeros((n, M))
            plan.eval_u(u, Bn)
            oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu)
            np.testing.assert_array_equal(u, ur)
    if kind == "flat":
        assert plan.stats()["prune_switched_off"] >= 1 and int(plan.stats()["prune_block"]) == 0
    plan.close()


def test_resident_interface_multiple_slots_one_launch(gpu_lib, oracle):
    """bench.py's path: inputs uploaded once, several subproblems walked by ONE persistent launch."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    insts = [wl.synthetic(n=90, B=149, seed=100 + s, levels=4, M=3, tie_heavy=(s == 1)) for s in range(3)]
    base = insts[0]
    plan = gpu_lib.TRMPlan(base.nu, base.iterator, 90, 149, 0.25, 1, base.dt, batch=3)
    for s, inst in enumerate(insts):
        plan.upload(s, inst.df, inst.u_old)
    launches0 = plan.stats()["launches"]
    plan.bellman_resident(0, 3)
    plan.sync()
    assert plan.stats()["launches"] - launches0 == 4          # 3 prep kernels + one wavefront kernel
    for s, inst in enumerate(insts):
        U, Phi, n_upd = oracle_tables(oracle, inst.nu, inst.iterator, 90, 149, inst.df, inst.u_old, 0.25, 1, inst.dt, plan.cost)
        np.testing.assert_array_equal(plan.export_phi(s), Phi)
        assert plan.count_updates(s) == n_upd
        plan.backtrack_resident(s, 100)
        u, ur = np.zeros((90, 3)), np.zeros((90, 3))
        plan.download(s, u)
        oracle.eval_u_TRM(ur, inst.u_old, U, Phi, 100, inst.nu)
        np.testing.assert_array_equal(u, ur)
    plan.close()


def test_drop_in_rebuilds_the_plan_when_parameters_change(gpu_lib, oracle):
    """The reference's bellman_TRM! is stateless: calling it again on the SAME U/Phi with another beta, p, dt or
    iterator must answer for the new parameters (the plan used to be keyed on (n, M, B) only)."""
    m, o = gpu_lib, oracle
    rng = np.random.default_rng(21)
    nu = [[0, 1, 2], [0, 1]]
    it = o.product_iterator(nu)
    n, B = 12, 6
    lv = o.level_values(nu, it)
    u_old = lv[rng.integers(0, len(it), size=n)].astype(np.float64)
    df = np.round(rng.standard_normal((n, 2)) * 4) / 4
    U, Phi = o.alloc_tables(nu, n, B)
    it_small = [t for t in it if t[0] <= 2]
    u_old_small = np.minimum(u_old, [1.0, 1.0])
    for beta, p, dt, iterator, uo in ((0.25, 1, 1.0, it, u_old), (0.75, 1, 1.0, it, u_old), (0.75, 2, 1.0, it, u_old),
                                      (0.75, 2, 0.5, it, u_old), (0.75, float("inf"), 0.5, it, u_old),
                                      (0.75, 1, 0.5, it_small, u_old_small)):
        Ur, Phir = o.alloc_tables(nu, n, B)
        m.bellman_TRM(df, uo, B, beta, p, dt, nu, U, Phi, iterator, write_back=True)
        o.bellman_TRM(df, uo, B, beta, p, dt, nu, Ur, Phir, iterator)
        np.testing.assert_array_equal(Phi, Phir)
        for Bn in (B, B // 2):
            u, ur = np.zeros((n, 2)), np.zeros((n, 2))
            m.eval_u_TRM(u, uo, U, Phi, Bn, nu)
            o.eval_u_TRM(ur, uo, Ur, Phir, Bn, nu)
            np.testing.assert_array_equal(u, ur)


def test_drop_in_sequence_is_one_graph_replay_per_inner_iteration(gpu_lib, oracle):
    """multi-trust.jl:112-113 always calls eval_u_TRM!(.., B, ..) right after bellman_TRM!: the drop-in serves that
    pair with ONE bb200_solve (graph replay); smaller radii afterwards go to bb200_select_and_backtrack."""
    m, o = gpu_lib, oracle
    wl = importlib.import_module(m.__name__ + ".workloads")
    api = importlib.import_module(m.__name__ + ".api")
    inst = wl.example_shaped("heat", n=96, seed=12, tie_heavy=True)
    U, Phi = o.alloc_tables(inst.nu, inst.n, inst.B)
    Ur, Phir = o.alloc_tables(inst.nu, inst.n, inst.B)
    rng = np.random.default_rng(3)
    for k in range(3):
        df = np.round(rng.standard_normal(inst.df.shape) * 4) / 4
        m.bellman_TRM(df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, U, Phi, inst.iterator)
        o.bellman_TRM(df, inst.u_old, inst.B, inst.beta, inst.p, inst.dt, inst.nu, Ur, Phir, inst.iterator)
        ent = api._plans[id(U)]
        replays = ent.plan.stats()["graph_replays"]
        assert replays == k + 1
        for Bn in (inst.B, inst.B // 2, inst.B // 4):
            u, ur = np.zeros_like(inst.u_old), np.zeros_like(inst.u_old)
            m.eval_u_TRM(u, inst.u_old, U, Phi, Bn, inst.nu)
            o.eval_u_TRM(ur, inst.u_old, Ur, Phir, Bn, inst.nu)
            np.testing.assert_array_equal(u, ur)
        assert ent.plan.stats()["graph_replays"] == replays       # the radii did not re-run the DP


def test_non_integer_levels_are_rejected(gpu_lib):
    with pytest.raises(ValueError):
        gpu_lib.TRMPlan([[0, 0.5, 1]], [(1,), (2,), (3,)], 4, 2, 0.5, 1, 1.0)


@pytest.mark.parametrize("kind", ["ties", "random", "flat", "far"])
def test_pruned_scan_is_exact_and_reports_what_it_skipped(gpu_lib, oracle, kind):
    """The branch-and-bound scan (tile 25: 4 + 3 rows, blocks of 4 successors) must give the exhaustive scan's bits on
    inputs that stress the bound test: heavy ties, beta = 0 (every block has the same jump-cost minimum), huge jump costs
    (almost everything skipped), and it reports how many candidates it really evaluated."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=60, B=999, seed=5, tie_heavy=(kind == "ties"))
    beta = {"ties": 0.25, "random": 0.5, "flat": 0.0, "far": 50.0}[kind]
    df = inst.df if kind != "flat" else np.zeros_like(inst.df)
    st = check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, df, inst.u_old, beta, inst.p,
                              inst.dt, 4, radii=[999, 500, 3, 0], tune=dict(variant=25))
    assert st["prune_block"] == 4 and st["ctas"] == 143
    plan = gpu_lib.TRMPlan(inst.nu, inst.iterator, inst.n, inst.B, beta, inst.p, inst.dt, flags=4)
    plan.tune(variant=25)
    plan.bellman(df, inst.u_old)
    st = plan.stats()
    # padded successors are counted, skipped ones are not.  (How much is skipped depends on the data: at this short
    # horizon most value rows are still +Inf and bound nothing; the long-horizon skip rate is what bench.py reports.)
    assert 0 < st["executed_updates"] <= plan.count_updates() * 1.05
    plan.close()


@pytest.mark.parametrize("kind", ["tiny", "large", "beyond_float", "float_exact", "nan_inf"])
def test_pruned_bound_tests_in_float_are_safe_at_extreme_magnitudes(gpu_lib, oracle, kind):
    """The bound tests of the pruned scan run in FP32 with directed rounding.  Magnitudes far outside the float range (the
    bounds saturate to FLT_MAX / +-Inf), far below it (float denormals, slack terms of zero), values that convert to float
    exactly (directed rounding adds no margin, ties everywhere) and NaN / Inf stage costs must all give the reference's
    bits: a bound may only ever be too weak."""
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    inst = wl.synthetic(n=90, B=999, seed=17, tie_heavy=(kind == "float_exact"))
    df, beta = inst.df.copy(), inst.beta
    if kind == "tiny":
        df, beta = df * 1e-42, beta * 1e-42          # below the smallest normal float
    elif kind == "large":
        df, beta = df * 1e25, beta * 1e25
    elif kind == "beyond_float":
        df, beta = df * 1e200, beta * 1e200          # every value overflows float
    elif kind == "nan_inf":
        df[7, 1] = np.nan; df[23, 0] = np.inf; df[40, 2] = -np.inf; df[61, :] = np.nan
    for tune in (None, dict(variant=28), dict(variant=25)):
        try:
            st = check_against_oracle(gpu_lib, oracle, inst.nu, inst.iterator, inst.n, inst.B, df, inst.u_old, beta, inst.p,
                                      inst.dt, 4, radii=[999, 400, 0], tune=tune)
        except gpu_lib.StaleCellError:
            assert kind == "nan_inf"                 # a +Inf / NaN optimum: the library reports the reference's stale read
            continue
        if tune is not None:
            assert st["prune_block"] == 4


def _multi_case(gpu_lib, S=9, n=80, B=99):
    wl = importlib.import_module(gpu_lib.__name__ + ".workloads")
    insts = [wl.synthetic(n=n, B=B, seed=20251018 + 2 * s, levels=3, M=3, tie_heavy=(s % 2 == 0)) for s in range(S)]
    return insts, np.stack([i.df for i in insts]), np.stack([i.u_old for i in insts])


def test_batched_pipeline_reports_per_entry_status_and_one_wait_per_wave(gpu_lib, oracle):
    """bb200_solve_batched: waves of `batch` slots, ONE selection and ONE backtrack launch per wave (CTA per (slot,
    radius)), one host wait per wave; a subproblem with a non-integer u_old is reported for itself only."""
    insts, df_all, uo_all = _multi_case(gpu_lib)
    uo_all = uo_all.copy()
    uo_all[4, 7, 1] = 0.5                                   # InexactError for subproblem 4 only
    base = insts[0]
    plan = gpu_lib.TRMPlan(base.nu, base.iterator, base.n, base.B, 0.25, 1, base.dt, batch=4)
    radii = [99, 49, 24, 0]
    launches0 = plan.stats()["launches"]
    u_all, phi, bs, ks, status = plan.solve_batched(df_all, uo_all, radii, strict=False)
    st = plan.stats()
    assert st["batch_waves"] == 3 and st["batch_syncs"] == 3                # ceil(9 / 4) waves, one wait each
    assert st["launches"] - launches0 == 3 * 3 + 9                           # per wave: DP + selection + backtrack; + one prep per subproblem
    assert (status[4] == gpu_lib._lib.ERR_INEXACT).all() and (np.delete(status, 4, axis=0) == 0).all()
    for s, inst in enumerate(insts):
        if s == 4:
            continue
        U, Phi, _ = oracle_tables(oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, 0.25, 1, inst.dt, plan.cost)
        for r, Bn in enumerate(radii):
            ur = np.zeros((inst.n, 3))
            info = {}
            oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu, info=info)
            np.testing.assert_array_equal(u_all[s, r], ur)
            assert phi[s, r] == info["phi_star"] and bs[s, r] == info["b_star"]
    with pytest.raises(gpu_lib.InexactError):                               # strict mode raises the worst entry
        plan.solve_batched(df_all, uo_all, radii)
    plan.close()


def test_multi_gpu_behind_the_c_abi(gpu_lib, oracle):
    """bb200_multi_*: one ccall drives every visible GPU -- subproblem s on device s mod G, NCCL all-gather of the
    16-byte best-candidate records inside the library, the winner's trajectory returned.  With one GPU the same entry
    point runs without NCCL."""
    G = min(gpu_lib.device_count(), 8)
    insts, df_all, uo_all = _multi_case(gpu_lib, S=11)
    base = insts[0]
    mp = gpu_lib.MultiPlan(list(range(G)), base.nu, base.iterator, base.n, base.B, 0.25, 1, base.dt, batch_per_device=2)
    radii = [99, 49, 24]
    res = mp.solve_batched(df_all, uo_all, radii)
    best = (np.inf, -1, -1)
    for s, inst in enumerate(insts):
        U, Phi, _ = oracle_tables(oracle, inst.nu, inst.iterator, inst.n, inst.B, inst.df, inst.u_old, 0.25, 1, inst.dt, mp.cost)
        for r, Bn in enumerate(radii):
            ur = np.zeros((inst.n, 3))
            info = {}
            oracle.eval_u_TRM(ur, inst.u_old, U, Phi, Bn, inst.nu, info=info)
            np.testing.assert_array_equal(res["u"][s, r], ur)
            assert res["phi"][s, r] == info["phi_star"]
            if info["phi_star"] < best[0]:
                best = (info["phi_star"], s, r)
    assert (res["best_value"], res["best_subproblem"], res["best_radius"]) == best
    np.testing.assert_array_equal(res["u_best"], res["u"][best[1], best[2]])
    # without the trajectories: the winner's u is recomputed on its device
    res2 = mp.solve_batched(df_all, uo_all, radii, want_u=False)
    np.testing.assert_array_equal(res2["u_best"], res["u_best"])
    assert mp.stats()["devices"] == G
    if G > 1:
        assert gpu_lib.nccl_version() > 0
    mp.close()
