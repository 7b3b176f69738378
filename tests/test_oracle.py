"""CPU tests of the oracle (oracle/bellman_oracle.c) against the committed known answers, a brute-force
enumerator and the structural properties of SURVEY.md section 8c.  The reference ships no tests for this
path, so these are the pins the oracle has: PARITY UNPINNED against real Julia."""
import numpy as np
import pytest

from helpers import inf_list, kat_iterator, load_kats, load_seeded, objective_of, phi_admissible, random_instance


def run_oracle(o, inst, with_U=True):
    U, Phi = o.alloc_tables(inst["nu"], inst["n"], inst["B"])
    cost = o.jump_cost_table(inst["beta"], inst["p"], inst["nu"], inst["it"])
    n_upd = o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"],
                          U if with_U else None, Phi, inst["it"], cost=cost)
    return U, Phi, cost, n_upd


@pytest.mark.parametrize("name", ["KAT-1", "KAT-2"])
def test_known_answers(oracle, name):
    o = oracle
    kat = load_kats()[name]
    it = kat_iterator(o, kat)
    inst = dict(nu=kat["nu"], it=it, n=kat["n"], B=kat["B"], df=np.array(kat["df"]), u_old=np.array(kat["u_old"]),
                beta=kat["beta"], p=kat["p"], dt=kat["dt"])
    U, Phi, cost, _ = run_oracle(o, inst)
    pa = phi_admissible(o, Phi, kat["nu"], it)
    np.testing.assert_array_equal(pa[0], inf_list(kat["phi_slot1"]))
    np.testing.assert_array_equal(pa[1], inf_list(kat["phi_slot2"]))
    # inadmissible grid cells stay +Inf in both slots
    g = o.grid_offsets(kat["nu"], it)
    B1 = kat["B"] + 1
    rest = np.delete(Phi.reshape(2, -1, B1), g, axis=1)
    assert np.isinf(rest).all()
    Ur = U.reshape(kat["n"] - 1, -1, B1, len(kat["nu"]))
    for i, b, k, tup in kat["U"]:
        assert Ur[i - 1, g[k], b].tolist() == tup
    if "U_all_others_point_to" in kat:
        named = {(i, b, k) for i, b, k, _ in kat["U"]}
        for i in range(1, kat["n"]):
            for k in range(len(it)):
                for b in range(B1):
                    cell = Ur[i - 1, g[k], b]
                    if (i, b, k) in named or not cell.any():
                        continue
                    assert cell.tolist() == kat["U_all_others_point_to"]
    for Bn, exp in kat["select"].items():
        u = np.zeros_like(inst["u_old"])
        info = {}
        o.eval_u_TRM(u, inst["u_old"], U, Phi, int(Bn), kat["nu"], info=info)
        assert info["b_star"] == exp["b"] and info["g_star"] == g[exp["k"]] and info["phi_star"] == exp["phi"]
        exp_u = inst["u_old"] if exp["u"] == "u_old" else np.array(exp["u"])
        np.testing.assert_array_equal(u, exp_u)


def test_iterator_order(oracle):
    o = oracle
    nu = [[0, 1]] * 3
    # Iterators.product varies the FIRST index fastest (AdmissibleIterators.jl:14-17)
    assert o.product_iterator([[7, 8], [1, 2, 3]])[:3] == [(1, 1), (2, 1), (1, 2)]
    it = o.bounded_sum_iterator(nu, 1, 1)
    assert it == [(2, 1, 1), (1, 2, 1), (1, 1, 2)]
    assert o.level_values(nu, it).tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
    assert (np.diff(o.grid_offsets(nu, it)) > 0).all()  # admissible order == ascending grid offset


def test_tv_p_docstring_values(oracle):
    # HelpFunctions.jl:236-248: u = [1 -1 1; 3 3 0; 2 2 1] (M x n) -> memory order (n, M)
    u = np.array([[1, 3, 2], [-1, 3, 2], [1, 0, 1]], dtype=np.float64)
    assert oracle.TV_p(u, 1) == 8
    assert oracle.TV_p(u, 2) == 5.741657386773941
    assert oracle.TV_p(u, float("inf")) == 5


def test_p_inf_jump_cost_is_constant(oracle):
    # with p = Inf the DP's jump cost is beta for EVERY pair, including j == l
    nu = [[0, 1, 2], [0, 3]]
    it = oracle.product_iterator(nu)
    c = oracle.jump_cost_table(0.125, float("inf"), nu, it)
    assert (c == 0.125).all()
    c1 = oracle.jump_cost_table(0.5, 1, nu, it)
    lv = oracle.level_values(nu, it)
    np.testing.assert_array_equal(c1, 0.5 * np.abs(lv[:, None, :] - lv[None, :, :]).sum(-1))


def test_inexact_u_old_raises(oracle):
    nu = [[0, 1]]
    it = oracle.product_iterator(nu)
    U, Phi = oracle.alloc_tables(nu, 3, 2)
    with pytest.raises(oracle.InexactError):
        oracle.bellman_TRM(np.zeros((3, 1)), np.array([[0.0], [0.5], [1.0]]), 2, 0.1, 1, 1.0, nu, U, Phi, it)


def test_dp_optimum_equals_brute_force(oracle):
    o = oracle
    rng = np.random.default_rng(1234)
    done = 0
    while done < 30:
        inst = random_instance(rng, o, n_max=4, B_max=5, K_choice=rng.choice([3, 5, 6]))
        U, Phi, cost, _ = run_oracle(o, inst)
        best, _ = o.brute_force(inst["df"], inst["u_old"], inst["B"], inst["dt"], inst["nu"], inst["it"], cost)
        pa = phi_admissible(o, Phi, inst["nu"], inst["it"])
        assert abs(pa[0].min() - best) <= 1e-9 * max(1.0, abs(best))
        done += 1


@pytest.mark.parametrize("tie_heavy", [False, True])
def test_structural_properties(oracle, tie_heavy):
    """(ii) poisoned table, (iii) used budget == selected row, (iv) monotone in B', (v) B'=0 => u = u_old."""
    o = oracle
    rng = np.random.default_rng(99 if tie_heavy else 98)
    for _ in range(100):
        inst = random_instance(rng, o, tie_heavy=tie_heavy)
        U, Phi = o.alloc_tables(inst["nu"], inst["n"], inst["B"])
        U[...] = -7  # poison: an index the backtrack must never consume
        cost = o.jump_cost_table(inst["beta"], inst["p"], inst["nu"], inst["it"])
        o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], U, Phi,
                      inst["it"], cost=cost)
        prev = -np.inf
        for Bn in range(0, inst["B"] + 1):
            u = np.zeros_like(inst["u_old"])
            info = {}
            o.eval_u_TRM(u, inst["u_old"], U, Phi, Bn, inst["nu"], info=info)  # raises on a poisoned cell
            used = int(np.abs(u - inst["u_old"]).sum())
            assert used == info["b_star"] <= Bn
            if Bn == 0:
                np.testing.assert_array_equal(u, inst["u_old"])
            if Bn > 0:
                assert info["phi_star"] <= prev
            prev = info["phi_star"]
            obj = objective_of(o, u, inst, cost)
            assert abs(obj - info["phi_star"]) <= 1e-9 * max(1.0, abs(obj))


def test_compact_table_matches_reference_table(oracle):
    o = oracle
    rng = np.random.default_rng(5)
    for _ in range(20):
        inst = random_instance(rng, o, tie_heavy=True)
        if inst["n"] < 2:
            continue
        U, Phi = o.alloc_tables(inst["nu"], inst["n"], inst["B"])
        K = len(inst["it"])
        argk = np.zeros((inst["n"] - 1, K, inst["B"] + 1), dtype=np.int16)
        o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], U, Phi,
                      inst["it"], argk=argk)
        g = o.grid_offsets(inst["nu"], inst["it"])
        Ur = U.reshape(inst["n"] - 1, -1, inst["B"] + 1, len(inst["nu"]))
        tuples = np.array(inst["it"])
        for i in range(inst["n"] - 1):
            for k in range(K):
                for b in range(inst["B"] + 1):
                    a = argk[i, k, b]
                    if a < 0:
                        assert not Ur[i, g[k], b].any()
                    else:
                        assert Ur[i, g[k], b].tolist() == tuples[a].tolist()


def test_openmp_build_is_identical(oracle):
    o = oracle
    rng = np.random.default_rng(77)
    inst = random_instance(rng, o, K_choice=36, n_max=8, B_max=9)
    inst["n"] = 8
    inst["df"] = rng.standard_normal((8, 2)); inst["u_old"] = np.zeros((8, 2))
    U1, P1 = o.alloc_tables(inst["nu"], 8, inst["B"]); U2, P2 = o.alloc_tables(inst["nu"], 8, inst["B"])
    n1 = o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], U1, P1, inst["it"])
    n2 = o.bellman_TRM(inst["df"], inst["u_old"], inst["B"], inst["beta"], inst["p"], inst["dt"], inst["nu"], U2, P2, inst["it"], threads=True)
    assert n1 == n2 == o.count_updates(inst["u_old"], inst["B"], inst["nu"], inst["it"])
    np.testing.assert_array_equal(P1, P2)
    np.testing.assert_array_equal(U1, U2)


def test_oracle_reproduces_committed_vectors(oracle):
    o = oracle
    for name, c in load_seeded().items():
        m = c["meta"]
        U, Phi = o.alloc_tables(m["nu"], m["n"], m["B"])
        n_upd = o.bellman_TRM(c["df"], c["u_old"], m["B"], m["beta"], m["p"], m["dt"], m["nu"], U, Phi, m["iterator"],
                              cost=c["cost"])
        np.testing.assert_array_equal(Phi, c["Phi"], err_msg=name)
        assert n_upd == int(c["n_updates"][0])
        for r, u_exp in zip(c["radii"], c["u"]):
            u = np.zeros_like(c["u_old"])
            o.eval_u_TRM(u, c["u_old"], U, Phi, int(r), m["nu"])
            np.testing.assert_array_equal(u, u_exp, err_msg=f"{name} B'={r}")
