"""CPU tests of the host-side logic: iterator mirror, table flattening, workload generators and the
multi-rank (world_size 2, gloo) sharding + best-candidate reduction."""
import importlib
import os
import socket

import numpy as np
import pytest


def test_iterators_mirror_the_oracle(mioc, oracle):
    for nu in ([[0, 1]] * 3, [[-2, -1, 0, 1, 2]], [[0, 1, 2], [5, 7]], [[0, 1, 2, 3, 4]] * 3):
        assert mioc.product_iterator(nu) == oracle.product_iterator(nu)
    nu = [[0, 1]] * 3
    assert mioc.bounded_sum_iterator(nu, 1, 1) == oracle.bounded_sum_iterator(nu, 1, 1)
    assert mioc.bounded_sum_iterator(nu, 0, 2) == oracle.bounded_sum_iterator(nu, 0, 2)


def test_flatten_and_jump_cost_match_oracle(mioc, oracle):
    nu = [[0, 1, 2, 3, 4, 5]] * 2
    it = mioc.product_iterator(nu)
    lv, goff, dims = mioc.flatten(nu, it)
    np.testing.assert_array_equal(lv, oracle.level_values(nu, it))
    np.testing.assert_array_equal(goff, oracle.grid_offsets(nu, it))
    assert dims.tolist() == [6, 6]
    for beta, p in ((1e-3, 2), (0.5, 1), (1e-4, float("inf")), (0.3, 3)):
        np.testing.assert_array_equal(mioc.jump_cost_table(beta, p, lv), oracle.jump_cost_table(beta, p, nu, it))
    with pytest.raises(ValueError):
        mioc.jump_cost_table(0.1, 0, lv)


def test_workloads_are_deterministic_and_admissible(mioc, oracle):
    wl = importlib.import_module(mioc.__name__ + ".workloads")
    a = wl.synthetic(n=500, B=99, seed=20251018)
    b = wl.synthetic(n=500, B=99, seed=20251018)
    np.testing.assert_array_equal(a.df, b.df)
    np.testing.assert_array_equal(a.u_old, b.u_old)
    assert a.K == 125 and a.M == 3
    lv = oracle.level_values(a.nu, a.iterator)
    assert all((lv == row).all(axis=1).any() for row in a.u_old[:50])
    assert 10 <= int((np.abs(np.diff(a.u_old, axis=0)).sum(axis=1) > 0).sum()) <= 50
    f = wl.example_shaped("fishing")
    assert (f.n, f.K, f.B) == (1024, 3, 170)
    h = wl.example_shaped("heat")
    assert (h.K, h.B, h.p) == (36, 204, 2)
    assert wl.example_shaped("vanderpol").B == 51 and wl.example_shaped("doubletank").B == 204
    assert wl.example_shaped("convolution").B == 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, ret):
    import torch.distributed as dist
    import mioc_b200
    d = importlib.import_module(mioc_b200.__name__ + ".distributed")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = 7
    values = np.array([5.0, 2.0, 9.0, 2.0, 2.5, 8.0, 3.0])      # tie between subproblems 1 and 3
    mine = d.shard(S, rank, world)
    bv, bi = d.local_best(values[mine], np.array(mine))
    gv, gi = d.best_candidate(bv, bi)
    owner = gi % world
    u = np.full((4, 2), float(rank))
    got = d.fetch_winner_control(u, owner, 4, 2)
    ret[rank] = (mine, gv, gi, float(got[0, 0]))
    dist.destroy_process_group()


def test_two_rank_sharding_and_best_candidate(mioc):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_rank_main, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][0] == [0, 2, 4, 6] and ret[1][0] == [1, 3, 5]
    for r in (0, 1):
        assert ret[r][1:3] == (2.0, 1)          # smallest value, then smallest global index
        assert ret[r][3] == 1.0                 # winner's control came from rank 1
