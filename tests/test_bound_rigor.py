"""Host-side mirror of the pruned scan's bound tests (csrc/pruned_scan.cuh), checked for SOUNDNESS without a GPU.

The device drops a block of successors when a bound test proves that none of its candidates can be a cell's minimum or tie
with it.  The tests run in FP32 with directed rounding; this file re-states them operation by operation with exact
(rational) arithmetic deciding every directed rounding, and asserts on random and adversarial inputs (float denormals, values
beyond the float range, float-exact ties, infinities) that

    a block the ROW test or the LEVEL test drops holds no candidate v = fl64(fl64(s + c) + P) with v <= UB,

where UB is the exactly evaluated candidate the device uses as upper bound.  A bound may be weak; it must never be wrong.
(The GPU parity tests check the same thing end to end; this one pins the arithmetic argument itself.)"""
import math
from fractions import Fraction

import numpy as np
import pytest

F32 = np.float32
INF = float("inf")


def _fr(x):
    return Fraction(float(x))


def rd32(x):
    """largest float32 <= x (x a finite or infinite double)"""
    if math.isnan(x):
        return F32(np.nan)
    with np.errstate(over="ignore"):
        f = F32(x)
    if float(f) > x:
        f = np.nextafter(f, F32(-np.inf))
    return f


def ru32(x):
    if math.isnan(x):
        return F32(np.nan)
    with np.errstate(over="ignore"):
        f = F32(x)
    if float(f) < x:
        f = np.nextafter(f, F32(np.inf))
    return f


def _directed(exact, up, dtype):
    """round the rational `exact` to dtype towards +inf (up) or -inf"""
    big = np.finfo(dtype).max
    if exact > _fr(big):
        return dtype(np.inf) if up else dtype(big)
    if exact < -_fr(big):
        return dtype(-big) if up else dtype(-np.inf)
    r = dtype(float(exact))
    while _fr(r) > exact and not up:
        r = np.nextafter(r, dtype(-np.inf))
    while _fr(r) < exact and up:
        r = np.nextafter(r, dtype(np.inf))
    # r is on the right side; step back while the neighbour is still on the right side (nearest such value)
    while True:
        with np.errstate(over="ignore"):
            nb = np.nextafter(r, dtype(-np.inf) if up else dtype(np.inf))
        if math.isinf(float(nb)):
            break
        if (up and _fr(nb) >= exact) or (not up and _fr(nb) <= exact):
            r = nb
        else:
            break
    return r


def add_dir(a, b, up, dtype=F32):
    """a + b with directed rounding in dtype (IEEE: inf/nan propagate)"""
    a, b = dtype(a), dtype(b)
    if math.isnan(float(a)) or math.isnan(float(b)) or math.isinf(float(a)) or math.isinf(float(b)):
        with np.errstate(invalid="ignore"):
            return dtype(a + b)
    return _directed(_fr(a) + _fr(b), up, dtype)


def mul_ru(a, b):
    a, b = F32(a), F32(b)
    if math.isnan(float(a)) or math.isnan(float(b)) or math.isinf(float(a)) or math.isinf(float(b)):
        with np.errstate(invalid="ignore"):
            return F32(a * b)
    return _directed(_fr(a) * _fr(b), True, F32)


def device_tests(s, c, P, l_levels, BK=4):
    """Mirror of pruned_bounds + the row / level tests for ONE row and the levels `l_levels` of one warp.
    s[l]: stage costs, c[j][l]: jump costs, P[j]: value row.  Returns (ub[l], row_keep[q], level_keep[q][l])."""
    K = len(P)
    nblk = K // BK
    live = [x for x in P if not math.isnan(x)]
    jseed = int(np.argmin([x if not math.isnan(x) else INF for x in P])) if live else 0
    ub, ubf, e = {}, {}, {}
    for l in l_levels:
        u = INF
        for j in (jseed, l):
            v = (s[l] + c[j][l]) + P[j]
            if u > v:
                u = v
        ub[l] = u
        ubf[l] = ru32(u)
        cmx = max([ru32(abs(c[j][l])) for j in range(K) if abs(c[j][l]) < INF] + [F32(0)])
        sl_l = mul_ru(add_dir(ru32(abs(s[l])), cmx, True), F32(2.0 ** -30))
        d = add_dir(np.float64(u), np.float64(-s[l]), True, np.float64)          # __dadd_ru(ub, -s)
        ev = add_dir(ru32(float(d)), sl_l, True)
        e[l] = F32(np.inf) if math.isnan(float(ev)) else ev
    U = max(float(e[l]) for l in l_levels)
    row_keep, level_keep = [], []
    for q in range(nblk):
        blk = range(q * BK, q * BK + BK)
        vals = [P[j] for j in blk if not math.isnan(P[j])]
        pm = min(vals) if vals else float("nan")
        pq = rd32(pm)
        cminf = {l: rd32(min([c[j][l] for j in blk if not math.isnan(c[j][l])] + [INF])) for l in l_levels}
        cw = F32(min(float(cminf[l]) for l in l_levels))
        with np.errstate(invalid="ignore", over="ignore"):
            slq = F32(min(float(mul_ru(abs(pq), F32(2.0 ** -30))), float(np.finfo(F32).max))) if not math.isnan(float(pq)) else F32(np.nan)
        t = add_dir(add_dir(cw, pq, False), -slq, False)
        row_keep.append(not (float(t) > U))
        lk = {}
        for l in l_levels:
            a = add_dir(rd32(s[l]), cminf[l], False)
            lb = add_dir(a, pq, False)
            lk[l] = not (float(lb) > float(ubf[l]))
        level_keep.append(lk)
    return ub, row_keep, level_keep


def check(s, c, P, l_levels, BK=4):
    ub, row_keep, level_keep = device_tests(s, c, P, l_levels, BK)
    K = len(P)
    for q in range(K // BK):
        for l in l_levels:
            if row_keep[q] and level_keep[q][l]:
                continue                                   # the block is scanned for this level: nothing to prove
            for j in range(q * BK, q * BK + BK):
                v = (s[l] + c[j][l]) + P[j]
                # a dropped candidate must lose strictly: it can neither be the minimum nor tie with it
                assert math.isnan(v) or v > ub[l], (q, l, j, v, ub[l], row_keep[q], level_keep[q][l])


def _instance(rng, kind, K=16):
    scale = {"unit": 1.0, "tiny": 1e-42, "large": 1e25, "beyond_float": 1e200, "mixed": 1.0, "exact": 1.0, "inf": 1.0}[kind]
    s = rng.standard_normal(K) * 3 * scale
    lv = rng.integers(0, 5, size=(K, 2)).astype(float)
    c = 0.5 * scale * np.abs(lv[:, None, :] - lv[None, :, :]).sum(2)
    P = rng.standard_normal(K) * 4 * scale + rng.choice([0.0, -7.0, 11.0]) * scale
    if kind == "exact":                                    # everything a multiple of 2^-2: conversions to float are exact
        s, c, P = np.round(s * 4) / 4, np.round(c * 4) / 4, np.round(P * 4) / 4
    if kind == "mixed":                                    # magnitudes 1e-30 .. 1e30 side by side
        P = P * 10.0 ** rng.integers(-30, 30, size=K)
        s = s * 10.0 ** rng.integers(-20, 20, size=K)
    if kind == "inf":
        P[rng.integers(0, K, size=3)] = np.inf
        P[rng.integers(0, K)] = -np.inf
        P[rng.integers(0, K)] = np.nan
        c[rng.integers(0, K), rng.integers(0, K)] = np.inf
        s[rng.integers(0, K)] = np.nan
    return [float(x) for x in s], [[float(x) for x in row] for row in c], [float(x) for x in P]


@pytest.mark.parametrize("kind", ["unit", "exact", "tiny", "large", "beyond_float", "mixed", "inf"])
def test_a_dropped_block_never_holds_a_winner_or_a_tie(kind):
    rng = np.random.default_rng(["unit", "exact", "tiny", "large", "beyond_float", "mixed", "inf"].index(kind) + 20251018)
    dropped = 0
    for trial in range(40):
        s, c, P = _instance(rng, kind)
        levels = list(range(0, 8)) if trial % 2 == 0 else list(range(8, 16))   # the 'warp': eight levels share the row test
        check(s, c, P, levels)
        _, rk, lk = device_tests(s, c, P, levels)
        dropped += sum(1 for q in range(len(rk)) if not rk[q] or not any(lk[q].values()))
    if kind in ("unit", "exact", "large"):
        assert dropped > 0                                  # the tests do drop blocks on ordinary data (not vacuous)


def test_directed_rounding_helpers():
    assert float(rd32(1e300)) == float(np.finfo(F32).max) and math.isinf(float(ru32(1e300)))
    assert float(rd32(0.1)) < 0.1 < float(ru32(0.1))
    assert float(rd32(0.25)) == 0.25 == float(ru32(0.25))
    assert float(rd32(1e-50)) == 0.0 and float(ru32(1e-50)) > 0.0
    assert float(add_dir(F32(1.0), F32(2.0 ** -30), False)) == 1.0 and float(add_dir(F32(1.0), F32(2.0 ** -30), True)) > 1.0
