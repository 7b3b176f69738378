"""Host-only checks of the wavefront kernel's geometry model (bb200_wave_geometry makes no CUDA call).

The model decides how a table shape is cut into CTAs, sub-slices, thread tiles, j-groups and scatter warps; these tests
pin the choice for BASELINE config 4 on a 148-SM B200 and check the invariants every choice has to satisfy."""
import ctypes
import importlib

import numpy as np
import pytest

import mioc_b200 as m

SMEM_MAX = 232448   # opt-in dynamic shared memory per CTA on sm_100 (227 KB)
FIELDS = ("ok", "variant", "tba", "tbb", "tl", "ctas", "rows", "jsplit", "jper", "kr", "scatter_warps", "threads", "smem",
          "prune", "ctas_full", "rows_top")


def geometry(n, M, K, B, sms=148, smem=SMEM_MAX, ctas=0, jsplit=0, variant=0):
    lib = importlib.import_module(m.__name__ + "._lib").load()
    out = np.zeros(16, dtype=np.int64)
    rc = lib.bb200_wave_geometry(n, M, K, B, sms, smem, ctas, jsplit, variant,
                                 out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), 16)
    assert rc == 0
    return dict(zip(FIELDS, out.tolist()))


def test_config4_geometry_is_pinned():
    # default: the pruned (branch-and-bound) scan -- lane = level, 2 rows per warp, 4 level blocks x 4 row groups = 16 compute
    # warps + comm + publisher, blocks of 4 successors; two-zone slices on all 148 SMs: 56 CTAs x 8 rows, 92 CTAs x 6 rows
    g = geometry(100_000, 3, 125, 999)
    assert g["ok"] == 1 and g["prune"] == 4
    assert (g["tba"], g["tbb"], g["tl"]) == (2, 8, 1) and (g["ctas"], g["rows"]) == (148, 8)
    assert (g["ctas_full"], g["rows_top"]) == (56, 6) and 56 * 8 + 92 * 6 == 1000
    assert (g["jsplit"], g["kr"], g["scatter_warps"], g["threads"]) == (1, 128, 0, 576) and g["smem"] <= SMEM_MAX
    # variant -1 = what the plan falls back to when the bound test does not pay on its data: the exhaustive tiles
    g = geometry(100_000, 3, 125, 999, variant=-1)
    assert g["ok"] == 1 and g["prune"] == 0
    assert (g["tba"], g["tbb"], g["tl"]) == (4, 3, 2)          # two sub-slices of 4 + 3 rows, 2 levels per thread
    assert (g["ctas"], g["rows"]) == (143, 7)                  # 143 * 7 = 1001 >= B + 1 source rows on 148 SMs
    assert (g["jsplit"], g["jper"], g["kr"]) == (4, 32, 128)   # whole trips of the unrolled scan, +Inf pad rows 125..127
    assert g["scatter_warps"] == 6 and g["threads"] == 512
    assert g["smem"] <= SMEM_MAX


@pytest.mark.parametrize("K,B,n", [(125, 999, 3000), (36, 1638, 8192), (36, 204, 1024), (64, 211, 70), (16, 120, 50),
                                   (81, 150, 45), (33, 97, 45), (7, 40, 45), (100, 333, 45), (150, 5000, 200), (3, 2, 3)])
@pytest.mark.parametrize("sms", [148, 132, 20])
def test_geometry_invariants(K, B, n, sms):
    g = geometry(n, 2, K, B, sms=sms)
    if not g["ok"]:
        pytest.skip("shape runs on the per-stage kernels")
    Kp = (K + 31) // 32 * 32
    assert 1 <= g["ctas"] <= sms
    ga, rt = g["ctas_full"], g["rows_top"]
    assert 0 <= ga <= g["ctas"] and 1 <= rt <= g["rows"]
    assert ga * g["rows"] + (g["ctas"] - ga) * rt >= B + 1     # every source row has an owner
    last_r0 = ga * g["rows"] + (g["ctas"] - 1 - ga) * rt if ga < g["ctas"] else (g["ctas"] - 1) * g["rows"]
    assert last_r0 < B + 1                                     # and no CTA is empty
    if g["prune"]:
        assert g["rows"] == g["tbb"] and g["jsplit"] == 1 and g["scatter_warps"] == 0 and g["kr"] % 32 == 0
    else:
        assert g["rows"] % (g["tba"] + g["tbb"]) == 0
    assert g["threads"] % 32 == 0 and 96 <= g["threads"] <= 640
    assert g["smem"] <= SMEM_MAX
    assert g["jsplit"] * g["jper"] >= K and g["jper"] % 2 == 0 # the j-groups cover every successor, in aligned pairs
    assert g["kr"] in (K, g["jsplit"] * g["jper"]) and g["kr"] <= max(K, Kp)
    assert g["prune"] or (g["tbb"] == 0) or g["scatter_warps"] >= 1   # two sub-slices one after the other need scatter warps


def test_shapes_the_wavefront_kernel_refuses():
    assert geometry(1000, 1, 300, 100)["ok"] == 0              # K > 255: uint16 argmin, per-stage kernels
    assert geometry(1000, 2, 200, 100)["ok"] == 0              # jump-cost table (200 x 224 x 8 B) does not fit
    assert geometry(1000, 2, 125, 999, smem=48 * 1024)["ok"] == 0


def test_forced_tuning_is_respected():
    g = geometry(3000, 3, 125, 999, jsplit=2, variant=213)     # tile 13 = (4+3) x 2 with two scatter warps
    assert g["ok"] == 1 and (g["variant"], g["jsplit"], g["scatter_warps"]) == (13, 2, 2)
    g = geometry(3000, 3, 125, 999, variant=1001)              # tile 1 = (7+0) x 2, compute warps finish their own stage
    assert g["ok"] == 1 and (g["variant"], g["tbb"], g["scatter_warps"]) == (1, 0, 0)
    assert geometry(3000, 3, 125, 999, ctas=100)["ctas"] <= 100
