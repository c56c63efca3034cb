"""Projection functions J_Llm(x) assembled on the GPU (song_los_build_projection, the device half
of bessel2_init) against the oracle's restatement of bessel2.c:1041-1588."""
import ctypes as C

import numpy as np
import pytest

from song_b200 import configs, workload
from song_b200.bessel2 import Bessel2, build_projection_direct
from song_b200.engine import MODE_EXACT, LosEngine

pytestmark = pytest.mark.gpu


def oracle_3j_blocks(oracle, wl, J_types, jl1):
    """first/second 3j of every (J,L,l,m) exactly as the oracle's bessel2_J_for_Llm computes them."""
    lay = wl.layout
    L = oracle.lib()
    ndim = 2 * (lay.L_size - 1) + 1
    n = lay.J_size * lay.L_size * lay.l_size * lay.m_size
    f1 = np.zeros((n, ndim)); f2 = np.zeros((n, ndim))
    l1_min = np.zeros(n, np.int32); cnt = np.zeros(n, np.int32); idx = np.zeros(n, np.int32)
    for J in range(lay.J_size):
        pf = int(J_types[J]); S = 0 if pf == 0 else 2
        for Lv in range(lay.L_size):
            for il in range(lay.l_size):
                for im in range(lay.m_size):
                    lv, mv = int(lay.l[il]), int(lay.m[im])
                    q = ((J * lay.L_size + Lv) * lay.l_size + il) * lay.m_size + im
                    if mv > min(Lv, lv) or S > lv or S > Lv or (pf == 2 and mv == 0):
                        continue
                    lo, hi = C.c_double(), C.c_double()
                    a = np.zeros(ndim); b = np.zeros(ndim)
                    assert L.oracle_drc3jj(Lv, lv, S, -S, C.byref(lo), C.byref(hi), a.ctypes.data_as(oracle.dp), ndim) == 0
                    l1min = int(lo.value + 1e-6); c = int(hi.value + 1e-6) - l1min + 1
                    if S != 0 or mv != 0:
                        assert L.oracle_drc3jj(Lv, lv, -mv, mv, C.byref(lo), C.byref(hi), b.ctypes.data_as(oracle.dp), ndim) == 0
                    else:
                        b[:c] = a[:c]
                    f1[q], f2[q] = a, b
                    l1_min[q], cnt[q], idx[q] = l1min, c, jl1["index_l1"][l1min]
    return f1, f2, l1_min, cnt, idx


@pytest.mark.parametrize("params", [configs.tiny(), configs.tiny(modes_song=[0, 1, 2], has_source_B=True),
                                    configs.get("quick_song_run")], ids=["tiny", "tiny_m012", "quick"])
def test_device_tables_bitwise_from_the_same_inputs(oracle, params):
    wl = workload.build_grids(params)
    oproj = oracle.build_projection(wl, with_jl1=True)
    b2 = Bessel2(wl)
    f1, f2, l1_min, cnt, idx = oracle_3j_blocks(oracle, wl, b2.J_types, oproj.jl1)
    eng = LosEngine(wl, 0)
    build_projection_direct(eng, wl, oproj.jl1["l1"], oproj.jl1["j_l1"], oproj.jl1["index_xmin_l1"], f1, f2,
                            l1_min, cnt, idx, b2.J_types)
    rows, xmin, xsize = b2.tables(eng)
    lay = wl.layout
    checked = 0
    for q, ref in enumerate(oproj.rows):
        if ref is None:
            assert xsize[q] == 0
            continue
        assert xmin[q] == oproj.index_xmin[q], q
        assert xsize[q] == oproj.x_size[q], q
        assert np.array_equal(rows[q], ref), (q, np.abs(rows[q] - ref).max())
        checked += 1
    assert checked > 0
    eng.close()


def test_device_second_derivatives_bitwise_and_cubic_los(oracle):
    """bessels_interpolation = cubic: ddJ_Llm_x built on the device (bessel2.c:395-435: EST_DERIV for l = 2, NATURAL
    otherwise) is bit-identical to the oracle's, and the LOS integral with the device-built {J, ddJ} tables
    (transfer2.c:2262-2271) is bit-identical to the oracle run on the oracle's tables."""
    from song_b200.engine import MODE_EXACT
    wl = workload.build_grids(configs.tiny(bessels_interpolation="cubic", sources_k3_interpolation="cubic"))
    workload.fill_sources(wl)
    oproj = oracle.build_projection(wl, with_jl1=True, spline=True)
    b2 = Bessel2(wl)
    f1, f2, l1_min, cnt, idx = oracle_3j_blocks(oracle, wl, b2.J_types, oproj.jl1)
    eng = LosEngine(wl, 0)
    build_projection_direct(eng, wl, oproj.jl1["l1"], oproj.jl1["j_l1"], oproj.jl1["index_xmin_l1"], f1, f2,
                            l1_min, cnt, idx, b2.J_types)
    rows, xmin, xsize = b2.tables(eng)
    dd = b2.dd_tables(eng, xsize)
    checked = 0
    for q, ref in enumerate(oproj.ddrows):
        if ref is None:
            continue
        assert np.array_equal(dd[q], ref), (q, np.abs(dd[q] - ref).max())
        checked += int(np.any(ref != 0))
    assert checked > 0
    ref_T, _ = oracle.OracleRun(wl, oproj).run()
    got = eng.transfer_pairs(wl, mode=MODE_EXACT)
    assert np.array_equal(got, ref_T)
    eng.close()


def test_product_bessel2_matches_oracle_tables(oracle):
    """Bessel2.init (host compat j_l / 3j + device assembly) against the oracle's own generators:
    different algorithms for the inputs, so agreement is to rounding, not bitwise."""
    wl = workload.build_grids(configs.tiny())
    oproj = oracle.build_projection(wl)
    eng = LosEngine(wl, 0)
    b2 = Bessel2(wl).init(eng)
    rows, xmin, xsize = b2.tables(eng)
    for q, ref in enumerate(oproj.rows):
        if ref is None:
            continue
        assert abs(int(xmin[q]) - int(oproj.index_xmin[q])) <= 1
        if xmin[q] == oproj.index_xmin[q]:
            assert np.allclose(rows[q], ref, rtol=0, atol=1e-11 * max(np.abs(ref).max(), 1e-300))
    eng.close()


def test_transfer_with_device_built_tables(oracle):
    """End to end through the product path only: Bessel2.init -> EXACT transfer; the oracle is fed
    the tables read back from the device and must agree bit for bit."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")))
    eng = LosEngine(wl, 0)
    b2 = Bessel2(wl).init(eng)
    rows, xmin, xsize = b2.tables(eng)
    got = eng.transfer_pairs(wl, mode=MODE_EXACT)
    proj = oracle.OracleProjection(rows, xmin, xsize, None)
    ref, _ = oracle.OracleRun(wl, proj).run()
    assert np.array_equal(got, ref)
    eng.close()
