"""The oracle pinned against the REFERENCE ITSELF, run here.

oracle/Makefile (target `ref`) compiles the reference's own source/transfer2.c and source/bessel2.c, where they lie under
/root/reference, into oracle/_ref/libsong_ref.so together with the harness that plays main/song.c:100-106.  The CLASS
fork the reference depends on is absent (an empty submodule), so its declarations and the handful of helpers the two
files call are stand-ins (oracle/ref_build/ref_prelude.h); everything between transfer2_init and transfer2_integrate,
and between bessel2_init and bessel2_J_Llm, is the reference's own code, unmodified.

What is compared, bit for bit (np.array_equal):
  * ptr2->transfer after the reference's transfer2_init  ==  the oracle's transfer functions, for the same sources and the
    same projection tables, over the options of the path (k3 interpolation, extrapolation, fields, m, time sampling and
    time interpolation, Bessel interpolation, long k3 grids) and on BASELINE configs[0] at full size;
  * struct bessels2 after the reference's bessel2_init  ==  the oracle's tables, when both start from the same j_l(x) and
    3j numbers (libsong_ref_op.so: the two providers of the absent fork answered by the oracle's);
  * bessel2_convolution; the product's host half of bessel2 against the reference's pbs2->j_l1; the disk format;
  * bispectra2_intrinsic_integrate_over_k3 and spectra2_cls of the reference's own bispectra2.c / spectra2.c
    (libsong_ref_bisp.so, further stand-ins under oracle/ref_build/fork_stubs/)  ==  oracle_bispectra2.c / oracle_spectra2.c.
The GPU path is bit-identical to the oracle in EXACT mode (tests -m gpu), hence to the reference's code.

The library does not exist on the GPU box unless it travelled with the snapshot; tests skip when it is missing.
"""
import ctypes
import os

import numpy as np
import pytest

import dropin_lib
from song_b200 import configs, workload

REF = dropin_lib.REF_PATH
REF_OP = REF.replace("libsong_ref.so", "libsong_ref_op.so")
needs_ref = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/libsong_ref.so not built (needs /root/reference)")
needs_ref_op = pytest.mark.skipif(not os.path.exists(REF_OP), reason="oracle/_ref/libsong_ref_op.so not built")


def reference_transfer2(wl, proj, lib_path=REF):
    with dropin_lib.Harness(wl, proj, lib_path=lib_path) as h:
        h.transfer2_init()
        got, count, labels_ok = h.read()
        h.transfer2_free()
    return got, count, labels_ok


VARIANTS = [
    dict(sources_k3_interpolation="linear"),
    dict(sources_k3_interpolation="cubic"),
    dict(output=["tBisp"], sources_k3_interpolation="cubic"),
    dict(bispectra_k3_extrapolation="no_extrapolation", sources_k3_interpolation="cubic"),
    dict(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True),
    dict(bessels_interpolation="cubic", sources_k3_interpolation="cubic"),
    dict(k3_size_min=300, sources_k3_interpolation="cubic"),
    dict(l_max_los_t=14, l_max_los_p=14, sources_k3_interpolation="cubic"),
    dict(extra_k3_oscillations_left=0.5, extra_k3_oscillations_right=12, sources_k3_interpolation="cubic"),
]


@needs_ref
@pytest.mark.parametrize("kind", ["smooth", "noise"])
@pytest.mark.parametrize("kw", VARIANTS, ids=lambda kw: ",".join(f"{k}={v}" for k, v in kw.items())[:60])
def test_reference_transfer2_init_equals_oracle(oracle, kw, kind):
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)), kind=kind)
    proj = oracle.build_projection(wl, spline=kw.get("bessels_interpolation") == "cubic")
    ref, _ = oracle.OracleRun(wl, proj).run()
    got, count, labels_ok = reference_transfer2(wl, proj)
    assert got.shape == ref.shape                      # the k3 grids of transfer2_get_k3_sizes have the oracle's / los_host.c's sizes
    assert count == wl.n_integrals and labels_ok       # count_memorised_transfers (transfer2.c:607-610), labels
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"
    assert np.abs(ref).max() > 0


@needs_ref
@pytest.mark.parametrize("time_interp", ["linear", "cubic"])
@pytest.mark.parametrize("sampling", ["custom", "bessel"])
def test_reference_k_dependent_time_grids_equal_oracle(oracle, sampling, time_interp):
    """transfer2_get_time_grid (transfer2.c:2366-2585) and transfer2_interpolate_sources_in_time (:2783-2908)."""
    p = configs.tiny(transfer2_tau_sampling=sampling, sources_time_interpolation=time_interp,
                     sources_k3_interpolation="cubic", tau_linstep_song=0.003, bessel_x_step_song=0.04)
    wl = workload.fill_sources(workload.build_grids(p), kind="noise")
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    got, count, _ = reference_transfer2(wl, proj)
    assert count == wl.n_integrals
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"


@needs_ref
def test_reference_equals_oracle_on_baseline_config0(oracle):
    """BASELINE.json configs[0] (quick_song_run) at full size through the reference's transfer2_init."""
    wl = workload.fill_sources(workload.build_grids(configs.get("quick_song_run")))
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    got, count, labels_ok = reference_transfer2(wl, proj)
    assert count == wl.n_integrals == 654112 and labels_ok
    assert np.array_equal(got, ref)


@needs_ref_op
@pytest.mark.parametrize("kw", [dict(sources_k3_interpolation="cubic"), dict(output=["tBisp"]),
                                dict(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True),
                                dict(bessels_interpolation="cubic", sources_k3_interpolation="cubic")],
                         ids=["T+E", "T", "m012+B", "cubicJ"])
def test_reference_bessel2_init_equals_oracle(oracle, kw):
    """struct bessels2 as the reference's bessel2_init leaves it (bessel2.c:75-468) against the oracle's restatement of
    bessel2_get_l1_list, bessel2_j_for_l1, bessel2_J_for_Llm and bessel2_J_Llm, from the same j_l(x) and 3j numbers; then
    transfer2_init on those tables."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)), kind="noise")
    spline = kw.get("bessels_interpolation") == "cubic"
    proj = oracle.build_projection(wl, with_jl1=True, spline=spline)
    dropin_lib.lib(REF_OP).ref_set_bessel_table_lmax(ctypes.c_int(int(proj.jl1["l1"].max())))
    with dropin_lib.Harness(wl, None, lib_path=REF_OP) as h:      # proj None: the harness calls bessel2_init
        rows, xmin, xsize, dd, info = h.bessel2_tables()
        l1, x_min_l1 = h.l1_info()
        h.transfer2_init()
        got, count, _ = h.read()
        h.transfer2_free()
    assert np.array_equal(l1, proj.jl1["l1"]) and np.array_equal(x_min_l1, proj.jl1["x_min_l1"])      # l1 list and lower cuts
    assert np.array_equal(xsize, proj.x_size)
    live = proj.x_size > 0
    assert np.array_equal(xmin[live], proj.index_xmin[live])
    n = 0
    for a, b in zip(rows, proj.rows):
        assert (a is None) == (b is None)
        if a is not None:
            assert np.array_equal(a, b)
            n += 1
    assert n == int(live.sum()) and n > 0
    if spline:                                          # pbs2->ddJ_Llm_x, bessel2.c:395-435
        # (tables shorter than 3 points: upstream reads past the end there; neither side uses them)
        assert all(np.array_equal(a, b) for a, b, r in zip(dd, proj.ddrows, proj.rows) if a is not None and r.size >= 3)
    ref, _ = oracle.OracleRun(wl, proj).run()
    assert count == wl.n_integrals and np.array_equal(got, ref)


@needs_ref_op
def test_reference_bessel2_convolution_equals_oracle(oracle):
    """bessel2_convolution (bessel2.c:499-605) of the reference on its own j_l1 table against the oracle's restatement on
    the oracle's table -- the same numbers, see test_reference_bessel2_init_equals_oracle."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")))
    proj = oracle.build_projection(wl, with_jl1=True)
    dropin_lib.lib(REF_OP).ref_set_bessel_table_lmax(ctypes.c_int(int(proj.jl1["l1"].max())))
    rng = np.random.default_rng(5)
    kk = np.sort(rng.uniform(1e-4, 0.05, 60)); dk = np.gradient(kk); f = rng.standard_normal(60); g = rng.standard_normal(60)
    f[7] = 0.0; g[11] = 0.0
    with dropin_lib.Harness(wl, None, lib_path=REF_OP) as h:
        l1, x_min_l1 = h.l1_info()
        vals = [(i, r, h.bessel2_convolution(kk, dk, f, g if i % 2 else None, i, r)) for i in (0, 3, len(l1) - 1) for r in (200.0, 1000.0)]   # k r stays inside the table (xx_max = 63)
    assert any(v != 0.0 for _, _, v in vals)
    assert np.array_equal(proj.jl1["l1"], l1) and np.array_equal(proj.jl1["x_min_l1"], x_min_l1)
    L = oracle.lib()
    for i, r, v in vals:
        out = ctypes.c_double(0)
        gp = g.ctypes.data_as(oracle.dp) if i % 2 else None
        row = np.ascontiguousarray(proj.jl1["j_l1"][i])
        assert L.oracle_bessel2_convolution(kk.ctypes.data_as(oracle.dp), dk.ctypes.data_as(oracle.dp), 60, f.ctypes.data_as(oracle.dp), gp,
                                            row.ctypes.data_as(oracle.dp), None, float(x_min_l1[i]), float(wl.xx_step), 0, float(r),
                                            ctypes.byref(out)) == 0
        assert out.value == v, (i, r, out.value, v)


@needs_ref
def test_reference_and_dropin_libraries_coexist():
    """Both libraries export transfer2_init & co.; the reference one is linked -Bsymbolic and loaded RTLD_LOCAL, so that
    neither binds to the other's symbols when a test process holds both."""
    if not os.path.exists(dropin_lib.PATH):
        pytest.skip("libsong_transfer2.so not built")
    a, b = dropin_lib.lib(), dropin_lib.lib(REF)
    assert ctypes.cast(a.transfer2_init, ctypes.c_void_p).value != ctypes.cast(b.transfer2_init, ctypes.c_void_p).value
    wl = workload.fill_sources(workload.build_grids(configs.tiny()))
    import oracle_lib as O
    proj = O.build_projection(wl)
    ref, _ = O.OracleRun(wl, proj).run()
    got, _, _ = reference_transfer2(wl, proj)          # with the drop-in loaded: still the reference's own code path
    assert np.array_equal(got, ref)


# ---- the GPU against the reference's own code, directly (driver-run: -m gpu) ----
@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("kw,kind", [(dict(sources_k3_interpolation="cubic"), "noise"),
                                     (dict(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True), "smooth")],
                         ids=["T+E", "m012+B"])
def test_gpu_equals_reference_transfer2_init(oracle, kw, kind, monkeypatch):
    """The engine through the C-ABI (EXACT: bit for bit; FAST: <= 1e-9) and the drop-in's transfer2_init (EXACT) against
    ptr2->transfer as the reference's own transfer2_init leaves it, same sources, same projection tables."""
    from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)), kind=kind)
    proj = oracle.build_projection(wl)
    ref, count, _ = reference_transfer2(wl, proj)
    assert count == wl.n_integrals
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(proj.rows, proj.index_xmin, proj.x_size, proj.ddrows))
    got = eng.transfer_pairs(wl, mode=MODE_EXACT)
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"
    fast = eng.transfer_pairs(wl, mode=MODE_FAST)
    eng.close()
    scale = np.abs(ref).max(axis=1, keepdims=True)
    sel = np.abs(ref) > 1e-6 * scale
    assert float(np.max(np.abs(fast - ref)[sel] / np.abs(ref)[sel])) < 1e-9
    if os.path.exists(dropin_lib.PATH):
        monkeypatch.setenv("SONG_LOS_MODE", "exact")
        got2, count2, labels_ok = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
        assert count2 == count and labels_ok and np.array_equal(got2, ref)


@needs_ref
def test_product_host_half_equals_reference_j_l1_table(oracle, monkeypatch):
    """The host half of the bessel2 drop-in (bessel2_host.c) against pbs2->j_l1 as the reference's bessel2_init leaves it
    (bessel2_j_for_l1, bessel2.c:1669-1815), same j_l provider on both sides: l1 list, lower cuts, first grid indices and,
    with the reference's point-by-point call pattern (the one a SONG tree gets, where CLASS's bessel_j is the provider),
    every value of the table, bit for bit.  The default pattern of the stand-in provider (one recurrence per x for all
    l1) gives the same numbers to the last 1-3 bits."""
    from song_b200.bessel2 import Bessel2
    wl = workload.fill_sources(workload.build_grids(configs.tiny(modes_song=[0, 1, 2], has_source_B=True)))
    with dropin_lib.Harness(wl, None, lib_path=REF) as h:          # the reference's bessel2_init, compat providers
        l1, x_min_l1 = h.l1_info()
        rows, xmin = h.j_l1_rows(wl.xx.size)
    monkeypatch.setenv("SONG_LOS_JL1_PER_POINT", "1")
    ht = Bessel2(wl).prepare().host_tables()
    assert np.array_equal(ht["l1"], l1) and np.array_equal(ht["x_min_l1"], x_min_l1) and np.array_equal(ht["index_xmin_l1"], xmin)
    assert all(np.array_equal(a, b) for a, b in zip(ht["j_l1"], rows)) and len(rows) == len(ht["j_l1"]) > 10
    monkeypatch.setenv("SONG_LOS_JL1_PER_POINT", "0")
    ht0 = Bessel2(wl).prepare().host_tables()
    worst = max(float(np.max(np.abs(a - b))) for a, b in zip(ht0["j_l1"], rows))
    assert np.array_equal(ht0["index_xmin_l1"], xmin) and worst < 1e-15


@needs_ref
def test_reference_disk_format_is_what_the_dropin_writes(oracle, tmp_path):
    """store_transfers_to_disk through the REFERENCE (transfer2.c:559-567, 2929-2963): one transfers_%03d.dat per type, the
    values in [k1][k2][k3] order -- the layout the drop-in's writer thread reproduces with one pwrite per batch and type
    (tests/test_dropin.py checks its files against the same expectation on the GPU) -- and read back through the
    reference's transfer2_load_transfers_from_disk / transfer2_free_type_level (the path bispectra2.c:1213,1298 takes)."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny()))
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    with dropin_lib.Harness(wl, proj, store_dir=str(tmp_path), lib_path=REF) as h:
        h.transfer2_init()
        got, count, _ = h.read()                       # with store_to_disk the harness reads every type back from its file
        h.transfer2_free()
    assert np.array_equal(got, ref)
    for tt in (0, wl.layout.tt2_size // 2, wl.layout.tt2_size - 1):
        f = tmp_path / f"transfers_{tt:03d}.dat"
        assert f.exists() and f.stat().st_size == 8 * wl.n_units
        assert np.array_equal(np.fromfile(f), ref[tt])


REF_BISP = REF.replace("libsong_ref.so", "libsong_ref_bisp.so")


@pytest.mark.skipif(not os.path.exists(REF_BISP), reason="oracle/_ref/libsong_ref_bisp.so not built")
@pytest.mark.parametrize("abs_M3,offset_L3", [(0, 0), (1, 0), (1, 2)])
def test_reference_integrate_over_k3_equals_oracle(oracle, abs_M3, offset_L3):
    """bispectra2_intrinsic_integrate_over_k3 (bispectra2.c:1107-1439) of the reference -- its own bispectra2.c, compiled
    where it lies -- on ptr2->transfer and pbs2->j_l1 as the reference's transfer2_init and bessel2_init left them, against
    the oracle's restatement (oracle_bispectra2.c; the loop song_los_integrate_k3 reproduces on the GPU) on the same
    numbers: the trapezoidal k3 measure of :1279-1284, bessel2_convolution, the M3 coefficient and the factor 2."""
    kw = dict(sources_k3_interpolation="cubic") if abs_M3 == 0 else dict(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True)
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)))
    lay = wl.layout
    index_M3 = int(np.nonzero(lay.m == abs_M3)[0][0])
    r = np.linspace(0.55, 0.9, 7) * wl.xx_max / float(wl.k3_tr.max())        # k3 r stays inside the j_l1 table
    coef = 0.7
    with dropin_lib.Harness(wl, None, lib_path=REF_BISP) as h:
        h.transfer2_init()
        T, count, _ = h.read()
        l1, x_min_l1 = h.l1_info()
        rows, _ = h.j_l1_rows(wl.xx.size)
        got = h.integrate_over_k3(lay.index_tt2_T, index_M3, offset_L3, abs_M3, coef, r)
        h.transfer2_free()
    assert count == wl.n_integrals
    checked = 0
    for il in range(lay.l_size):
        l3 = int(lay.l[il])
        L3 = abs(l3 - abs_M3) + offset_L3
        if abs_M3 > l3 or L3 > l3 + abs_M3:
            assert not got[il].any()                   # skipped configurations stay zero (:1181-1190)
            continue
        index_L3 = int(np.nonzero(l1 == L3)[0][0])
        tt = lay.index_tt2_T + int(lay.lm_array[il, index_M3])
        ref = oracle.integral_over_k3(wl, T[tt], rows[index_L3], None, float(x_min_l1[index_L3]), r, abs_M3=abs_M3, M3_coefficient=coef)
        assert np.array_equal(got[il], ref), (il, float(np.abs(got[il] - ref).max()))
        assert np.abs(ref).max() > 0
        checked += 1
    assert checked >= 2


@pytest.mark.skipif(not os.path.exists(REF_BISP), reason="oracle/_ref/libsong_ref_bisp.so not built")
@pytest.mark.parametrize("rescale", [True, False])
def test_reference_spectra2_cls_equals_oracle(oracle, rescale):
    """spectra2_cls (spectra2.c:241-570) of the reference -- its own spectra2.c, compiled where it lies -- on ptr2->transfer
    as the reference's transfer2_init left it, against the oracle's restatement (oracle_spectra2.c; the sum song_los_cl_sum
    reproduces on the GPU): tt2, ee2 and te2 for every l, m = 0, with the k1/k2/k3 trapezoidal steps, the triangular
    clipping of the k3 step and the primordial spectra of :371-515."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")))
    lay = wl.layout
    prim = 2.1e-9 * (wl.k / 0.05) ** (0.96 - 1.0)
    proj = oracle.build_projection(wl)
    with dropin_lib.Harness(wl, proj, lib_path=REF_BISP) as h:
        h.transfer2_init()
        T, count, _ = h.read()
        got = h.spectra2_cls(prim, rescale_cmb_sources=rescale)
        h.transfer2_free()
    assert count == wl.n_integrals and got.shape == (3, lay.l_size)
    for il in range(lay.l_size):
        tT = lay.index_tt2_T + int(lay.lm_array[il, 0]); tE = lay.index_tt2_E + int(lay.lm_array[il, 0])
        for ct, (a, b) in enumerate(((tT, tT), (tE, tE), (tT, tE))):
            ref = oracle.spectra2_cl(wl, T[a], T[b], 0, prim, rescale_cmb_sources=rescale)
            assert got[ct, il] == ref, (ct, il, got[ct, il], ref)
    assert np.abs(got).max() > 0


@needs_ref
@pytest.mark.parametrize("name,n_k", [("sn_t_10percent_lmax2000", 6), ("sn_pol_10percent_lmax2000", 6),
                                      ("sn_pol_5percent_lmax3000", 3), ("sn_pol_5percent_lmax4000", 2)])
def test_reference_equals_oracle_on_baseline_triangles(oracle, name, n_k):
    """BASELINE.json configs[1]-[4] at their own parameters (l lists up to 4000, x grids of 17-53 thousand points, L up to
    17, 60-240 time samples): the reference's transfer2_init on the whole (k1,k2) triangle of n_k of the run's k values
    (the sample bench.py's CPU arm times) against the oracle, bit for bit."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    wl, k_size = bench.reference_triangle(name, n_k)
    assert wl.n_pairs == n_k * (n_k + 1) // 2 and wl.k.size == n_k < k_size
    workload.fill_sources(wl)
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    got, count, labels_ok = reference_transfer2(wl, proj)
    assert count == wl.n_integrals and labels_ok
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"


@needs_ref
@pytest.mark.parametrize("kw", [dict(), dict(bispectra_k3_extrapolation="no_extrapolation"),
                                dict(extra_k3_oscillations_left=0.5, extra_k3_oscillations_right=12), dict(q_linstep_song=0.13)],
                         ids=["default", "no extrapolation", "asymmetric wings", "finer q step"])
def test_reference_k3_grids_equal_the_products(oracle, kw):
    """transfer2_get_k3_list of the reference (transfer2.c:1566-1737), pair by pair, against the fine k3 grids the product's
    host code builds (song_b200/host/los_host.c, what the GPU integrates on) and against the oracle's: bit for bit, and the
    physical ranges of transfer2_get_k3_size (:1329-1550) through the wing handling of the transfer functions."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)))
    proj = oracle.build_projection(wl)
    with dropin_lib.Harness(wl, proj, lib_path=REF) as h:
        h.transfer2_init()
        grids = h.k3_lists()
        h.transfer2_free()
    assert np.array_equal(grids, wl.k3_tr)
    _, ogrids = oracle.OracleRun(wl, proj).k3_grids()
    assert np.array_equal(grids, ogrids)


@needs_ref
@pytest.mark.parametrize("kw", [dict(), dict(output=["tBisp"]), dict(modes_song=[0, 1, 2], has_source_B=True),
                                dict(bispectra_k3_extrapolation="no_extrapolation")], ids=["T+E", "T", "m012+B", "no extrapolation"])
def test_dropin_host_module_against_reference_without_a_gpu(oracle, kw, tmp_path):
    """The host side of the drop-in (song_b200/host/transfer2_gpu.c) against the reference's transfer2.c where no GPU is
    involved: with load_transfers_from_disk the module runs transfer2_indices_of_transfers and transfer2_get_k3_sizes and
    returns (transfer2.c:151-157).  (1) Both modules leave the same index tables, labels, counters, grid sizes and physical
    ranges in struct transfers2; (2) the drop-in's transfer2_load_transfers_from_disk / transfer2_free_type_level read the
    transfers_%03d.dat files THE REFERENCE wrote and return its values bit for bit (the path bispectra2.c:1213 takes)."""
    if not os.path.exists(dropin_lib.PATH):
        pytest.skip("libsong_transfer2.so not built")
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)))
    proj = oracle.build_projection(wl)
    with dropin_lib.Harness(wl, proj, store_dir=str(tmp_path), lib_path=REF) as h:       # the reference computes and stores
        h.transfer2_init()
        ref_vals, count, _ = h.read()
        ref_digest = h.indices_digest()
        h.transfer2_free()
    assert count == wl.n_integrals
    with dropin_lib.Harness(wl, proj, load_dir=str(tmp_path)) as h:                       # the drop-in: indices + load, no GPU
        h.transfer2_init()
        digest = h.indices_digest()
        got, _, labels_ok = h.read()
        h.transfer2_free()
    assert digest == ref_digest
    assert labels_ok and np.array_equal(got, ref_vals)
