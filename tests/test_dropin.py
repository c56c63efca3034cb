"""The SONG-facing drop-in (song_b200/host/transfer2_gpu.c): same symbols as source/transfer2.c, same results."""
import ctypes as C
import os

import numpy as np
import pytest

import dropin_lib
from song_b200 import _lib, configs, workload
from song_b200.engine import Projection


def test_dropin_exports_the_reference_module_api():
    dll = C.CDLL(dropin_lib.PATH) if os.path.exists(dropin_lib.PATH) else None
    if dll is None:
        pytest.skip("libsong_transfer2.so not built (needs the SONG headers)")
    _lib.host()
    for name in dropin_lib.DROPIN_SYMBOLS:
        assert hasattr(dll, name), f"transfer2.h symbol {name} is not exported"


def _tiny(**kw):
    wl = workload.build_grids(configs.tiny(**kw))
    workload.fill_sources(wl)
    return wl


def test_dropin_fails_loudly_without_a_gpu(oracle):
    """Error convention of the module API: _FAILURE_ + text in ptr2->error_message, no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    if not os.path.exists(dropin_lib.PATH):
        pytest.skip("libsong_transfer2.so not built")
    wl = _tiny()
    proj = oracle.build_projection(wl)
    with pytest.raises(_lib.SongLosError) as exc:
        dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
    assert "no CUDA device" in str(exc.value) or "CUDA" in str(exc.value)


def _rel_err(got, ref):
    scale = np.abs(ref).max(axis=1, keepdims=True)
    sel = np.abs(ref) > 1e-6 * scale
    return float(np.max(np.abs(got - ref)[sel] / np.abs(ref)[sel]))


@pytest.mark.gpu
@pytest.mark.parametrize("k3_interp", ["linear", "cubic"])
def test_transfer2_init_matches_oracle(oracle, k3_interp, monkeypatch):
    wl = _tiny(sources_k3_interpolation=k3_interp)
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    p = Projection(proj.rows, proj.index_xmin, proj.x_size)
    monkeypatch.setenv("SONG_LOS_MODE", "exact")
    got, count, labels_ok = dropin_lib.run_transfer2(wl, p)
    assert np.array_equal(got, ref), "EXACT mode through transfer2_init must be bit-identical to the oracle"
    assert count == wl.n_integrals and labels_ok           # count_memorised_transfers, transfer2.c:607-610
    monkeypatch.setenv("SONG_LOS_MODE", "fast")
    got, count, _ = dropin_lib.run_transfer2(wl, p)
    assert _rel_err(got, ref) < 1e-9 and count == wl.n_integrals


@pytest.mark.gpu
def test_transfer2_init_small_batches_and_polarisation(oracle, monkeypatch):
    """Several k1 batches (SONG_LOS_BATCH_UNITS) and non-scalar modes with B-mode mixing."""
    wl = _tiny(modes_song=[0, 1, 2], has_source_B=True, sources_k3_interpolation="cubic")
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    monkeypatch.setenv("SONG_LOS_BATCH_UNITS", "300")
    monkeypatch.setenv("SONG_LOS_MODE", "fast")
    got, count, labels_ok = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
    assert _rel_err(got, ref) < 1e-9 and count == wl.n_integrals and labels_ok


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_transfer2_init_multi_gpu_schedule_and_pipeline(oracle, monkeypatch, mode):
    """Many small k1 batches spread over up to 8 GPUs (one host thread and one three-deep submit/wait pipeline per GPU,
    static LPT schedule).  On a single-GPU box the module finds one device and runs the same pipeline there; the result
    must not depend on the number of devices."""
    wl = _tiny(sources_k3_interpolation="cubic")
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    monkeypatch.setenv("SONG_LOS_BATCH_UNITS", "150")
    monkeypatch.setenv("SONG_LOS_DEVICES", "8")
    monkeypatch.setenv("SONG_LOS_MODE", mode)
    got, count, labels_ok = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
    assert count == wl.n_integrals and labels_ok
    if mode == "exact":
        assert np.array_equal(got, ref)
    else:
        assert _rel_err(got, ref) < 1e-9


@pytest.mark.gpu
def test_transfer2_custom_time_sampling(oracle, monkeypatch):
    """ptr2->tau_sampling = custom_tau_sampling through the drop-in (k-dependent time grids, cubic time interpolation)."""
    wl = _tiny(transfer2_tau_sampling="custom", sources_time_interpolation="cubic", sources_k3_interpolation="cubic",
               tau_linstep_song=0.003)
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    got, count, _ = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
    assert np.array_equal(got, ref) and count == wl.n_integrals


@pytest.mark.gpu
def test_transfer2_store_and_load_from_disk(oracle, tmp_path, monkeypatch):
    """store_transfers_to_disk: one transfers_%03d.dat per type in [k1][k2][k3] order, read back with
    transfer2_load_transfers_from_disk (the path bispectra2.c:1213 takes)."""
    wl = _tiny()
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    monkeypatch.setenv("SONG_LOS_MODE", "exact")
    got, _, _ = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size), store_dir=str(tmp_path))
    assert np.array_equal(got, ref)
    f = tmp_path / "transfers_000.dat"
    assert f.exists() and f.stat().st_size == 8 * wl.n_units
    assert np.array_equal(np.fromfile(f), ref[0])


# ---------------------------------------------------------------------------------------------------------------
# bessel2 drop-in (song_b200/host/bessel2_gpu.c) and the hand-over of its engines to transfer2_init
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(sources_k3_interpolation="cubic"),
                                dict(modes_song=[0, 1, 2], has_source_B=True),
                                dict(bessels_interpolation="cubic", sources_k3_interpolation="cubic")],
                         ids=["scalar", "m012_B", "cubic_bessel"])
def test_bessel2_init_then_transfer2_init(oracle, kw, monkeypatch):
    """main/song.c:100-106: bessel2_init fills struct bessels2 (device-built J_Llm_x, host j_l1) and leaves the engines
    on the GPUs; transfer2_init takes them over.  The oracle, fed the tables read from pbs2, must agree bit for bit in
    EXACT mode; the tables agree with the oracle's own generators to rounding (different j_l / 3j providers)."""
    wl = _tiny(**kw)
    oproj = oracle.build_projection(wl)
    monkeypatch.setenv("SONG_LOS_MODE", "exact")
    with dropin_lib.Harness(wl, None) as h:
        rows, xmin, xsize, dd, info = h.bessel2_tables()
        assert info["J_size"] == wl.layout.J_size and info["L_size"] == wl.layout.L_size and info["xx_size"] == wl.xx.size
        assert info["x_size_max_J"] == max(int(x) for x in xsize)
        n_same = 0
        for q, ref in enumerate(oproj.rows):
            if ref is None:
                assert rows[q] is None
                continue
            assert abs(int(xmin[q]) - int(oproj.index_xmin[q])) <= 1 and xsize[q] == wl.xx.size - xmin[q]
            if xmin[q] == oproj.index_xmin[q]:
                assert np.allclose(rows[q], ref, rtol=0, atol=1e-11 * max(np.abs(ref).max(), 1e-300))
                n_same += 1
        assert n_same > 0
        cubic = kw.get("bessels_interpolation") == "cubic"
        proj = oracle.OracleProjection(rows, xmin, xsize, None)
        if cubic:
            proj.ddrows = dd
            assert any(d is not None and np.any(d != 0) for d in dd)
        ref_T, _ = oracle.OracleRun(wl, proj).run()
        h.transfer2_init()
        got, count, labels_ok = h.read()
        assert np.array_equal(got, ref_T) and count == wl.n_integrals and labels_ok
        h.transfer2_free()
        if not cubic:
            # a second run on the same structures (page-locked blocks come back from the pool), tensor-core arithmetic
            monkeypatch.setenv("SONG_LOS_MODE", "fast")
            h.fill_sources()
            h.transfer2_init()
            got2, count2, _ = h.read()
            assert _rel_err(got2, ref_T) < 1e-9 and count2 == wl.n_integrals
            assert abs(h.checksum() - np.abs(got2).sum()) <= 1e-12 * np.abs(got2).sum()
            h.transfer2_free()


@pytest.mark.gpu
@pytest.mark.parametrize("cubic", [False, True], ids=["linear", "cubic"])
def test_bessel2_convolution_matches_oracle(oracle, cubic):
    """bessel2_convolution (bispectra2.c:1326) on the j_l1 table of the drop-in against the oracle's restatement of
    bessel2.c:499-605 on the oracle's own table; zero entries of f and g and x below the table are skipped."""
    wl = _tiny(bessels_interpolation="cubic" if cubic else "linear")
    oproj = oracle.build_projection(wl, with_jl1=True, spline=cubic)
    rng = np.random.default_rng(5)
    kk = np.sort(rng.uniform(1e-5, 0.9 * wl.xx_max / 14000.0, 200))
    dk = np.empty_like(kk); dk[1:-1] = kk[2:] - kk[:-2]; dk[0] = kk[1] - kk[0]; dk[-1] = kk[-1] - kk[-2]
    f = rng.normal(size=kk.size); g = rng.normal(size=kk.size)
    f[::7] = 0.0; g[3::11] = 0.0
    with dropin_lib.Harness(wl, None) as h:
        l1, x_min_l1 = h.l1_info()
        assert np.array_equal(l1, oproj.jl1["l1"])
        checked = 0
        for index_l1 in (0, 1, len(l1) // 2, len(l1) - 1):
            if x_min_l1[index_l1] != oproj.jl1["x_min_l1"][index_l1]:
                continue                      # the two j_l providers put the cut one grid point apart: different tables
            checked += 1
            ddj = None
            if cubic:                         # natural spline of the row (bessel2.c:439-463)
                row = oproj.jl1["j_l1"][index_l1]
                x = np.ascontiguousarray(wl.xx[oproj.jl1["index_xmin_l1"][index_l1]:][:row.size])
                ddj = np.zeros_like(row)
                oracle.lib().oracle_spline_table_columns(oracle._d(x), row.size, oracle._d(row), 1, oracle._d(ddj), 0)
            for r in (13000.0, 14000.0):
                for gg in (g, None):
                    got = h.bessel2_convolution(kk, dk, f, gg, index_l1, r)
                    ref = oracle.bessel2_convolution(kk, dk, f, gg, oproj.jl1["j_l1"][index_l1], ddj,
                                                     oproj.jl1["x_min_l1"][index_l1], wl.xx_step, r, cubic=cubic)
                    assert abs(got - ref) <= 1e-9 * max(abs(ref), 1e-30), (index_l1, r, got, ref)
        assert checked >= 2


@pytest.mark.gpu
def test_k1_level_memory_semantics(oracle, monkeypatch):
    """ptr2->transfer lives in per-batch slabs; transfer2_free_k1_level / transfer2_free keep their per-level meaning
    (has_allocated_transfers guard, count_allocated == count_memorised, transfer2.c:607-610, 2976-3104) when a batch
    covers parts of several k1 levels."""
    wl = _tiny(sources_k3_interpolation="cubic")
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    monkeypatch.setenv("SONG_LOS_MODE", "exact")
    monkeypatch.setenv("SONG_LOS_BATCH_UNITS", "97")          # batches end in the middle of k1 levels
    for pinned in ("1", "0"):
        monkeypatch.setenv("SONG_LOS_PINNED_TRANSFERS", pinned)
        got, count, labels_ok = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size))
        assert np.array_equal(got, ref) and count == wl.n_integrals and labels_ok


@pytest.mark.gpu
@pytest.mark.parametrize("devices", ["1", "8"])
def test_transfer2_store_to_disk_many_batches(oracle, tmp_path, monkeypatch, devices):
    """store_transfers_to_disk with many batches in flight on one or several GPUs: the writer thread puts every finished
    batch at its offset of the per-type files (the file order is the pair order, transfer2.c:2940-2957), levels are freed
    once written (:565), and the files read back through transfer2_load_transfers_from_disk equal the oracle."""
    wl = _tiny(sources_k3_interpolation="cubic")
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    monkeypatch.setenv("SONG_LOS_MODE", "exact")
    monkeypatch.setenv("SONG_LOS_BATCH_UNITS", "131")
    monkeypatch.setenv("SONG_LOS_CHUNKS_PER_DEVICE", "3")
    monkeypatch.setenv("SONG_LOS_DEVICES", devices)
    got, count, _ = dropin_lib.run_transfer2(wl, Projection(proj.rows, proj.index_xmin, proj.x_size), store_dir=str(tmp_path))
    assert np.array_equal(got, ref)
    for tt in (0, wl.layout.tt2_size - 1):
        f = tmp_path / f"transfers_{tt:03d}.dat"
        assert f.stat().st_size == 8 * wl.n_units and np.array_equal(np.fromfile(f), ref[tt])
