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
