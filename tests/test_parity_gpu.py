"""GPU parity: the CUDA path, called through the C-ABI, against the CPU oracle.

EXACT mode must be bit-identical to the oracle (same statement order, no FMA contraction);
FAST mode must agree to BASELINE.json's tolerance: max relative error 1e-9 on transfer
functions above 1e-6 of their per-(l,m) maximum.
"""
import numpy as np
import pytest

from song_b200 import configs, workload
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9      # BASELINE.json north_star
FLOOR = 1e-6        # fraction of the per-(l,m) maximum below which values are not compared


def rel_err_above_floor(got, ref):
    """max over tt of max |got-ref|/|ref| restricted to |ref| > FLOOR * max_k |ref[tt]|."""
    worst = 0.0
    for tt in range(ref.shape[0]):
        mx = np.abs(ref[tt]).max()
        if mx == 0:
            assert np.all(got[tt] == 0)
            continue
        sel = np.abs(ref[tt]) > FLOOR * mx
        worst = max(worst, float(np.max(np.abs(got[tt][sel] - ref[tt][sel]) / np.abs(ref[tt][sel]))))
    return worst


def run_both(oracle, params, kind="smooth", mode=MODE_EXACT, spline_J=False):
    wl = workload.build_grids(params)
    workload.fill_sources(wl, kind=kind)
    oproj = oracle.build_projection(wl, spline=spline_J)
    ref, nterms = oracle.OracleRun(wl, oproj).run()
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size, oproj.ddrows))
    got = eng.transfer_pairs(wl, mode=mode)
    eng.close()
    return wl, ref, got


@pytest.mark.parametrize("k3_interp", ["linear", "cubic"])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_exact_bitwise_tiny(oracle, k3_interp, kind):
    p = configs.tiny(sources_k3_interpolation=k3_interp)
    wl, ref, got = run_both(oracle, p, kind=kind)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"


def test_exact_bitwise_temperature_only(oracle):
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p)
    assert np.array_equal(got, ref)


def test_exact_bitwise_no_extrapolation(oracle):
    p = configs.tiny(bispectra_k3_extrapolation="no_extrapolation", sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p)
    assert np.all(wl.phys_start == 0) and np.array_equal(wl.phys_size, wl.k3_tr_size)
    assert np.array_equal(got, ref)


def test_exact_bitwise_cubic_bessel_interpolation(oracle):
    p = configs.tiny(bessels_interpolation="cubic", sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p, spline_J=True)
    assert np.array_equal(got, ref)


def test_exact_bitwise_nonscalar_modes(oracle):
    """m = 0,1,2: E/B mixing through J_EB (transfer2.c:1960-2061)."""
    p = configs.tiny(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True)
    wl, ref, got = run_both(oracle, p)
    assert wl.layout.index_J_EB >= 0 and wl.layout.has_source_B
    assert np.array_equal(got, ref)


def test_exact_bitwise_quick_song_run(oracle):
    """BASELINE configs[0] at full size (6.5e5 integrals)."""
    wl, ref, got = run_both(oracle, configs.get("quick_song_run"))
    assert wl.n_integrals == 654112
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("time_interp", ["linear", "cubic"])
@pytest.mark.parametrize("sampling", ["custom", "bessel"])
def test_exact_bitwise_k_dependent_time_grids(oracle, sampling, time_interp):
    """transfer2_tau_sampling = custom / bessel: transfer2_get_time_grid inserts points so that no step exceeds
    2 pi/k tau_linstep_song (transfer2.c:2422) or xx_step/k (:2432); the sources are resampled in time, linearly or
    with a cubic spline (transfer2.c:2816-2880).  Both arithmetic modes route to the exact kernels."""
    # tau_linstep_song small enough that 2 pi/k * tau_linstep_song (k3 ~ 0.004-0.008 here) falls below the source spacing
    # (bessel sampling: xx_step/k below the source spacing needs a finer x grid than the default 0.2)
    p = configs.tiny(transfer2_tau_sampling=sampling, sources_time_interpolation=time_interp,
                     sources_k3_interpolation="cubic", tau_linstep_song=0.003, bessel_x_step_song=0.04)
    for mode in (MODE_EXACT, MODE_FAST):
        wl, ref, got = run_both(oracle, p, mode=mode)
        assert np.array_equal(got, ref), f"max abs diff {np.abs(got - ref).max()}"
    # the grids really are k dependent: extra integration points between the source times for the largest k3
    import ctypes as C
    s = oracle.OracleRun(wl, oracle.build_projection(wl)).setup
    kmax = float(wl.k3_tr.max())
    L = oracle.lib()
    nmax = L.oracle_time_grid_size_max(C.byref(s), kmax)
    grid = np.zeros(nmax); d = np.zeros(nmax); w = np.zeros(nmax); left = np.zeros(nmax, dtype=np.int32); n = C.c_int(0)
    assert L.oracle_get_time_grid(C.byref(s), kmax, grid.ctypes.data_as(oracle.dp), left.ctypes.data_as(oracle.ip),
                                  d.ctypes.data_as(oracle.dp), w.ctypes.data_as(oracle.dp), C.byref(n)) == 0
    assert n.value > wl.tau.size, (n.value, wl.tau.size)


# ----------------------------------------------------------------------------------------------
# FAST mode (k3-sorted DMMA kernel): BASELINE.json tolerance
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k3_interp", ["linear", "cubic"])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_fast_tolerance_tiny(oracle, k3_interp, kind):
    p = configs.tiny(sources_k3_interpolation=k3_interp)
    wl, ref, got = run_both(oracle, p, kind=kind, mode=MODE_FAST)
    err = rel_err_above_floor(got, ref)
    assert err < REL_TOL, err


def test_fast_tolerance_temperature_only(oracle):
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p, mode=MODE_FAST)
    assert rel_err_above_floor(got, ref) < REL_TOL


def test_fast_tolerance_no_extrapolation(oracle):
    p = configs.tiny(bispectra_k3_extrapolation="no_extrapolation", sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p, mode=MODE_FAST)
    assert rel_err_above_floor(got, ref) < REL_TOL


def test_fast_tolerance_nonscalar_modes(oracle):
    p = configs.tiny(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True)
    wl, ref, got = run_both(oracle, p, mode=MODE_FAST)
    assert rel_err_above_floor(got, ref) < REL_TOL


def test_fast_tolerance_quick_song_run(oracle):
    wl, ref, got = run_both(oracle, configs.get("quick_song_run"), mode=MODE_FAST)
    assert rel_err_above_floor(got, ref) < REL_TOL


@pytest.mark.parametrize("shape", [None, "12,1,7", "12,1,10"])
def test_fast_many_columns_layouts(oracle, monkeypatch, shape):
    """Column counts that exercise the (WM,WN) = (4,2) and (2,4) warp grids and partial sub-blocks; with a forced
    (12,1,x) grid the wide blocks are covered in several column passes (the choice made for large batches)."""
    if shape:
        monkeypatch.setenv("SONG_LOS_TC_SHAPE", shape)
    for lmax, linstep in ((300, 4), (700, 4)):
        p = configs.tiny(l_max_scalars=lmax, l_linstep=linstep, l_logstep=1.05, sources_k3_interpolation="cubic",
                         k_max_tau0_over_l_max_song=0.25)
        wl, ref, got = run_both(oracle, p, mode=MODE_FAST)
        assert wl.layout.l_size > 64
        assert rel_err_above_floor(got, ref) < REL_TOL


def test_fast_several_k_chunks(oracle):
    """l_max_los = 14: 15 (T) and 13 (E) terms -> two K chunks per column block, i.e. two kernel launches per block,
    the second one continuing from the partial sums of the first (the path the 5 % BASELINE workloads take)."""
    p = configs.tiny(l_max_los_t=14, l_max_los_p=14, sources_k3_interpolation="cubic")
    for kind in ("smooth", "noise"):
        wl, ref, got = run_both(oracle, p, kind=kind, mode=MODE_FAST)
        assert wl.layout.L_size == 15
        assert rel_err_above_floor(got, ref) < REL_TOL


def test_fast_rejects_cubic_bessel_interpolation(oracle):
    p = configs.tiny(bessels_interpolation="cubic")
    from song_b200._lib import SongLosError
    with pytest.raises(SongLosError, match="linear Bessel interpolation"):
        run_both(oracle, p, mode=MODE_FAST, spline_J=True)


def test_long_coarse_k3_grids(oracle):
    """k3_size above 256 (a high-resolution sources2 sampling): the spline workspaces move from thread-local memory to
    global scratch; the reference's array_spline_table_columns has no size limit (transfer2.c:2639)."""
    p = configs.tiny(k3_size_min=300, sources_k3_interpolation="cubic")
    wl, ref, got = run_both(oracle, p, mode=MODE_EXACT)
    assert wl.k3_pt_size.max() > 256
    assert np.array_equal(got, ref)
    wl, ref, got = run_both(oracle, p, kind="noise", mode=MODE_FAST)
    assert rel_err_above_floor(got, ref) < REL_TOL


@pytest.mark.parametrize("n_bands", [2, 3])
def test_k3_band_slices_match_whole_pairs(oracle, n_bands):
    """(pair groups x k3 bands) schedule, song_b200/shard.py: every GPU gets slices of the pairs' fine grids.  The slices
    of all bands, computed in separate batches, must reproduce the oracle on the whole pairs -- including the wings of
    the flat extrapolation (transfer2.c:2743-2757), whose sources are interpolated at the first / last PHYSICAL k3 of
    the pair, a point that usually lies in another band (song_los_batch.phys_k_first / phys_k_last)."""
    from song_b200 import shard
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl, kind="noise")
    assert np.any(wl.phys_start > 0) and np.any(wl.phys_start + wl.phys_size < wl.k3_tr_size)     # there are wings on both sides
    oproj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, oproj).run()
    sched = shard.schedule_bands(wl, 2 * n_bands, n_bands=n_bands)                                # two pair groups as well
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size, oproj.ddrows))
    slices_without_their_physical_end = 0
    for mode in (MODE_EXACT, MODE_FAST):
        full = np.full_like(ref, np.nan)
        for rank in range(2 * n_bands):
            sub, own, first, last = shard.shard_units(wl, rank, 2 * n_bands, schedule=sched)
            slices_without_their_physical_end += int(np.sum((sub.phys_start < 0) | (sub.phys_start + sub.phys_size > sub.k3_tr_size)))
            got = eng.transfer_pairs(sub, mode=mode)
            o = 0
            for ip, a, b in zip(own, first, last):
                t = int(wl.tr_off[ip])
                full[:, t + a:t + b] = got[:, o:o + (b - a)]
                o += b - a
        assert not np.isnan(full).any()
        if mode == MODE_EXACT:
            assert np.array_equal(full, ref)
        else:
            assert rel_err_above_floor(full, ref) < REL_TOL
    assert slices_without_their_physical_end > 0          # the case the two extra arrays exist for
    eng.close()
