"""the k3 integral of the intrinsic bispectrum (bispectra2_intrinsic_integrate_over_k3,
source/bispectra2.c:1107-1439, inner loop bessel2_convolution, source/bessel2.c:499-605) on the transfer functions left
in device memory by the LOS kernels, against the oracle's restatement: bit for bit (same statement order, no FMA
contraction) when both sides read the same T and the same j_l1 table."""
import numpy as np
import pytest

from song_b200 import configs, workload
from song_b200.bessel2 import Bessel2
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection

pytestmark = pytest.mark.gpu


def r_grid(wl, n):
    """pwb->r for bispectra_r_sampling: around tau0 - tau_rec (the region where the integrand lives)."""
    c = wl.cosmo["tau0"] - wl.cosmo["tau_rec"]
    return np.linspace(c - 400.0, c + 250.0, n)


def natural_spline(oracle, x, y):
    dd = np.zeros_like(y)
    if y.size >= 3:
        oracle.lib().oracle_spline_table_columns(oracle._d(np.ascontiguousarray(x)), y.size, oracle._d(y), 1, oracle._d(dd), 0)
    return dd


@pytest.mark.parametrize("kw,mode", [(dict(sources_k3_interpolation="cubic"), MODE_EXACT),
                                     (dict(sources_k3_interpolation="cubic"), MODE_FAST),
                                     (dict(modes_song=[0, 1, 2], has_source_B=True), MODE_EXACT),
                                     (dict(bessels_interpolation="cubic", sources_k3_interpolation="cubic"), MODE_EXACT)],
                         ids=["scalar_exact", "scalar_fast", "m012", "cubic_bessel"])
def test_k3_integral_bitwise(oracle, kw, mode):
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)))
    cubic = kw.get("bessels_interpolation") == "cubic"
    oproj = oracle.build_projection(wl, with_jl1=True, spline=cubic)
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size, oproj.ddrows))
    T = eng.transfer_pairs(wl, mode=mode)                 # what the device integrates is exactly what it returned here
    jl = oproj.jl1
    ddj = None
    if cubic:
        ddj = [natural_spline(oracle, wl.xx[jl["index_xmin_l1"][i]:][:jl["j_l1"][i].size], jl["j_l1"][i]) for i in range(len(jl["l1"]))]
    eng.set_bessel_table(jl["j_l1"], jl["x_min_l1"], ddj)
    lay = wl.layout
    r = r_grid(wl, 37)
    # every (l3, M3) of the T block and of the E block: L3 = |l3 - |M3|| + offset_L3 with offset 0 (bispectra2.c:1183)
    tts, rows, coefs, Ms = [], [], [], []
    for base in ([lay.index_tt2_T] if lay.has_source_T else []) + ([lay.index_tt2_E] if lay.has_source_E else []):
        for il, l3 in enumerate(lay.l):
            for im, M3 in enumerate(lay.m):
                if M3 > l3:
                    continue
                L3 = abs(int(l3) - int(M3))
                tts.append(base + int(lay.lm_array[il, im])); rows.append(int(jl["index_l1"][L3])); coefs.append(1.0 + 0.25 * M3); Ms.append(int(M3))
    tts, rows, coefs, Ms = map(np.asarray, (tts, rows, coefs, Ms))
    checked = 0
    for M3 in np.unique(Ms):
        sel = Ms == M3
        got = eng.integrate_k3(tts[sel], rows[sel], r, wl.n_pairs, abs_M3=int(M3), M3_coefficient=coefs[sel])
        for q, (tt, row, c) in enumerate(zip(tts[sel], rows[sel], coefs[sel])):
            ref = oracle.integral_over_k3(wl, T[tt], jl["j_l1"][row], None if ddj is None else ddj[row], jl["x_min_l1"][row], r,
                                          abs_M3=int(M3), M3_coefficient=float(c), cubic=cubic)
            assert np.array_equal(got[q], ref), (int(tt), float(np.abs(got[q] - ref).max()))
            checked += int(np.any(ref != 0))
    assert checked > 0
    eng.close()


def test_k3_integral_on_a_baseline_workload(oracle):
    """Sampled pairs of configs[1] through the product path only (device-built J tables, tensor-core LOS), then the k3
    integral of every l3 for r_size = 75 (pre/sn_t_10percent_lmax2000.pre:47): bit-identical to the oracle loop fed the
    same T and the product's j_l1 table."""
    grids = workload.build_grids(configs.get("sn_t_10percent_lmax2000"))
    sub = workload.fill_sources(grids.select(np.unique(np.linspace(0, grids.n_pairs - 1, 24).astype(int))))
    eng = LosEngine(sub, device=0)
    b2 = Bessel2(sub).init(eng)
    ht = b2.host_tables()
    T = eng.transfer_pairs(sub, mode=MODE_FAST)
    eng.set_bessel_table(ht["j_l1"], ht["x_min_l1"])
    lay = sub.layout
    index_l1 = {int(l1): i for i, l1 in enumerate(ht["l1"])}
    r = r_grid(sub, 75)
    tts = np.asarray([lay.index_tt2_T + int(lay.lm_array[il, 0]) for il in range(lay.l_size)])
    rows = np.asarray([index_l1[int(l3)] for l3 in lay.l])
    got = eng.integrate_k3(tts, rows, r, sub.n_pairs)
    assert got.shape == (lay.l_size, 75, sub.n_pairs) and eng.k3_ms > 0
    for q in (0, 1, lay.l_size // 2, lay.l_size - 1):
        ref = oracle.integral_over_k3(sub, T[tts[q]], ht["j_l1"][rows[q]], None, ht["x_min_l1"][rows[q]], r)
        assert np.array_equal(got[q], ref) and np.any(ref != 0)
    eng.close()


def test_k3_integral_error_convention(oracle):
    """x = k3 r beyond the j_l1 table is the reference's DEBUG test (bessel2.c:520-523): _FAILURE_ + message."""
    from song_b200._lib import SongLosError
    wl = workload.fill_sources(workload.build_grids(configs.tiny()))
    oproj = oracle.build_projection(wl, with_jl1=True)
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size))
    with pytest.raises(SongLosError, match="no transfer functions"):
        eng.integrate_k3([0], [0], [14000.0], wl.n_pairs)
    eng.transfer_pairs(wl, mode=MODE_EXACT)
    with pytest.raises(SongLosError, match="j_l1 table not set"):
        eng.integrate_k3([0], [0], [14000.0], wl.n_pairs)
    eng.set_bessel_table(oproj.jl1["j_l1"], oproj.jl1["x_min_l1"])
    with pytest.raises(SongLosError, match="x_max"):
        eng.integrate_k3([0], [0], [1e6], wl.n_pairs)
    eng.close()


# ---------------------------------------------------------------------------------------------------------------
# spectra2_cls (source/spectra2.c:241-570) on the device-resident transfer functions
# ---------------------------------------------------------------------------------------------------------------
def primordial(k, A_s=2.1e-9, n_s=0.965, k_pivot=0.05):
    """primordial_spectrum_at_k for the analytic power law of CLASS (the module itself is outside the path)."""
    return A_s * (k / k_pivot) ** (n_s - 1.0)


@pytest.mark.parametrize("kw", [dict(sources_k3_interpolation="cubic"), dict(modes_song=[0, 1, 2], has_source_B=True)],
                         ids=["scalar", "m012"])
def test_intrinsic_cl_sum(oracle, kw):
    """TT, EE, TE (and BB for m != 0) of every l, m against the oracle's restatement of the reference loops: the same
    products in the same k3 order per pair, pairs added in the reference order; not one running sum, hence 1e-12."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**kw)))
    oproj = oracle.build_projection(wl)
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size))
    T = eng.transfer_pairs(wl, mode=MODE_EXACT)
    lay = wl.layout
    prim = primordial(wl.k)
    tt1, tt2, ms = [], [], []
    blocks = [(lay.index_tt2_T, lay.index_tt2_T), (lay.index_tt2_E, lay.index_tt2_E), (lay.index_tt2_T, lay.index_tt2_E)]
    if lay.has_source_B:
        blocks.append((lay.index_tt2_B, lay.index_tt2_B))
    for a, b in blocks:
        for il, l in enumerate(lay.l):
            for im, m in enumerate(lay.m):
                if m > l or (a == getattr(lay, "index_tt2_B", -1) and lay.has_source_B and m == 0):
                    continue
                tt1.append(a + int(lay.lm_array[il, im])); tt2.append(b + int(lay.lm_array[il, im])); ms.append(int(m))
    for rescale in (True, False):
        got = eng.cl_sum(wl, tt1, tt2, ms, prim, rescale_cmb_sources=rescale)
        ref = np.asarray([oracle.spectra2_cl(wl, T[a], T[b], m, prim, rescale) for a, b, m in zip(tt1, tt2, ms)])
        assert np.any(ref != 0)
        scale = np.abs(ref).max()
        assert np.max(np.abs(got - ref)) <= 1e-12 * scale, float(np.max(np.abs(got - ref)) / scale)
        m0 = np.asarray(ms) == 0
        assert np.max(np.abs(got[m0] - ref[m0]) / np.maximum(np.abs(ref[m0]), 1e-300)) < 1e-12
    eng.close()


def test_intrinsic_cl_sum_is_additive_over_batches(oracle):
    """The caller adds the contributions of batches / GPUs: two halves of the pairs give the whole."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")))
    oproj = oracle.build_projection(wl)
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size))
    lay = wl.layout
    prim = primordial(wl.k)
    tt = [lay.index_tt2_T + int(lay.lm_array[il, 0]) for il in range(lay.l_size)]
    ms = [0] * len(tt)
    eng.transfer_pairs(wl, mode=MODE_FAST)
    whole = eng.cl_sum(wl, tt, tt, ms, prim)
    parts = np.zeros_like(whole)
    for idx in (np.arange(0, wl.n_pairs // 2), np.arange(wl.n_pairs // 2, wl.n_pairs)):
        sub = wl.select(idx)
        eng.transfer_pairs(sub, mode=MODE_FAST)
        parts += eng.cl_sum(sub, tt, tt, ms, prim)
    assert np.max(np.abs(parts - whole)) <= 1e-12 * np.abs(whole).max()
    eng.close()
