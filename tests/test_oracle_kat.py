"""Known-answer tests that pin the CPU oracle.

The reference ships no golden vectors for this path.  Its own transfer2.c and bessel2.c, compiled where they lie, pin
the oracle's restatement of those two files bit for bit (tests/test_reference_pin.py); the helpers of its absent CLASS
fork (splines, 3j, j_l, index maps) cannot be pinned that way, so they -- and the LOS sum once more -- are pinned here by
answers that depend on neither implementation: closed forms, exact rational arithmetic, scipy special functions and an
independent vectorised numpy restatement of the LOS sum.
"""
import ctypes as C
import math
from fractions import Fraction

import numpy as np
import pytest
from scipy.interpolate import CubicSpline
from scipy.special import spherical_jn

from song_b200 import configs, workload


# ----------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------
def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def spline_dd(oracle, x, y, mode=1):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(np.atleast_2d(y), dtype=np.float64)
    dd = np.zeros_like(y)
    rc = oracle.lib().oracle_spline_table_columns(_d(x), x.size, _d(y), y.shape[0], _d(dd), mode)
    assert rc == 0
    return dd


def wigner3j_exact(j1, j2, j3, m1, m2, m3):
    """Racah formula in exact rational arithmetic (integers only)."""
    if m1 + m2 + m3 != 0 or j3 < abs(j1 - j2) or j3 > j1 + j2:
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    f = math.factorial
    delta = Fraction(f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3), f(j1 + j2 + j3 + 1))
    pref = delta * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    s = Fraction(0)
    for t in range(0, j1 + j2 + j3 + 1):
        args = [t, j3 - j2 + t + m1, j3 - j1 + t - m2, j1 + j2 - j3 - t, j1 - t - m1, j2 - t + m2]
        if min(args) < 0:
            continue
        den = 1
        for a in args:
            den *= f(a)
        s += Fraction((-1) ** t, den)
    sign = (-1) ** (j1 - j2 - m3)
    # sqrt of an exact rational: do it in high-precision floats via integer sqrt
    num, den = pref.numerator, pref.denominator
    root = math.sqrt(num / den) if num < 10 ** 300 and den < 10 ** 300 else math.exp(0.5 * (math.log(num) - math.log(den)))
    return sign * float(s) * root


def drc3jj(oracle, l2, l3, m2, m3, ndim=512):
    lo, hi = C.c_double(0), C.c_double(0)
    out = np.zeros(ndim)
    rc = oracle.lib().oracle_drc3jj(float(l2), float(l3), float(m2), float(m3), C.byref(lo), C.byref(hi), _d(out), ndim)
    assert rc == 0, oracle.lib().oracle_last_error()
    n = int(round(hi.value - lo.value)) + 1
    return int(round(lo.value)), out[:n]


# ----------------------------------------------------------------------------------------------
# CLASS-fork helpers restated in oracle_tools.c
# ----------------------------------------------------------------------------------------------
def test_spline_reproduces_quadratic(oracle):
    """A clamped spline whose end slopes come from the 3-point parabola is exact for quadratics:
    y'' = 2a everywhere."""
    x = np.array([0.1, 0.25, 0.3, 0.55, 0.9, 1.4, 1.5])
    y = 3.0 * x * x - 2.0 * x + 0.7
    dd = spline_dd(oracle, x, y)
    assert np.allclose(dd, 6.0, rtol=0, atol=1e-10)


def test_spline_linear_source_is_exact(oracle):
    """KAT (4): a source linear in k3 has zero second derivatives, so the cubic formula of
    transfer2.c:2703 reduces to the linear one."""
    x = np.sort(np.random.default_rng(1).uniform(0, 1, 9))
    dd = spline_dd(oracle, x, 4.0 * x - 1.0)
    assert np.abs(dd).max() < 1e-9


def test_spline_matches_scipy_clamped(oracle):
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(0, 2, 12))
    y = np.sin(3 * x) + 0.3 * x
    dd = spline_dd(oracle, x, y)[0]

    def parabola_slope(xs, ys):   # slope at xs[0] of the parabola through three points
        return ((xs[2] - xs[0]) ** 2 * (ys[1] - ys[0]) - (xs[1] - xs[0]) ** 2 * (ys[2] - ys[0])) / \
               ((xs[2] - xs[0]) * (xs[1] - xs[0]) * (xs[2] - xs[1]))
    d0 = parabola_slope(x[:3], y[:3])
    d1 = parabola_slope(x[::-1][:3], y[::-1][:3])
    cs = CubicSpline(x, y, bc_type=((1, d0), (1, d1)))
    assert np.allclose(dd, cs(x, 2), rtol=1e-9, atol=1e-9)


def test_spline_natural_mode(oracle):
    x = np.linspace(0, 1, 8)
    y = np.cos(2 * x)
    dd = spline_dd(oracle, x, y, mode=0)[0]
    cs = CubicSpline(x, y, bc_type="natural")
    assert dd[0] == 0 and abs(dd[-1]) < 1e-14
    assert np.allclose(dd, cs(x, 2), atol=1e-10)


def test_lin_space_endpoints(oracle):
    out = np.zeros(7)
    oracle.lib().oracle_lin_space.argtypes = [C.POINTER(C.c_double), C.c_double, C.c_double, C.c_int]
    oracle.lib().oracle_lin_space(_d(out), 0.0, 3360.21, 7)
    assert out[0] == 0.0 and out[-1] == 3360.21 and np.all(np.diff(out) > 0)


@pytest.mark.parametrize("case", [(2, 3, 1, -1), (4, 4, 2, -2), (5, 2, 0, 0), (3, 7, -2, 1), (6, 6, 0, 0),
                                  (8, 30, 2, -2), (12, 40, -1, 1), (0, 5, 0, 2), (24, 24, 2, -2)])
def test_drc3jj_against_exact_racah(oracle, case):
    l2, l3, m2, m3 = case
    l1min, vals = drc3jj(oracle, l2, l3, m2, m3)
    assert l1min == max(abs(l2 - l3), abs(m2 + m3))
    for i, v in enumerate(vals):
        exact = wigner3j_exact(l1min + i, l2, l3, -m2 - m3, m2, m3)
        assert abs(v - exact) < 1e-12, (case, l1min + i, v, exact)


@pytest.mark.parametrize("L,l,S", [(8, 2000, 0), (8, 2000, 2), (24, 4000, 2), (6, 37, 2), (17, 3000, 0), (5, 2, 2)])
def test_drc3jj_song_call_sites(oracle, L, l, S):
    """The two calls of bessel2.c:1131,1154: normalisation sum (2 l1+1) f^2 = 1, parity zeros and
    the closed form for L=0."""
    l1min, f = drc3jj(oracle, L, l, S, -S)
    l1 = l1min + np.arange(f.size)
    assert abs(np.sum((2 * l1 + 1) * f * f) - 1.0) < 1e-12
    if S == 0:
        assert np.all(f[(l1 + L + l) % 2 == 1] == 0.0)      # (l1 L l;000) vanishes for odd sum
    # symmetry: (l1 L l; 0 S -S) = (-1)^(l1+L+l) (l1 L l; 0 -S S)
    _, g = drc3jj(oracle, L, l, -S, S)
    assert np.allclose(g, ((-1.0) ** (l1 + L + l)) * f, atol=1e-13)


def test_drc3jj_L0_closed_form(oracle):
    for l in (2, 17, 350):
        l1min, f = drc3jj(oracle, 0, l, 0, 0)
        assert l1min == l and f.size == 1
        assert abs(f[0] - (-1.0) ** l / math.sqrt(2 * l + 1)) < 1e-15


def test_spherical_bessel_against_scipy(oracle):
    L = oracle.lib()
    for l in (0, 1, 2, 7, 30, 131, 500, 2024):
        for x in (0.0, 1e-3, 0.7, 5.3, 29.9, 30.0, 131.5, 480.0, 1999.3, 3360.0):
            got = L.oracle_sph_bessel_j(l, x)
            ref = float(spherical_jn(l, x))
            assert abs(got - ref) <= 2e-12 * max(abs(ref), 1e-300) + 1e-16, (l, x, got, ref)


def test_lm_offsets(oracle):
    L = oracle.lib()
    l = np.array([2, 3, 5, 9], dtype=np.int32)
    m = np.array([0, 1, 3], dtype=np.int32)
    n = L.oracle_size_indexl_indexm(_i(l), 4, _i(m), 3)
    assert n == 2 + 3 + 3 + 3                       # l=2 cannot have m=3
    seen = []
    for off in range(n):
        il, im = C.c_int(0), C.c_int(0)
        assert L.oracle_offset2multipole_indexl_indexm(off, _i(l), 4, _i(m), 3, C.byref(il), C.byref(im)) == 0
        assert L.oracle_multipole2offset_indexl_indexm(int(l[il.value]), int(m[im.value]), _i(l), 4, _i(m), 3) == off
        seen.append((il.value, im.value))
    assert seen == sorted(seen)
    # lm(L,m) for m-list {0}: lm(L,0) = L   
    m0 = np.array([0], dtype=np.int32)
    assert [L.oracle_multipole2offset_l_indexm(Lv, 0, _i(m0), 1) for Lv in range(6)] == list(range(6))
    # workload.make_layout agrees with the oracle's restatement
    lay = workload.make_layout(l, m, 4, 4, True, True, True)
    for il in range(4):
        for im in range(3):
            assert lay.lm_array[il, im] == L.oracle_multipole2offset_indexl_indexm(int(l[il]), int(m[im]), _i(l), 4, _i(m), 3)
    for Lv in range(lay.L_size):
        for im in range(3):
            assert lay.lm_src[Lv, im] == L.oracle_multipole2offset_l_indexm(Lv, int(m[im]), _i(m), 3)


# ----------------------------------------------------------------------------------------------
# bessel2 restatement
# ----------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def tiny_proj(oracle, tiny_workload):
    return oracle.build_projection(tiny_workload, with_jl1=True)


def test_J_TT_L0_is_2lp1_jl(oracle, tiny_workload, tiny_proj):
    """KAT (1): J^TT_{0,l,0}(x) = (2l+1) j_l(x)  (bessel2.c:1543-1555 with L=0)."""
    wl, lay = tiny_workload, tiny_workload.layout
    for il in range(lay.l_size):
        jidx = ((lay.index_J_TT * lay.L_size + 0) * lay.l_size + il) * lay.m_size
        row, xm = tiny_proj.rows[jidx], tiny_proj.index_xmin[jidx]
        x = wl.xx[xm:xm + row.size]
        ref = (2 * lay.l[il] + 1) * spherical_jn(int(lay.l[il]), x)
        assert np.allclose(row, ref, rtol=1e-11, atol=1e-13)
        # the cut: everything before index_xmin is below J_Llm_cut
        if xm > 0:
            below = (2 * lay.l[il] + 1) * spherical_jn(int(lay.l[il]), wl.xx[:xm + 1])
            assert np.all(np.abs(below) < wl.params["bessel_J_cut_song"] * 1.0000001)


def test_J_selection_rules(oracle, tiny_workload, tiny_proj):
    """KAT (5): J_EE vanishes for L<2 (spin constraint, bessel2.c:1090): marker row of size 1."""
    wl, lay = tiny_workload, tiny_workload.layout
    for Lv in (0, 1):
        for il in range(lay.l_size):
            jidx = ((lay.index_J_EE * lay.L_size + Lv) * lay.l_size + il) * lay.m_size
            assert tiny_proj.x_size[jidx] == 1 and tiny_proj.index_xmin[jidx] == wl.xx.size - 1
            assert tiny_proj.rows[jidx][0] == 0.0
    # L == l rows start at x = 0 (bessel2.c:1202-1206)
    for il, lv in enumerate(lay.l):
        if lv < lay.L_size:
            jidx = ((lay.index_J_TT * lay.L_size + int(lv)) * lay.l_size + il) * lay.m_size
            assert tiny_proj.index_xmin[jidx] == 0


def test_J_EB_vanishes_for_m0(oracle):
    p = configs.tiny(modes_song=[0, 1], has_source_B=True)
    wl = workload.build_grids(p)
    proj = oracle.build_projection(wl)
    lay = wl.layout
    assert lay.index_J_EB >= 0
    for Lv in range(2, lay.L_size):
        for il in range(lay.l_size):
            jidx = ((lay.index_J_EB * lay.L_size + Lv) * lay.l_size + il) * lay.m_size + 0   # m = 0
            assert proj.x_size[jidx] == 1 and proj.rows[jidx][0] == 0.0                      # bessel2.c:1093
            jidx1 = jidx + 1                                                                 # m = 1
            assert proj.x_size[jidx1] > 1 and np.abs(proj.rows[jidx1]).max() > 0


def test_J_TT_general_L_against_direct_sum(oracle, tiny_workload, tiny_proj):
    """J_TT for L>0 against eq. 5.97 evaluated with exact 3j symbols and scipy Bessels."""
    wl, lay = tiny_workload, tiny_workload.layout
    for Lv, il in ((1, 2), (2, 5), (4, lay.l_size - 1)):
        l = int(lay.l[il])
        jidx = ((lay.index_J_TT * lay.L_size + Lv) * lay.l_size + il) * lay.m_size
        row, xm = tiny_proj.rows[jidx], tiny_proj.index_xmin[jidx]
        x = wl.xx[xm:xm + row.size][5::11]
        ref = np.zeros_like(x)
        for l1 in range(abs(l - Lv), l + Lv + 1):
            if (l - l1 - Lv) % 2:
                continue
            tj = wigner3j_exact(l1, Lv, l, 0, 0, 0)
            ref += (-1.0) ** ((l - l1 - Lv) // 2) * (2 * l1 + 1) * tj * tj * spherical_jn(l1, x)
        ref *= (2 * l + 1)
        assert np.allclose(row[5::11], ref, rtol=1e-9, atol=1e-12)


# ----------------------------------------------------------------------------------------------
# transfer2 restatement against an independent numpy restatement of appendix C
# ----------------------------------------------------------------------------------------------
def numpy_los(wl, proj, pair):
    """T_l(k3) for one pair, temperature only, m=0, linear k3 interpolation, linear J --
    vectorised numpy, written from SURVEY.md appendix C rather than from the oracle."""
    lay = wl.layout
    o, n = int(wl.pt_off[pair]), int(wl.k3_pt_size[pair])
    kp = wl.k3_pt[o:o + n]
    t0, nt = int(wl.tr_off[pair]), int(wl.k3_tr_size[pair])
    kt = wl.k3_tr[t0:t0 + nt]
    first, last = int(wl.phys_start[pair]), int(wl.phys_start[pair] + wl.phys_size[pair] - 1)
    ntau = wl.tau.size
    S = wl.src[wl.src_off[pair]:wl.src_off[pair] + lay.tp2_size * ntau * n].reshape(lay.tp2_size, ntau, n)
    ke = kt[np.clip(np.arange(nt), first, last)]
    Sk = np.empty((lay.tp2_size, ntau, nt))
    for tp in range(lay.tp2_size):
        for it in range(ntau):
            Sk[tp, it] = np.interp(ke, kp, S[tp, it])
    d = wl.cosmo["tau0"] - wl.tau
    w = np.empty(ntau)
    w[0] = wl.tau[1] - wl.tau[0]
    w[1:-1] = wl.tau[2:] - wl.tau[:-2]
    w[-1] = wl.tau[-1] - wl.tau[-2]
    x = kt[:, None] * d[None, :]
    ix = (x / wl.xx_step).astype(np.int64)
    a = (wl.xx[ix + 1] - x) / wl.xx_step
    out = np.zeros((lay.l_size, nt))
    for il in range(lay.l_size):
        g = np.zeros_like(x)
        for Lv in range(lay.l_max_los_t + 1):
            jidx = ((lay.index_J_TT * lay.L_size + Lv) * lay.l_size + il) * lay.m_size
            row, xm = proj.rows[jidx], int(proj.index_xmin[jidx])
            full = np.zeros(wl.xx.size + 1)
            full[xm:xm + row.size] = row
            ok = ix >= xm
            Jv = np.where(ok, a * full[ix] + (1 - a) * full[ix + 1], 0.0)
            g += Jv * Sk[lay.index_tp2_T + Lv].T
        out[il] = 0.5 * (g * w[None, :]).sum(axis=1) / (8.0 * (2 * lay.l[il] + 1))
    return out


def test_oracle_against_numpy_restatement(oracle):
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="linear")
    wl = workload.fill_sources(workload.build_grids(p))
    proj = oracle.build_projection(wl)
    ref, nterms = oracle.OracleRun(wl, proj).run()
    assert nterms > 0
    scale = np.abs(ref).max()
    for pair in (0, 7, wl.n_pairs - 1):
        mine = numpy_los(wl, proj, pair)
        t0, nt = int(wl.tr_off[pair]), int(wl.k3_tr_size[pair])
        assert np.abs(mine - ref[:, t0:t0 + nt]).max() < 1e-13 * scale


def test_first_order_limit(oracle):
    """KAT (2): with only S_00 != 0 the result is the first-order LOS form
    T_l = (1/8) int dtau S_00 j_l(k(tau0-tau)), evaluated here with scipy's j_l on the same
    trapezoid grid (the table lerp error is O(xx_step^2))."""
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="linear", bessel_x_step_song=0.02,
                     bessel_J_cut_song=1e-14)
    wl = workload.fill_sources(workload.build_grids(p))
    lay = wl.layout
    ntau = wl.tau.size
    for ip in range(wl.n_pairs):           # zero every multipole except L = 0
        n = int(wl.k3_pt_size[ip])
        blk = wl.src[wl.src_off[ip]:wl.src_off[ip] + lay.tp2_size * ntau * n].reshape(lay.tp2_size, ntau, n)
        blk[1:] = 0.0
        blk[0] = 1.0 + 0.0 * blk[0]        # S_00 = 1: the k3 interpolation is then exact
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    d = wl.cosmo["tau0"] - wl.tau
    w = np.empty(ntau)
    w[0] = wl.tau[1] - wl.tau[0]; w[1:-1] = wl.tau[2:] - wl.tau[:-2]; w[-1] = wl.tau[-1] - wl.tau[-2]
    pair = wl.n_pairs - 1
    t0, nt = int(wl.tr_off[pair]), int(wl.k3_tr_size[pair])
    kt = wl.k3_tr[t0:t0 + nt]
    for il in (0, 4, lay.l_size - 1):
        exact = 0.5 * (spherical_jn(int(lay.l[il]), kt[:, None] * d[None, :]) * w[None, :]).sum(axis=1) / 8.0
        assert np.allclose(ref[il, t0:t0 + nt], exact, rtol=0, atol=2e-4 * np.abs(exact).max())


def test_impulse_source(oracle):
    """KAT (3): a source that is non-zero at a single time gives T = (1/2) dtau J(x) S / (8(2l+1))
    exactly, including the index_x < index_xmin_J skip (transfer2.c:2232)."""
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="linear")
    wl = workload.fill_sources(workload.build_grids(p))
    lay = wl.layout
    ntau, it0 = wl.tau.size, 5
    for ip in range(wl.n_pairs):
        n = int(wl.k3_pt_size[ip])
        blk = wl.src[wl.src_off[ip]:wl.src_off[ip] + lay.tp2_size * ntau * n].reshape(lay.tp2_size, ntau, n)
        blk[...] = 0.0
        blk[2, it0, :] = 1.0               # S_20(tau_5) = 1 for every k3
    proj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, proj).run()
    dtau = wl.tau[it0 + 1] - wl.tau[it0 - 1]
    d = wl.cosmo["tau0"] - wl.tau[it0]
    skipped = 0
    for pair in range(wl.n_pairs):
        t0, nt = int(wl.tr_off[pair]), int(wl.k3_tr_size[pair])
        for il in range(lay.l_size):
            jidx = ((lay.index_J_TT * lay.L_size + 2) * lay.l_size + il) * lay.m_size
            row, xm = proj.rows[jidx], int(proj.index_xmin[jidx])
            for i in range(nt):
                x = wl.k3_tr[t0 + i] * d
                ix = int(x / wl.xx_step)
                a = (wl.xx[ix + 1] - x) / wl.xx_step
                if ix < xm:
                    want = 0.0
                    skipped += 1
                else:
                    J = a * row[ix - xm] + (1 - a) * row[ix - xm + 1]
                    want = ((J * 1.0) * dtau) * 0.5
                    want /= 4.; want /= (2. * lay.l[il] + 1); want /= 2.
                assert ref[il, t0 + i] == want
    assert skipped > 0          # the skip branch was exercised


def test_time_grid_custom_sampling(oracle, tiny_workload):
    """custom/bessel time sampling (transfer2.c:2420-2521): sorted, keeps every source time,
    steps never exceed tau_step_max, trapezoid weights sum to twice the interval."""
    wl = workload.build_grids(configs.tiny(transfer2_tau_sampling="custom", tau_linstep_song=0.5))
    proj = oracle.OracleProjection([None] * (wl.layout.J_size * wl.layout.L_size * wl.layout.l_size * wl.layout.m_size),
                                   np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1))
    run = oracle.OracleRun(wl, proj)
    L = oracle.lib()
    kmax = float(wl.k[-1])
    nmax = L.oracle_time_grid_size_max(C.byref(run.setup), kmax)
    for k in (float(wl.k[0]), 0.5 * kmax, kmax):
        grid = np.zeros(nmax); d = np.zeros(nmax); w = np.zeros(nmax)
        left = np.zeros(nmax, dtype=np.int32); n = C.c_int(0)
        assert L.oracle_get_time_grid(C.byref(run.setup), k, _d(grid), _i(left), _d(d), _d(w), C.byref(n)) == 0
        g = grid[:n.value]
        assert n.value >= wl.tau.size and np.all(np.diff(g) > 0)
        assert g[0] == wl.tau[0] and g[-1] == wl.tau[-1]
        assert np.all(np.isin(wl.tau, g))
        assert np.diff(g).max() <= 2 * math.pi / k * 0.5 * (1 + 1e-12)
        assert abs(w[:n.value].sum() - 2 * (g[-1] - g[0])) < 1e-9
        assert np.all(left[:n.value] <= wl.tau.size - 2)


@pytest.mark.parametrize("interp", ["linear", "cubic"])
def test_time_interpolation_reproduces_linear_sources(oracle, interp):
    """transfer2_interpolate_sources_in_time (transfer2.c:2816-2880): a source row that is linear in tau is resampled
    exactly on the custom grid, by the linear formula and by the cubic spline (second derivatives vanish)."""
    wl = workload.build_grids(configs.tiny(transfer2_tau_sampling="custom", tau_linstep_song=0.5,
                                           sources_time_interpolation=interp))
    proj = oracle.OracleProjection([None] * (wl.layout.J_size * wl.layout.L_size * wl.layout.l_size * wl.layout.m_size),
                                   np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1))
    run = oracle.OracleRun(wl, proj)
    L = oracle.lib()
    k = 1.0                                                           # short wavelength: steps of 2 pi/k * 0.5 = 3.1 Mpc between the source times
    nmax = L.oracle_time_grid_size_max(C.byref(run.setup), k)
    grid = np.zeros(nmax); d = np.zeros(nmax); w = np.zeros(nmax); left = np.zeros(nmax, dtype=np.int32); n = C.c_int(0)
    assert L.oracle_get_time_grid(C.byref(run.setup), k, _d(grid), _i(left), _d(d), _d(w), C.byref(n)) == 0
    row = np.ascontiguousarray(3.0 - 0.01 * wl.tau)                  # interp_k row of index_k = 0
    out = np.zeros(nmax); spline = np.zeros(wl.tau.size)
    L.oracle_interpolate_sources_in_time.argtypes = [C.POINTER(oracle.Setup), C.c_int, oracle.dp, oracle.dp,
                                                     C.POINTER(C.c_int), C.c_int, oracle.dp, oracle.dp]
    assert L.oracle_interpolate_sources_in_time(C.byref(run.setup), 0, _d(row), _d(grid), _i(left), n.value,
                                                _d(spline), _d(out)) == 0
    want = 3.0 - 0.01 * grid[:n.value]
    assert np.max(np.abs(out[:n.value] - want)) < 1e-12
    assert n.value > wl.tau.size                                        # extra points were really inserted
