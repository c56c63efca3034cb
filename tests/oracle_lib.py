"""ctypes wrapper of the CPU oracle (oracle/liblos_oracle.so).  Test infrastructure only:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liblos_oracle.so")

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
lp = C.POINTER(C.c_long)
dpp = C.POINTER(dp)

INTERP = {"linear": 0, "cubic": 1}
TAU_SAMPLING = {"sources": 0, "custom": 1, "smart": 1, "bessel": 2}
K3_SAMPLING = {"bessel": 0, "smart": 1, "class": 1}
K3_EXTRAPOLATION = {"no_extrapolation": 0, "flat_extrapolation": 1}
J_TT, J_EE, J_EB = 0, 1, 2


class Setup(C.Structure):
    _fields_ = [
        ("tau0", C.c_double), ("tau_rec", C.c_double), ("rs_rec", C.c_double),
        ("k_size", C.c_int), ("k", dp),
        ("tau_size", C.c_int), ("tau_sampling", dp),
        ("k3_extrapolation", C.c_int),
        ("extra_k3_oscillations_left", C.c_double), ("extra_k3_oscillations_right", C.c_double),
        ("k_sampling", C.c_int), ("q_linstep_song", C.c_double), ("old_run_step", C.c_int),
        ("smallest_allowed_variation", C.c_double),
        ("tau_sampling_mode", C.c_int), ("tau_linstep_song", C.c_double),
        ("sources_k3_interpolation", C.c_int), ("sources_time_interpolation", C.c_int),
        ("bessels_interpolation", C.c_int),
        ("xx_size", C.c_int), ("xx", dp), ("xx_step", C.c_double), ("xx_max", C.c_double),
        ("J_size", C.c_int), ("index_J_TT", C.c_int), ("index_J_EE", C.c_int), ("index_J_EB", C.c_int),
        ("L_size", C.c_int),
        ("l_size", C.c_int), ("l", ip),
        ("m_size", C.c_int), ("m", ip),
        ("J_rows", dpp), ("ddJ_rows", dpp), ("index_xmin_J", ip), ("x_size_J", ip),
        ("has_source_T", C.c_int), ("has_source_E", C.c_int), ("has_source_B", C.c_int),
        ("index_tp2_T", C.c_int), ("index_tp2_E", C.c_int), ("index_tp2_B", C.c_int),
        ("tp2_size", C.c_int), ("l_max_los_t", C.c_int), ("l_max_los_p", C.c_int),
        ("lm_src", ip),
        ("n_transfers", C.c_int), ("tt2_size", C.c_int),
        ("index_tt2_T", C.c_int), ("index_tt2_E", C.c_int), ("index_tt2_B", C.c_int),
        ("tt_index_l", ip), ("tt_index_m", ip),
    ]


class K3Info(C.Structure):
    _fields_ = [("k_min", C.c_double), ("k_max", C.c_double), ("k_size", C.c_int),
                ("k_physical_start", C.c_int), ("k_physical_size", C.c_int)]


class JL1(C.Structure):
    _fields_ = [("l1_size", C.c_int), ("l1", ip), ("index_l1", ip), ("index_l1_len", C.c_int),
                ("j_l1", dpp), ("index_xmin_l1", ip), ("x_size_l1", ip), ("x_min_l1", dp)]


_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.oracle_last_error.restype = C.c_char_p
        L.oracle_spline_table_columns.argtypes = [dp, C.c_int, dp, C.c_int, dp, C.c_int]
        L.oracle_drc3jj.argtypes = [C.c_double] * 4 + [dp, dp, dp, C.c_int]
        L.oracle_sph_bessel_j.argtypes = [C.c_int, C.c_double]
        L.oracle_sph_bessel_j.restype = C.c_double
        L.oracle_bessel2_l1_list.argtypes = [ip, C.c_int, C.c_int, ip]
        L.oracle_bessel2_j_for_all_l1.argtypes = [ip, C.c_int, C.c_int, C.c_int, dp, C.c_int, C.c_double,
                                                  C.c_double, C.c_double, C.c_double, C.POINTER(JL1)]
        L.oracle_jl1_free.argtypes = [C.POINTER(JL1)]
        L.oracle_bessel2_J_for_Llm.argtypes = [C.c_int] * 5 + [C.POINTER(JL1), dp, C.c_int, C.c_double, C.c_double,
                                                              dpp, ip, ip, dp]
        L.oracle_get_k3_size.argtypes = [C.POINTER(Setup), dp, C.c_int, C.POINTER(K3Info)]
        L.oracle_get_k3_list.argtypes = [C.POINTER(Setup), dp, C.c_int, C.POINTER(K3Info), dp, ip]
        L.oracle_get_time_grid.argtypes = [C.POINTER(Setup), C.c_double, dp, ip, dp, dp, ip]
        L.oracle_time_grid_size_max.argtypes = [C.POINTER(Setup), C.c_double]
        L.oracle_transfer2_run.argtypes = [C.POINTER(Setup), C.c_int, ip, lp, dp, lp, dp, C.POINTER(K3Info),
                                           lp, dp, C.c_long, dp, lp]
        L.oracle_size_indexl_indexm.argtypes = [ip, C.c_int, ip, C.c_int]
        L.oracle_multipole2offset_indexl_indexm.argtypes = [C.c_int, C.c_int, ip, C.c_int, ip, C.c_int]
        L.oracle_multipole2offset_l_indexm.argtypes = [C.c_int, C.c_int, ip, C.c_int]
        L.oracle_offset2multipole_indexl_indexm.argtypes = [C.c_int, ip, C.c_int, ip, C.c_int, ip, ip]
        L.oracle_bessel2_convolution.argtypes = [dp, dp, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_double, dp]
        L.oracle_integral_over_k3.argtypes = [C.c_int, ip, lp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, dp,
                                              C.c_int, C.c_double, dp]
        L.oracle_spectra2_cl.argtypes = [C.c_int, ip, ip, C.c_int, dp, dp, ip, lp, dp, dp, dp, C.c_int, C.c_int]
        L.oracle_spectra2_cl.restype = C.c_double
        L.oracle_free = C.CDLL(None).free
        L.oracle_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


def _l(a):
    return a.ctypes.data_as(lp)


# ----------------------------------------------------------------------------------------------
# projection tables through the oracle's bessel2 restatement
# ----------------------------------------------------------------------------------------------
class OracleProjection:
    """J tables as bessel2_init would leave them in struct bessels2 (ragged)."""

    def __init__(self, rows, index_xmin, x_size, x_min, jl1=None):
        self.rows = rows
        self.index_xmin = index_xmin
        self.x_size = x_size
        self.x_min = x_min
        self.jl1 = jl1   # dict with the j_l1 table (numpy copies)
        self.ddrows = None


def build_projection(wl, with_jl1=False, spline=False) -> OracleProjection:
    """bessel2_init (bessel2.c:75-468) for the layout of workload `wl`."""
    L = lib()
    lay, p = wl.layout, wl.params
    l = np.ascontiguousarray(lay.l, dtype=np.int32)
    m_max = int(lay.m.max())
    L_max = lay.L_size - 1
    jl = JL1()
    rc = L.oracle_bessel2_j_for_all_l1(_i(l), l.size, L_max, m_max, _d(wl.xx), wl.xx.size, wl.xx_max,
                                       p["bessel_j_cut_song"], p["bessel_tol_x_min"],
                                       p["smallest_allowed_variation"], C.byref(jl))
    if rc:
        raise RuntimeError("oracle_bessel2_j_for_all_l1 failed")
    n = lay.J_size * lay.L_size * lay.l_size * lay.m_size
    rows = [None] * n
    xmin = np.full(n, wl.xx.size - 1, dtype=np.int32)
    xsize = np.zeros(n, dtype=np.int32)
    xminv = np.zeros(n, dtype=np.float64)
    types = {}
    if lay.index_J_TT >= 0:
        types[lay.index_J_TT] = J_TT
    if lay.index_J_EE >= 0:
        types[lay.index_J_EE] = J_EE
    if lay.index_J_EB >= 0:
        types[lay.index_J_EB] = J_EB
    for J, pf in types.items():
        for Lv in range(lay.L_size):
            for il in range(lay.l_size):
                for im in range(lay.m_size):
                    mv = int(lay.m[im])
                    if mv > min(Lv, int(l[il])):
                        continue                     # not allocated by the reference (bessel2.c:288)
                    jidx = ((J * lay.L_size + Lv) * lay.l_size + il) * lay.m_size + im
                    row = dp()
                    a, b, c = C.c_int(0), C.c_int(0), C.c_double(0)
                    rc = L.oracle_bessel2_J_for_Llm(pf, Lv, int(l[il]), mv, L_max, C.byref(jl), _d(wl.xx), wl.xx.size,
                                                    wl.xx_max, p["bessel_J_cut_song"], C.byref(row),
                                                    C.byref(a), C.byref(b), C.byref(c))
                    if rc:
                        raise RuntimeError(f"oracle_bessel2_J_for_Llm failed at J={J} L={Lv} l={l[il]} m={mv}")
                    rows[jidx] = np.ctypeslib.as_array(row, shape=(b.value,)).copy()
                    L.oracle_free(row)
                    xmin[jidx], xsize[jidx], xminv[jidx] = a.value, b.value, c.value
    jl1 = None
    if with_jl1:
        n1 = jl.l1_size
        jl1 = {
            "l1": np.ctypeslib.as_array(jl.l1, shape=(n1,)).copy(),
            "index_l1": np.ctypeslib.as_array(jl.index_l1, shape=(jl.index_l1_len,)).copy(),
            "index_xmin_l1": np.ctypeslib.as_array(jl.index_xmin_l1, shape=(n1,)).copy(),
            "x_size_l1": np.ctypeslib.as_array(jl.x_size_l1, shape=(n1,)).copy(),
            "x_min_l1": np.ctypeslib.as_array(jl.x_min_l1, shape=(n1,)).copy(),
        }
        jl1["j_l1"] = [np.ctypeslib.as_array(jl.j_l1[i], shape=(int(jl1["x_size_l1"][i]),)).copy() for i in range(n1)]
    L.oracle_jl1_free(C.byref(jl))
    proj = OracleProjection(rows, xmin, xsize, xminv, jl1)
    if spline:
        proj.ddrows = spline_projection(wl, proj)
    return proj


def spline_projection(wl, proj):
    """ddJ_Llm_x as bessel2.c:395-435 (EST_DERIV for l=2, NATURAL otherwise)."""
    L = lib()
    lay = wl.layout
    dd = [None] * len(proj.rows)
    for jidx, row in enumerate(proj.rows):
        if row is None:
            continue
        il = (jidx // lay.m_size) % lay.l_size
        mode = 1 if int(lay.l[il]) == 2 else 0
        out = np.zeros_like(row)
        if row.size >= 3:
            x = np.ascontiguousarray(wl.xx[proj.index_xmin[jidx]:proj.index_xmin[jidx] + row.size])
            L.oracle_spline_table_columns(_d(x), row.size, _d(row), 1, _d(out), mode)
        dd[jidx] = out
    return dd


# ----------------------------------------------------------------------------------------------
# transfer2
# ----------------------------------------------------------------------------------------------
class OracleRun:
    def __init__(self, wl, proj):
        lay, p = wl.layout, wl.params
        self.wl = wl
        n = len(proj.rows)
        self._rows = (dp * n)()
        self._ddrows = (dp * n)()
        for i, r in enumerate(proj.rows):
            if r is not None:
                self._rows[i] = _d(r)
            if proj.ddrows is not None and proj.ddrows[i] is not None:
                self._ddrows[i] = _d(proj.ddrows[i])
        self._keep = [proj, np.ascontiguousarray(lay.l), np.ascontiguousarray(lay.m), np.ascontiguousarray(lay.lm_src),
                      np.ascontiguousarray(lay.tt_index_l), np.ascontiguousarray(lay.tt_index_m),
                      np.ascontiguousarray(proj.index_xmin), np.ascontiguousarray(proj.x_size)]
        _, l, m, lm_src, tl, tm, xmin, xsize = self._keep
        self.setup = Setup(
            tau0=wl.cosmo["tau0"], tau_rec=wl.cosmo["tau_rec"], rs_rec=wl.cosmo["rs_rec"],
            k_size=wl.k.size, k=_d(wl.k), tau_size=wl.tau.size, tau_sampling=_d(wl.tau),
            k3_extrapolation=K3_EXTRAPOLATION[p["bispectra_k3_extrapolation"]],
            extra_k3_oscillations_left=float(p["extra_k3_oscillations_left"]),
            extra_k3_oscillations_right=float(p["extra_k3_oscillations_right"]),
            k_sampling=K3_SAMPLING[p["transfer2_k3_sampling"]], q_linstep_song=float(p["q_linstep_song"]),
            old_run_step=0, smallest_allowed_variation=float(p["smallest_allowed_variation"]),
            tau_sampling_mode=TAU_SAMPLING[p["transfer2_tau_sampling"]],
            tau_linstep_song=float(p.get("tau_linstep_song", 0.0)),
            sources_k3_interpolation=INTERP[p["sources_k3_interpolation"]],
            sources_time_interpolation=INTERP[p["sources_time_interpolation"]],
            bessels_interpolation=INTERP[p["bessels_interpolation"]],
            xx_size=wl.xx.size, xx=_d(wl.xx), xx_step=wl.xx_step, xx_max=wl.xx_max,
            J_size=lay.J_size, index_J_TT=lay.index_J_TT, index_J_EE=lay.index_J_EE, index_J_EB=lay.index_J_EB,
            L_size=lay.L_size, l_size=lay.l_size, l=_i(l), m_size=lay.m_size, m=_i(m),
            J_rows=self._rows, ddJ_rows=self._ddrows, index_xmin_J=_i(xmin), x_size_J=_i(xsize),
            has_source_T=int(lay.has_source_T), has_source_E=int(lay.has_source_E), has_source_B=int(lay.has_source_B),
            index_tp2_T=lay.index_tp2_T, index_tp2_E=lay.index_tp2_E, index_tp2_B=lay.index_tp2_B,
            tp2_size=lay.tp2_size, l_max_los_t=lay.l_max_los_t, l_max_los_p=lay.l_max_los_p, lm_src=_i(lm_src),
            n_transfers=lay.n_transfers, tt2_size=lay.tt2_size,
            index_tt2_T=lay.index_tt2_T, index_tt2_E=lay.index_tt2_E, index_tt2_B=lay.index_tt2_B,
            tt_index_l=_i(tl), tt_index_m=_i(tm))

    def k3_grids(self):
        """transfer2_get_k3_sizes + _list through the oracle; returns (infos, packed k3_tr)."""
        L, wl = lib(), self.wl
        infos = (K3Info * wl.n_pairs)()
        for i in range(wl.n_pairs):
            o, n = int(wl.pt_off[i]), int(wl.k3_pt_size[i])
            seg = wl.k3_pt[o:o + n]
            if L.oracle_get_k3_size(C.byref(self.setup), _d(seg), n, C.byref(infos[i])):
                raise RuntimeError(f"oracle_get_k3_size failed for pair {i}")
        sizes = np.array([x.k_size for x in infos], dtype=np.int64)
        offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        out = np.zeros(int(sizes.sum()))
        for i in range(wl.n_pairs):
            o, n = int(wl.pt_off[i]), int(wl.k3_pt_size[i])
            seg = wl.k3_pt[o:o + n]
            dst = out[offs[i]:offs[i] + sizes[i]]
            if L.oracle_get_k3_list(C.byref(self.setup), _d(seg), n, C.byref(infos[i]), _d(dst), None):
                raise RuntimeError(f"oracle_get_k3_list failed for pair {i}")
        return infos, out

    def run(self):
        """transfer2_init's compute phase for every pair of the workload -> ([tt][unit], n_terms)."""
        L, wl = lib(), self.wl
        infos = (K3Info * wl.n_pairs)()
        for i in range(wl.n_pairs):
            infos[i] = K3Info(float(wl.k_min_tr[i]), float(wl.k_max_tr[i]), int(wl.k3_tr_size[i]),
                              int(wl.phys_start[i]), int(wl.phys_size[i]))
        out = np.zeros((wl.layout.tt2_size, wl.n_units))
        nterms = C.c_long(0)
        pt_off = wl.pt_off.astype(np.int64); tr_off = wl.tr_off.astype(np.int64); src_off = wl.src_off.astype(np.int64)
        rc = L.oracle_transfer2_run(C.byref(self.setup), wl.n_pairs, _i(wl.k3_pt_size), _l(pt_off), _d(wl.k3_pt),
                                    _l(src_off), _d(wl.src), infos, _l(tr_off), _d(wl.k3_tr), wl.n_units,
                                    _d(out), C.byref(nterms))
        if rc:
            raise RuntimeError("oracle_transfer2_run failed: " + L.oracle_last_error().decode())
        return out, int(nterms.value)


def bessel2_convolution(kk, delta_kk, f, g, j_row, ddj_row, x_min_l1, xx_step, r, cubic=False) -> float:
    """bessel2_convolution (bessel2.c:499-605) on one row of the j_l1 table."""
    kk = np.ascontiguousarray(kk, dtype=np.float64); delta_kk = np.ascontiguousarray(delta_kk, dtype=np.float64)
    f = np.ascontiguousarray(f, dtype=np.float64); j_row = np.ascontiguousarray(j_row, dtype=np.float64)
    g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
    ddj_row = None if ddj_row is None else np.ascontiguousarray(ddj_row, dtype=np.float64)
    out = C.c_double(0)
    lib().oracle_bessel2_convolution(_d(kk), _d(delta_kk), int(kk.size), _d(f), None if g is None else _d(g), _d(j_row),
                                     None if ddj_row is None else _d(ddj_row), float(x_min_l1), float(xx_step), int(cubic),
                                     float(r), C.byref(out))
    return float(out.value)


def integral_over_k3(wl, transfer_row, j_row, ddj_row, x_min_l1, r, abs_M3=0, M3_coefficient=1.0, cubic=False) -> np.ndarray:
    """bispectra2_intrinsic_integrate_over_k3 (bispectra2.c:1107-1439) for one (l3, M3): [r_size][n_pairs]."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    out = np.zeros((r.size, wl.n_pairs))
    tr_off = wl.tr_off.astype(np.int64)
    t = np.ascontiguousarray(transfer_row, dtype=np.float64)
    j_row = np.ascontiguousarray(j_row, dtype=np.float64)
    ddj_row = None if ddj_row is None else np.ascontiguousarray(ddj_row, dtype=np.float64)
    rc = lib().oracle_integral_over_k3(wl.n_pairs, _i(wl.k3_tr_size), _l(tr_off), _d(wl.k3_tr), _d(t), _d(j_row),
                                       None if ddj_row is None else _d(ddj_row), float(x_min_l1), float(wl.xx_step), int(cubic),
                                       int(r.size), _d(r), int(abs_M3), float(M3_coefficient), _d(out))
    if rc:
        raise RuntimeError("oracle_integral_over_k3: an integration grid has less than two elements")
    return out


def spectra2_cl(wl, transfer_1, transfer_2, m, primordial, rescale_cmb_sources=True) -> float:
    """spectra2_cls (spectra2.c:241-570) for one (index_ct, index_l, index_M) over the pairs of `wl`."""
    k1 = np.ascontiguousarray(wl.pairs[:, 0], dtype=np.int32); k2 = np.ascontiguousarray(wl.pairs[:, 1], dtype=np.int32)
    tr_off = wl.tr_off.astype(np.int64)
    t1 = np.ascontiguousarray(transfer_1, dtype=np.float64); t2 = np.ascontiguousarray(transfer_2, dtype=np.float64)
    prim = np.ascontiguousarray(primordial, dtype=np.float64); k = np.ascontiguousarray(wl.k, dtype=np.float64)
    return float(lib().oracle_spectra2_cl(wl.n_pairs, _i(k1), _i(k2), int(k.size), _d(k), _d(prim), _i(wl.k3_tr_size), _l(tr_off),
                                          _d(wl.k3_tr), _d(t1), _d(t2), int(m), int(rescale_cmb_sources)))
