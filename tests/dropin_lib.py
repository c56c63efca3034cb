"""ctypes driver of tests/harness/dropin_harness.c: runs SONG's transfer2_init() -- the drop-in of
song_b200/host/transfer2_gpu.c -- on a synthetic workload, the way main/song.c:106 does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from song_b200 import _lib

PATH = os.path.join(os.path.dirname(_lib.LOS_PATH), "libsong_transfer2.so")
# the REFERENCE's own transfer2.c + bessel2.c behind the same harness (oracle/Makefile, target `ref`; test infrastructure)
REF_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsong_ref.so")
DROPIN_SYMBOLS = [  # called from outside the module in the reference + the rest of transfer2.h:327-541
    "transfer2_init", "transfer2_free", "transfer2_get_k3_list", "transfer2_load_transfers_from_disk",
    "transfer2_free_type_level", "transfer2_indices_of_transfers", "transfer2_get_l_list", "transfer2_get_lm_lists",
    "transfer2_get_k3_sizes", "transfer2_get_k3_size", "transfer2_interpolate_sources_in_k", "transfer2_get_time_grid",
    "transfer2_interpolate_sources_in_time", "transfer2_compute", "transfer2_integrate",
    "transfer2_store_transfers_to_disk", "transfer2_allocate_type_level", "transfer2_allocate_k1_level",
    "transfer2_free_k1_level",
    # bessel2 module (include/bessel2.h:161-260); bessel2_J_Llm_at_x is declared there but defined nowhere in the reference
    "bessel2_init", "bessel2_free", "bessel2_convolution", "bessel2_get_l1_list", "bessel2_get_xx_list", "bessel2_j_for_l1",
    "bessel2_J_for_Llm", "bessel2_J_Llm", "bessel2_l1_at_x", "bessel2_l1_at_x_linear"]

dp, ip, lp = _lib.c_double_p, _lib.c_int_p, C.POINTER(C.c_longlong)


class HarnessInput(C.Structure):
    _fields_ = [
        ("tau0", C.c_double), ("tau_rec", C.c_double), ("rs_rec", C.c_double),
        ("k_size", C.c_int), ("k", dp),
        ("tau_size", C.c_int), ("tau_sampling", dp),
        ("l_size", C.c_int), ("l", ip),
        ("m_size", C.c_int), ("m", ip),
        ("L_max", C.c_int), ("l_max_los_t", C.c_int), ("l_max_los_p", C.c_int),
        ("has_T", C.c_int), ("has_E", C.c_int), ("has_B", C.c_int),
        ("cubic_k3", C.c_int), ("cubic_bessel", C.c_int), ("k3_extrapolation", C.c_int),
        ("extra_left", C.c_double), ("extra_right", C.c_double), ("q_linstep_song", C.c_double),
        ("smallest_allowed_variation", C.c_double),
        ("xx_size", C.c_int), ("xx", dp), ("xx_step", C.c_double), ("xx_max", C.c_double),
        ("J_size", C.c_int), ("index_J_TT", C.c_int), ("index_J_EE", C.c_int), ("index_J_EB", C.c_int),
        ("J_rows", C.POINTER(dp)), ("ddJ_rows", C.POINTER(dp)), ("index_xmin_J", ip), ("x_size_J", ip),
        ("k3_pt_size", ip), ("pt_off", lp), ("k3_pt", dp),
        ("src_off", lp), ("src", dp),
        ("store_to_disk", C.c_int), ("load_from_disk", C.c_int), ("transfers_dir", C.c_char_p),
        ("verbose", C.c_int),
        ("tau_sampling_mode", C.c_int), ("cubic_time", C.c_int), ("tau_linstep_song", C.c_double),
        ("use_bessel2_init", C.c_int), ("j_l1_cut", C.c_double), ("J_Llm_cut", C.c_double), ("bessel_tol_x_min", C.c_double),
        ("keep_sources", C.c_int),
    ]


_dlls = {}


def lib(path: str | None = None):
    """The harness library: the drop-in (default) or, with path=REF_PATH, the reference's own modules.  The reference
    library is loaded RTLD_LOCAL (and linked -Bsymbolic): both export transfer2_init & co."""
    path = path or PATH
    _dll = _dlls.get(path)
    if _dll is None:
        if not os.path.exists(path):
            raise _lib.SongLosError(f"{path} is missing: run __graft_entry__.build() where the SONG headers are available")
        if path == PATH:
            _lib.host()
        _dll = C.CDLL(path, mode=C.RTLD_GLOBAL if path == PATH else C.RTLD_LOCAL)
        _dlls[path] = _dll
        vp = C.c_void_p
        _dll.harness_transfer2.argtypes = [C.POINTER(HarnessInput), dp, C.c_longlong, lp, ip, C.c_char_p, C.c_int]
        _dll.harness_create.argtypes = [C.POINTER(HarnessInput), C.c_char_p, C.c_int]
        _dll.harness_create.restype = vp
        _dll.harness_fill_sources.argtypes = [vp]
        _dll.harness_transfer2_init.argtypes = [vp, C.c_char_p, C.c_int]
        _dll.harness_checksum.argtypes = [vp]
        _dll.harness_checksum.restype = C.c_double
        _dll.harness_read.argtypes = [vp, dp, C.c_longlong, lp, ip, C.c_char_p, C.c_int]
        _dll.harness_transfer2_free.argtypes = [vp, C.c_char_p, C.c_int]
        _dll.harness_bessel2_tables.argtypes = [vp, ip, ip, C.POINTER(dp), C.POINTER(dp), lp]
        _dll.harness_bessel2_convolution.argtypes = [vp, dp, dp, C.c_int, dp, dp, C.c_int, C.c_double, dp]
        _dll.harness_l1_info.argtypes = [vp, ip, dp]
        _dll.harness_j_l1_row.argtypes = [vp, C.c_int, dp, C.c_int, ip]
        _dll.harness_k3_list.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int]
        _dll.harness_indices_digest.argtypes = [vp, C.c_char_p, C.c_longlong]
        _dll.harness_indices_digest.restype = C.c_longlong
        _dll.harness_destroy.argtypes = [vp]
        if hasattr(_dll, "ref_integrate_over_k3"):       # oracle/_ref/libsong_ref_bisp.so only (the reference's bispectra2.c)
            _dll.ref_integrate_over_k3.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, dp, C.c_int, dp, C.c_char_p, C.c_int]
            _dll.ref_spectra2_cls.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_char_p, C.c_int]
    return _dll


def _input(wl, proj, store_dir=None, verbose=0, use_bessel2_init=False, keep_sources=False, load_dir=None):
    """harness_input for workload `wl`; returns (struct, objects that must stay alive)."""
    lay, p = wl.layout, wl.params
    assert wl.pairs.shape[0] == wl.k.size * (wl.k.size + 1) // 2, "the harness needs every (k1,k2) pair"
    keep = dict(k=np.ascontiguousarray(wl.k), tau=np.ascontiguousarray(wl.tau), l=np.ascontiguousarray(lay.l),
                m=np.ascontiguousarray(lay.m), xx=np.ascontiguousarray(wl.xx),
                pt_off=wl.pt_off.astype(np.longlong), src_off=wl.src_off.astype(np.longlong), wl=wl, proj=proj)
    rows = ddrows = None
    xmin_p = xsize_p = None
    if proj is not None:
        keep["xmin"] = np.ascontiguousarray(proj.index_xmin, dtype=np.int32)
        keep["xsize"] = np.ascontiguousarray(proj.x_size, dtype=np.int32)
        xmin_p, xsize_p = _lib.iptr(keep["xmin"]), _lib.iptr(keep["xsize"])
        rows = (dp * len(proj.rows))()
        for i, r in enumerate(proj.rows):
            if r is not None:
                rows[i] = _lib.dptr(r)
        if getattr(proj, "ddrows", None) is not None:
            ddrows = (dp * len(proj.ddrows))()
            for i, r in enumerate(proj.ddrows):
                if r is not None:
                    ddrows[i] = _lib.dptr(r)
    else:
        assert use_bessel2_init, "without projection tables the harness must call bessel2_init"
    keep["rows"], keep["ddrows"] = rows, ddrows
    assert not (store_dir and load_dir)
    keep["dir"] = store_dir.encode() if store_dir else (load_dir.encode() if load_dir else None)
    inp = HarnessInput(
        tau0=wl.cosmo["tau0"], tau_rec=wl.cosmo["tau_rec"], rs_rec=wl.cosmo["rs_rec"],
        k_size=int(wl.k.size), k=_lib.dptr(keep["k"]), tau_size=int(wl.tau.size), tau_sampling=_lib.dptr(keep["tau"]),
        l_size=lay.l_size, l=_lib.iptr(keep["l"]), m_size=lay.m_size, m=_lib.iptr(keep["m"]),
        L_max=lay.L_size - 1, l_max_los_t=lay.l_max_los_t, l_max_los_p=lay.l_max_los_p,
        has_T=int(lay.has_source_T), has_E=int(lay.has_source_E), has_B=int(lay.has_source_B),
        cubic_k3=int(p["sources_k3_interpolation"] == "cubic"), cubic_bessel=int(p["bessels_interpolation"] == "cubic"),
        k3_extrapolation=int(p["bispectra_k3_extrapolation"] != "no_extrapolation"),
        extra_left=float(p["extra_k3_oscillations_left"]), extra_right=float(p["extra_k3_oscillations_right"]),
        q_linstep_song=float(p["q_linstep_song"]), smallest_allowed_variation=float(p["smallest_allowed_variation"]),
        xx_size=int(wl.xx.size), xx=_lib.dptr(keep["xx"]), xx_step=float(wl.xx_step), xx_max=float(wl.xx_max),
        J_size=lay.J_size, index_J_TT=lay.index_J_TT, index_J_EE=lay.index_J_EE, index_J_EB=lay.index_J_EB,
        J_rows=rows, ddJ_rows=ddrows, index_xmin_J=xmin_p, x_size_J=xsize_p,
        k3_pt_size=_lib.iptr(wl.k3_pt_size), pt_off=keep["pt_off"].ctypes.data_as(lp), k3_pt=_lib.dptr(wl.k3_pt),
        src_off=keep["src_off"].ctypes.data_as(lp), src=_lib.dptr(wl.src),
        store_to_disk=int(store_dir is not None), load_from_disk=int(load_dir is not None),
        transfers_dir=keep["dir"], verbose=verbose,
        tau_sampling_mode={"sources": 0, "custom": 1, "smart": 1, "bessel": 2}[p["transfer2_tau_sampling"]],
        cubic_time=int(p["sources_time_interpolation"] == "cubic"), tau_linstep_song=float(p.get("tau_linstep_song", 0.0)),
        use_bessel2_init=int(use_bessel2_init), j_l1_cut=float(p["bessel_j_cut_song"]), J_Llm_cut=float(p["bessel_J_cut_song"]),
        bessel_tol_x_min=float(p["bessel_tol_x_min"]), keep_sources=int(keep_sources))
    return inp, keep


class Harness:
    """The SONG structures of one run, kept alive between calls: main/song.c up to transfer2_init, step by step.

    h = Harness(wl, proj)            # input2, perturbations2, bessel, bessel2 (bessel2_init drop-in when proj is None)
    h.transfer2_init()               # main/song.c:106 -- the call the benchmark times
    T = h.read()                     # ptr2->transfer in transfers_%03d.dat order
    h.transfer2_free(); h.fill_sources(); h.transfer2_init()   # next step
    """

    def __init__(self, wl, proj=None, store_dir=None, verbose=0, keep_sources=False, lib_path=None, load_dir=None):
        self.wl = wl
        self.lib_path = lib_path
        self.inp, self._keep = _input(wl, proj, store_dir, verbose, use_bessel2_init=proj is None, keep_sources=keep_sources,
                                      load_dir=load_dir)
        err = C.create_string_buffer(4096)
        self.h = self._lib().harness_create(C.byref(self.inp), err, 4096)
        if not self.h:
            raise _lib.SongLosError(err.value.decode(errors="replace"))

    def _lib(self):
        return lib(self.lib_path)

    def _check(self, rc, err):
        if rc:
            raise _lib.SongLosError(err.value.decode(errors="replace"))

    def fill_sources(self):
        self._lib().harness_fill_sources(self.h)

    def transfer2_init(self):
        err = C.create_string_buffer(4096)
        self._check(self._lib().harness_transfer2_init(self.h, err, 4096), err)

    def checksum(self) -> float:
        return float(self._lib().harness_checksum(self.h))

    def read(self):
        lay, wl = self.wl.layout, self.wl
        out = np.zeros((lay.tt2_size, wl.n_units), dtype=np.float64)
        sizes = (C.c_longlong * 3)()
        labels_ok = C.c_int(0)
        err = C.create_string_buffer(4096)
        self._check(self._lib().harness_read(self.h, _lib.dptr(out), out.size, sizes, C.byref(labels_ok), err, 4096), err)
        assert sizes[0] == lay.tt2_size and sizes[1] == wl.n_units, (list(sizes), lay.tt2_size, wl.n_units)
        return out, int(sizes[2]), bool(labels_ok.value)

    def transfer2_free(self):
        err = C.create_string_buffer(4096)
        self._check(self._lib().harness_transfer2_free(self.h, err, 4096), err)

    def bessel2_tables(self):
        """(rows, index_xmin_J, x_size_J, ddrows, info) of struct bessels2, flat jidx order (copies)."""
        lay = self.wl.layout
        n = lay.J_size * lay.L_size * lay.l_size * lay.m_size
        sizes = (C.c_longlong * 6)()
        xmin = np.zeros(n, dtype=np.int32); xsize = np.zeros(n, dtype=np.int32)
        rows = (dp * n)(); ddrows = (dp * n)()
        self._lib().harness_bessel2_tables(self.h, _lib.iptr(xmin), _lib.iptr(xsize), rows, ddrows, sizes)
        out = [np.ctypeslib.as_array(rows[i], shape=(int(xsize[i]),)).copy() if xsize[i] > 0 and rows[i] else None for i in range(n)]
        dd = [np.ctypeslib.as_array(ddrows[i], shape=(int(xsize[i]),)).copy() if xsize[i] > 0 and ddrows[i] else None for i in range(n)]
        info = dict(J_size=sizes[0], L_size=sizes[1], xx_size=sizes[2], l1_size=sizes[3], x_size_max_J=sizes[4], count_allocated_Js=sizes[5])
        return out, xmin, xsize, dd, info

    def l1_info(self):
        sizes = (C.c_longlong * 6)()
        self._lib().harness_bessel2_tables(self.h, None, None, None, None, sizes)
        l1 = np.zeros(int(sizes[3]), dtype=np.int32); xmin = np.zeros(int(sizes[3]))
        self._lib().harness_l1_info(self.h, _lib.iptr(l1), _lib.dptr(xmin))
        return l1, xmin

    def indices_digest(self) -> str:
        """Index tables, labels, grid sizes and physical ranges of struct transfers2 as text (harness_indices_digest)."""
        cap = 64 * 1024 * 1024
        buf = C.create_string_buffer(cap)
        n = self._lib().harness_indices_digest(self.h, buf, cap)
        assert 0 < n < cap
        return buf.value.decode()

    def k3_lists(self) -> np.ndarray:
        """transfer2_get_k3_list of the module for every pair, packed in the reference loop order (after transfer2_init)."""
        out = np.zeros(self.wl.n_units)
        for ip in range(self.wl.n_pairs):
            a, n = int(self.wl.tr_off[ip]), int(self.wl.k3_tr_size[ip])
            got = self._lib().harness_k3_list(self.h, int(self.wl.pairs[ip, 0]), int(self.wl.pairs[ip, 1]), _lib.dptr(out[a:a + n]), n)
            if got != n:
                raise _lib.SongLosError(f"pair {ip}: transfer2_get_k3_list gives {got} values, the workload has {n}")
        return out

    def j_l1_rows(self, xx_size: int):
        """pbs2->j_l1 as bessel2_init left it: (rows, index_xmin_l1)."""
        sizes = (C.c_longlong * 6)()
        self._lib().harness_bessel2_tables(self.h, None, None, None, None, sizes)
        rows, xmin = [], []
        buf = np.zeros(int(xx_size)); ix = C.c_int(0)
        for i in range(int(sizes[3])):
            n = self._lib().harness_j_l1_row(self.h, i, _lib.dptr(buf), int(xx_size), C.byref(ix))
            rows.append(buf[:n].copy()); xmin.append(int(ix.value))
        return rows, np.asarray(xmin, dtype=np.int32)

    def integrate_over_k3(self, index_tt2_k3, index_M3, offset_L3, abs_M3, M3_coefficient, r) -> np.ndarray:
        """The reference's bispectra2_intrinsic_integrate_over_k3 (libsong_ref_bisp.so): [l_size][r_size][n_pairs]."""
        r = np.ascontiguousarray(r, dtype=np.float64)
        out = np.zeros((self.wl.layout.l_size, r.size, self.wl.n_pairs))
        err = C.create_string_buffer(4096)
        self._check(self._lib().ref_integrate_over_k3(self.h, int(index_tt2_k3), int(index_M3), int(offset_L3), int(abs_M3),
                                                      float(M3_coefficient), _lib.dptr(r), int(r.size), _lib.dptr(out), err, 4096), err)
        return out

    def spectra2_cls(self, primordial, has_tt2=True, has_ee2=True, has_te2=True, rescale_cmb_sources=True) -> np.ndarray:
        """The reference's spectra2_cls (libsong_ref_bisp.so): [requested C_l types in the order tt2, ee2, te2][l_size]."""
        prim = np.ascontiguousarray(primordial, dtype=np.float64)
        n_ct = int(has_tt2) + int(has_ee2) + int(has_te2)
        out = np.zeros((n_ct, self.wl.layout.l_size))
        err = C.create_string_buffer(4096)
        self._check(self._lib().ref_spectra2_cls(self.h, _lib.dptr(prim), int(has_tt2), int(has_ee2), int(has_te2),
                                                 int(rescale_cmb_sources), _lib.dptr(out), err, 4096), err)
        return out

    def bessel2_convolution(self, kk, delta_kk, f, g, index_l1, r) -> float:
        kk = np.ascontiguousarray(kk, dtype=np.float64); delta_kk = np.ascontiguousarray(delta_kk, dtype=np.float64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        g = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
        out = C.c_double(0)
        rc = self._lib().harness_bessel2_convolution(self.h, _lib.dptr(kk), _lib.dptr(delta_kk), int(kk.size), _lib.dptr(f),
                                               None if g is None else _lib.dptr(g), int(index_l1), float(r), C.byref(out))
        if rc:
            raise _lib.SongLosError("bessel2_convolution failed")
        return float(out.value)

    def close(self):
        if self.h:
            self._lib().harness_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def run_transfer2(wl, proj, store_dir: str | None = None, verbose: int = 0):
    """transfer2_init on workload `wl` with projection tables `proj` (rows in bessels2 layout; None: the bessel2_init
    drop-in builds them).  Returns (transfer [tt2_size][n_units], count_memorised_transfers, labels_ok)."""
    with Harness(wl, proj, store_dir=store_dir, verbose=verbose) as h:
        h.transfer2_init()
        out = h.read()
        h.transfer2_free()
    return out
