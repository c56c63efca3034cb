"""ctypes driver of tests/harness/dropin_harness.c: runs SONG's transfer2_init() -- the drop-in of
song_b200/host/transfer2_gpu.c -- on a synthetic workload, the way main/song.c:106 does."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from song_b200 import _lib

PATH = os.path.join(os.path.dirname(_lib.LOS_PATH), "libsong_transfer2.so")
DROPIN_SYMBOLS = [  # called from outside the module in the reference + the rest of transfer2.h:327-541
    "transfer2_init", "transfer2_free", "transfer2_get_k3_list", "transfer2_load_transfers_from_disk",
    "transfer2_free_type_level", "transfer2_indices_of_transfers", "transfer2_get_l_list", "transfer2_get_lm_lists",
    "transfer2_get_k3_sizes", "transfer2_get_k3_size", "transfer2_interpolate_sources_in_k", "transfer2_get_time_grid",
    "transfer2_interpolate_sources_in_time", "transfer2_compute", "transfer2_integrate",
    "transfer2_store_transfers_to_disk", "transfer2_allocate_type_level", "transfer2_allocate_k1_level",
    "transfer2_free_k1_level"]

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
    ]


_dll = None


def lib():
    global _dll
    if _dll is None:
        if not os.path.exists(PATH):
            raise _lib.SongLosError(f"{PATH} is missing: run __graft_entry__.build() where the SONG headers are available")
        _lib.host()
        _dll = C.CDLL(PATH, mode=C.RTLD_GLOBAL)
        _dll.harness_transfer2.argtypes = [C.POINTER(HarnessInput), dp, C.c_longlong, lp, ip, C.c_char_p, C.c_int]
    return _dll


def run_transfer2(wl, proj, store_dir: str | None = None, verbose: int = 0):
    """transfer2_init on workload `wl` with projection tables `proj` (rows in bessels2 layout).
    Returns (transfer [tt2_size][n_units], count_memorised_transfers, labels_ok)."""
    lay, p = wl.layout, wl.params
    assert wl.pairs.shape[0] == wl.k.size * (wl.k.size + 1) // 2, "the harness needs every (k1,k2) pair"
    keep = dict(k=np.ascontiguousarray(wl.k), tau=np.ascontiguousarray(wl.tau), l=np.ascontiguousarray(lay.l),
                m=np.ascontiguousarray(lay.m), xx=np.ascontiguousarray(wl.xx),
                pt_off=wl.pt_off.astype(np.longlong), src_off=wl.src_off.astype(np.longlong),
                xmin=np.ascontiguousarray(proj.index_xmin, dtype=np.int32), xsize=np.ascontiguousarray(proj.x_size, dtype=np.int32))
    rows = (dp * len(proj.rows))()
    for i, r in enumerate(proj.rows):
        if r is not None:
            rows[i] = _lib.dptr(r)
    ddrows = None
    if getattr(proj, "ddrows", None) is not None:
        ddrows = (dp * len(proj.ddrows))()
        for i, r in enumerate(proj.ddrows):
            if r is not None:
                ddrows[i] = _lib.dptr(r)
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
        J_rows=rows, ddJ_rows=ddrows, index_xmin_J=_lib.iptr(keep["xmin"]), x_size_J=_lib.iptr(keep["xsize"]),
        k3_pt_size=_lib.iptr(wl.k3_pt_size), pt_off=keep["pt_off"].ctypes.data_as(lp), k3_pt=_lib.dptr(wl.k3_pt),
        src_off=keep["src_off"].ctypes.data_as(lp), src=_lib.dptr(wl.src),
        store_to_disk=int(store_dir is not None), load_from_disk=0,
        transfers_dir=(store_dir.encode() if store_dir else None), verbose=verbose,
        tau_sampling_mode={"sources": 0, "custom": 1, "smart": 1, "bessel": 2}[p["transfer2_tau_sampling"]],
        cubic_time=int(p["sources_time_interpolation"] == "cubic"), tau_linstep_song=float(p.get("tau_linstep_song", 0.0)))
    out = np.zeros((lay.tt2_size, wl.n_units), dtype=np.float64)
    sizes = (C.c_longlong * 3)()
    labels_ok = C.c_int(0)
    err = C.create_string_buffer(4096)
    rc = lib().harness_transfer2(C.byref(inp), _lib.dptr(out), out.size, sizes, C.byref(labels_ok), err, 4096)
    if rc:
        raise _lib.SongLosError(err.value.decode(errors="replace"))
    assert sizes[0] == lay.tt2_size and sizes[1] == wl.n_units, (list(sizes), lay.tt2_size, wl.n_units)
    return out, int(sizes[2]), bool(labels_ok.value)
