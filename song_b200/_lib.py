"""ctypes bindings of the two in-tree shared libraries.

* ``libsong_los.so``  -- CUDA kernels + the C-ABI of ``include/song_los.h`` (nvcc, sm_100a)
* ``libsong_host.so`` -- host C of the transfer2 / bessel2 drop-in (gcc)

There is no fallback of any kind: if a library is missing or no CUDA device is present the
import / the first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LOS_PATH = os.path.join(_HERE, "libsong_los.so")
HOST_PATH = os.path.join(_HERE, "libsong_host.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)


class SongLosError(RuntimeError):
    """A C-ABI call returned SONG_LOS_FAILURE; the message is the library's error text."""


class LosConfig(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("tau0", C.c_double),
        ("tau_size", C.c_int),
        ("tau_sampling", c_double_p),
        ("tau_sampling_mode", C.c_int),
        ("tau_linstep_song", C.c_double),
        ("sources_k3_interpolation", C.c_int),
        ("sources_time_interpolation", C.c_int),
        ("bessels_interpolation", C.c_int),
        ("k3_extrapolation", C.c_int),
        ("xx_size", C.c_int),
        ("xx", c_double_p),
        ("xx_step", C.c_double),
        ("l_size", C.c_int), ("l", c_int_p),
        ("m_size", C.c_int), ("m", c_int_p),
        ("L_size", C.c_int),
        ("n_transfers", C.c_int),
        ("lm_array", c_int_p),
        ("lm_src", c_int_p),
        ("has_source_T", C.c_int), ("has_source_E", C.c_int), ("has_source_B", C.c_int),
        ("index_tp2_T", C.c_int), ("index_tp2_E", C.c_int), ("index_tp2_B", C.c_int),
        ("tp2_size", C.c_int),
        ("l_max_los_t", C.c_int), ("l_max_los_p", C.c_int),
        ("index_tt2_T", C.c_int), ("index_tt2_E", C.c_int), ("index_tt2_B", C.c_int),
        ("tt2_size", C.c_int),
        ("tt_index_l", c_int_p),
        ("tt_index_m", c_int_p),
        ("J_size", C.c_int),
        ("index_J_TT", C.c_int), ("index_J_EE", C.c_int), ("index_J_EB", C.c_int),
    ]


class LosBatch(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_int),
        ("k3_pt_size", c_int_p),
        ("k3_tr_size", c_int_p),
        ("phys_start", c_int_p),
        ("phys_size", c_int_p),
        ("pt_off", c_int64_p),
        ("tr_off", c_int64_p),
        ("src_off", c_int64_p),
        ("n_pt_total", C.c_int64),
        ("n_units", C.c_int64),
        ("n_src_total", C.c_int64),
        ("k3_pt", c_double_p),
        ("k3_tr", c_double_p),
        ("src", c_double_p),
        ("phys_k_first", c_double_p),     # optional: slices of fine grids (k3 bands), see include/song_los.h
        ("phys_k_last", c_double_p),
    ]


class LosTiming(C.Structure):
    _fields_ = [
        ("ms_total", C.c_float),
        ("ms_spline", C.c_float),
        ("ms_sort", C.c_float),
        ("ms_interp", C.c_float),
        ("ms_integrate", C.c_float),
        ("ms_reorder", C.c_float),
        ("n_launches", C.c_int),
        ("n_integrals", C.c_int64),
        ("n_terms_upper", C.c_int64),
        ("ms_h2d", C.c_float),
        ("ms_d2h", C.c_float),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class K3Request(C.Structure):
    """struct song_los_k3_request (include/song_los.h)."""
    _fields_ = [("n_requests", C.c_int), ("tt", c_int_p), ("index_l1", c_int_p), ("abs_M3", C.c_int),
                ("M3_coefficient", c_double_p), ("r_size", C.c_int), ("r", c_double_p)]


class ClRequest(C.Structure):
    """struct song_los_cl_request (include/song_los.h)."""
    _fields_ = [("n_requests", C.c_int), ("tt1", c_int_p), ("tt2", c_int_p), ("m", c_int_p), ("rescale_cmb_sources", C.c_int),
                ("k_size", C.c_int), ("k", c_double_p), ("primordial", c_double_p), ("pair_k1", c_int_p), ("pair_k2", c_int_p)]


class K3Params(C.Structure):
    _fields_ = [
        ("tau0", C.c_double), ("tau_rec", C.c_double), ("rs_rec", C.c_double),
        ("k_first", C.c_double), ("k_last", C.c_double),
        ("k3_extrapolation", C.c_int),
        ("extra_k3_oscillations_left", C.c_double),
        ("extra_k3_oscillations_right", C.c_double),
        ("k_sampling", C.c_int),
        ("q_linstep_song", C.c_double),
        ("old_run_step", C.c_int),
        ("smallest_allowed_variation", C.c_double),
        ("xx_max", C.c_double), ("xx_step", C.c_double),
    ]


class K3Range(C.Structure):
    _fields_ = [
        ("k_min", C.c_double), ("k_max", C.c_double),
        ("k_size", C.c_int),
        ("k_physical_start", C.c_int),
        ("k_physical_size", C.c_int),
    ]


class Bessel2Host(C.Structure):
    """struct song_bessel2_host (song_b200/host/bessel2_host.h)."""
    _fields_ = [
        ("l1_size", C.c_int), ("l1", c_int_p), ("index_l1", c_int_p), ("index_l1_len", C.c_int),
        ("j_l1", C.POINTER(c_double_p)), ("index_xmin_l1", c_int_p), ("x_size_l1", c_int_p), ("x_min_l1", c_double_p),
        ("n_jidx", C.c_int), ("ndim", C.c_int),
        ("first_3j", c_double_p), ("second_3j", c_double_p),
        ("l1_min", c_int_p), ("l1_count", c_int_p), ("index_l1_of_l1min", c_int_p),
        ("count_allocated_Js", C.c_long),
    ]


#: every symbol include/song_los.h declares (checked by tests/test_abi.py without a GPU)
LOS_SYMBOLS = [
    "song_los_create", "song_los_destroy", "song_los_last_error", "song_los_set_projection",
    "song_los_build_projection", "song_los_get_projection", "song_los_upload_pairs",
    "song_los_compute", "song_los_download", "song_los_transfer_pairs",
    "song_los_device_result", "song_los_copy_result", "song_los_set_stream", "song_los_measure_peaks",
    "song_los_device_info", "song_los_version",
    "song_los_submit", "song_los_wait", "song_los_host_alloc", "song_los_host_free", "song_los_max_batch_units", "song_los_max_resident_units", "song_los_stage_sources", "song_los_reserve", "song_los_get_projection_dd",
    "song_los_device_count", "song_los_set_time", "song_los_config_matches", "song_los_drop_projection_rows",
    "song_los_select_slot", "song_los_set_bessel_table", "song_los_integrate_k3", "song_los_last_k3_ms", "song_los_cl_sum",
]
HOST_SYMBOLS = ["song_host_k3_size", "song_host_k3_list", "song_host_k3_sizes", "song_host_k3_lists",
                "song_host_bessel2_prepare", "song_host_bessel2_free", "song_host_bessel2_upload",
                "song_host_lpt_schedule", "song_host_schedule_pairs", "song_host_pair_costs",
                "song_host_band_edges", "song_host_band_cuts", "song_host_band_count", "song_host_pair_costs_model"]

_los = None
_host = None


def los():
    """The CUDA library.  Raises if it has not been built (no silent fallback)."""
    global _los
    if _los is None:
        if not os.path.exists(LOS_PATH):
            raise SongLosError(
                f"{LOS_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        lib = C.CDLL(LOS_PATH, mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        lib.song_los_create.argtypes = [C.POINTER(LosConfig), C.POINTER(vp)]
        lib.song_los_destroy.argtypes = [vp]
        lib.song_los_last_error.argtypes = [vp]
        lib.song_los_last_error.restype = C.c_char_p
        lib.song_los_set_projection.argtypes = [vp, C.POINTER(c_double_p), C.POINTER(c_double_p), c_int_p, c_int_p]
        lib.song_los_build_projection.argtypes = [
            vp, C.c_int, c_int_p, C.POINTER(c_double_p), c_int_p, c_int_p,
            C.c_int, c_double_p, c_double_p, c_int_p, c_int_p, c_int_p, c_int_p, C.c_double, C.c_double]
        lib.song_los_get_projection.argtypes = [vp, c_int_p, c_int_p, C.POINTER(c_double_p)]
        lib.song_los_get_projection_dd.argtypes = [vp, C.POINTER(c_double_p)]
        lib.song_los_upload_pairs.argtypes = [vp, C.POINTER(LosBatch)]
        lib.song_los_compute.argtypes = [vp, C.c_int, C.POINTER(LosTiming)]
        lib.song_los_download.argtypes = [vp, c_double_p]
        lib.song_los_transfer_pairs.argtypes = [vp, C.POINTER(LosBatch), C.c_int, c_double_p, C.POINTER(LosTiming)]
        lib.song_los_device_result.argtypes = [vp, c_int64_p]
        lib.song_los_device_result.restype = C.c_uint64
        lib.song_los_set_stream.argtypes = [vp, vp]
        lib.song_los_measure_peaks.argtypes = [vp, c_double_p, c_double_p, c_double_p]
        lib.song_los_copy_result.argtypes = [vp, C.c_uint64, C.c_int64]
        lib.song_los_device_info.argtypes = [vp, c_int64_p, c_int64_p, c_int_p, C.c_char_p, C.c_int]
        lib.song_los_version.argtypes = []
        lib.song_los_submit.argtypes = [vp, C.POINTER(LosBatch), C.c_int, c_double_p]
        lib.song_los_wait.argtypes = [vp, C.POINTER(LosTiming)]
        lib.song_los_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
        lib.song_los_host_free.argtypes = [vp]
        lib.song_los_max_batch_units.argtypes = [vp, C.c_int, C.c_double, c_int64_p]
        lib.song_los_stage_sources.argtypes = [vp, c_double_p, C.c_int64, C.c_int64, C.c_int64]
        lib.song_los_max_resident_units.argtypes = [vp, C.c_int, C.c_double, C.c_int, c_int64_p]
        lib.song_los_reserve.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64]
        lib.song_los_device_count.argtypes = [c_int_p]
        lib.song_los_set_time.argtypes = [vp, C.c_double, C.c_int, C.c_double]
        lib.song_los_config_matches.argtypes = [vp, C.POINTER(LosConfig)]
        lib.song_los_drop_projection_rows.argtypes = [vp]
        lib.song_los_select_slot.argtypes = [vp, C.c_int]
        lib.song_los_set_bessel_table.argtypes = [vp, C.c_int, C.POINTER(c_double_p), C.POINTER(c_double_p), c_double_p, c_int_p]
        lib.song_los_integrate_k3.argtypes = [vp, C.POINTER(K3Request), c_double_p]
        lib.song_los_cl_sum.argtypes = [vp, C.POINTER(ClRequest), c_double_p]
        lib.song_los_last_k3_ms.argtypes = [vp]
        lib.song_los_last_k3_ms.restype = C.c_float
        _los = lib
    return _los


def host():
    """The host C library (k3 grids, drop-in module)."""
    global _host
    if _host is None:
        if not os.path.exists(HOST_PATH):
            raise SongLosError(f"{HOST_PATH} is missing: run __graft_entry__.build()")
        los()  # libsong_host links libsong_los
        lib = C.CDLL(HOST_PATH, mode=C.RTLD_GLOBAL)
        lib.song_host_k3_size.argtypes = [C.POINTER(K3Params), c_double_p, C.c_int, C.POINTER(K3Range), C.c_char_p, C.c_int]
        lib.song_host_k3_list.argtypes = [C.POINTER(K3Params), c_double_p, C.c_int, C.POINTER(K3Range), c_double_p, c_int_p, C.c_char_p, C.c_int]
        lib.song_host_k3_sizes.argtypes = [C.POINTER(K3Params), C.c_int, c_int_p, c_int64_p, c_double_p,
                                           C.POINTER(K3Range), c_int_p, c_int64_p, C.c_char_p, C.c_int]
        lib.song_host_k3_lists.argtypes = [C.POINTER(K3Params), C.c_int, c_int_p, c_int64_p, c_double_p,
                                           C.POINTER(K3Range), c_int64_p, c_double_p, C.c_char_p, C.c_int]
        lib.song_host_bessel2_prepare.argtypes = [c_int_p, C.c_int, c_int_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_double,
                                                  C.c_int, c_int_p, C.c_double, C.c_double, C.c_double,
                                                  C.POINTER(C.POINTER(Bessel2Host)), C.c_char_p, C.c_int]
        lib.song_host_bessel2_free.argtypes = [C.POINTER(Bessel2Host)]
        lib.song_host_bessel2_upload.argtypes = [C.c_void_p, C.POINTER(Bessel2Host), c_int_p, C.c_double, C.c_double,
                                                 C.c_char_p, C.c_int]
        lib.song_host_lpt_schedule.argtypes = [c_int64_p, C.c_int, C.c_int, c_int_p, c_int64_p]
        lib.song_host_schedule_pairs.argtypes = [c_int64_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int64_p]
        lib.song_host_pair_costs.argtypes = [C.c_int, c_int_p, c_double_p, c_double_p, C.c_double, C.c_int, c_int64_p]
        lib.song_host_band_edges.argtypes = [C.c_int, c_int_p, c_int64_p, c_double_p, c_int_p, C.c_void_p, C.c_double, C.c_int, C.c_double,
                                             C.c_int, c_double_p]
        lib.song_host_pair_costs_model.argtypes = [C.c_int, c_int_p, c_int64_p, c_double_p, c_int_p, C.c_double, C.c_int, C.c_double, C.c_int,
                                                   c_int64_p]
        lib.song_host_band_cuts.argtypes = [C.c_int, c_int_p, c_int64_p, c_double_p, C.c_int, c_double_p, c_int_p]
        lib.song_host_band_count.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int]
        lib.compat_drc3jj.argtypes = [C.c_double] * 4 + [c_double_p, c_double_p, c_double_p, C.c_int]
        lib.compat_bessel_j.argtypes = [C.c_int, C.c_double]
        lib.compat_bessel_j.restype = C.c_double
        _host = lib
    return _host


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int_p)


def lptr(a):
    return a.ctypes.data_as(c_int64_p)
