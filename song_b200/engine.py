"""Thin Python handle on the C-ABI LOS engine (``include/song_los.h``).

This is plumbing for the tests, ``bench.py`` and the multi-GPU driver; SONG itself reaches the
same entry points from C through ``song_b200/host/transfer2_gpu.c``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .workload import INTERP, K3_EXTRAPOLATION, TAU_SAMPLING, Workload, make_batch

MODE_EXACT = 0
MODE_FAST = 1


class Projection:
    """Host copy of the bessels2 tables the engine consumes (bessel2.h:64-91), flattened as
    jidx = ((index_J*L_size + L)*l_size + index_l)*m_size + index_m."""

    def __init__(self, rows, index_xmin, x_size, ddrows=None):
        self.rows = rows                  # list of 1-D float64 arrays or None
        self.index_xmin = np.ascontiguousarray(index_xmin, dtype=np.int32)
        self.x_size = np.ascontiguousarray(x_size, dtype=np.int32)
        self.ddrows = ddrows

    def _ptrs(self, rows):
        arr = (_lib.c_double_p * len(rows))()
        for i, r in enumerate(rows):
            if r is not None:
                arr[i] = _lib.dptr(r)
        return arr

    def row_ptrs(self):
        return self._ptrs(self.rows)

    def ddrow_ptrs(self):
        return self._ptrs(self.ddrows) if self.ddrows is not None else None


class PinnedArray:
    """A numpy view of page-locked host memory from song_los_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype=np.float64):
        self.lib = _lib.los()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        if self.lib.song_los_host_alloc(C.byref(self.ptr), n):
            raise _lib.SongLosError(self.lib.song_los_last_error(None).decode())
        buf = (C.c_char * max(n, 1)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self.ptr:
                self.array = None
                self.lib.song_los_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


class LosEngine:
    """One engine per GPU.  Fails loudly when the CUDA library or a device is missing."""

    def __init__(self, wl: Workload, device: int = 0):
        self.lib = _lib.los()
        lay, p = wl.layout, wl.params
        self._keep = dict(tau=np.ascontiguousarray(wl.tau), xx=np.ascontiguousarray(wl.xx),
                          l=np.ascontiguousarray(lay.l), m=np.ascontiguousarray(lay.m),
                          lm_array=np.ascontiguousarray(lay.lm_array), lm_src=np.ascontiguousarray(lay.lm_src),
                          tl=np.ascontiguousarray(lay.tt_index_l), tm=np.ascontiguousarray(lay.tt_index_m))
        k = self._keep
        cfg = _lib.LosConfig(
            device=device, tau0=wl.cosmo["tau0"], tau_size=int(wl.tau.size), tau_sampling=_lib.dptr(k["tau"]),
            tau_sampling_mode=TAU_SAMPLING[p["transfer2_tau_sampling"]],
            tau_linstep_song=float(p.get("tau_linstep_song", 0.0)),
            sources_k3_interpolation=INTERP[p["sources_k3_interpolation"]],
            sources_time_interpolation=INTERP[p["sources_time_interpolation"]],
            bessels_interpolation=INTERP[p["bessels_interpolation"]],
            k3_extrapolation=K3_EXTRAPOLATION[p["bispectra_k3_extrapolation"]],
            xx_size=int(wl.xx.size), xx=_lib.dptr(k["xx"]), xx_step=float(wl.xx_step),
            l_size=lay.l_size, l=_lib.iptr(k["l"]), m_size=lay.m_size, m=_lib.iptr(k["m"]),
            L_size=lay.L_size, n_transfers=lay.n_transfers,
            lm_array=_lib.iptr(k["lm_array"]), lm_src=_lib.iptr(k["lm_src"]),
            has_source_T=int(lay.has_source_T), has_source_E=int(lay.has_source_E), has_source_B=int(lay.has_source_B),
            index_tp2_T=lay.index_tp2_T, index_tp2_E=lay.index_tp2_E, index_tp2_B=lay.index_tp2_B,
            tp2_size=lay.tp2_size, l_max_los_t=lay.l_max_los_t, l_max_los_p=lay.l_max_los_p,
            index_tt2_T=lay.index_tt2_T, index_tt2_E=lay.index_tt2_E, index_tt2_B=lay.index_tt2_B,
            tt2_size=lay.tt2_size, tt_index_l=_lib.iptr(k["tl"]), tt_index_m=_lib.iptr(k["tm"]),
            J_size=lay.J_size, index_J_TT=lay.index_J_TT, index_J_EE=lay.index_J_EE, index_J_EB=lay.index_J_EB)
        self.ctx = C.c_void_p()
        if self.lib.song_los_create(C.byref(cfg), C.byref(self.ctx)):
            raise _lib.SongLosError(self.lib.song_los_last_error(None).decode())
        self.tt2_size = lay.tt2_size
        self.n_units = 0
        self.timing = _lib.LosTiming()

    # -- helpers ---------------------------------------------------------------------------
    def _check(self, rc):
        if rc:
            raise _lib.SongLosError(self.lib.song_los_last_error(self.ctx).decode())

    def close(self):
        if self.ctx:
            self.lib.song_los_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- projection functions --------------------------------------------------------------
    def set_projection(self, proj: Projection):
        dd = proj.ddrow_ptrs()
        self._check(self.lib.song_los_set_projection(self.ctx, proj.row_ptrs(), dd,
                                                     _lib.iptr(proj.index_xmin), _lib.iptr(proj.x_size)))

    # -- batches ---------------------------------------------------------------------------
    def select_slot(self, slot: int):
        """Which of the three resident batch slots upload / compute / download work on."""
        self._check(self.lib.song_los_select_slot(self.ctx, int(slot)))

    def upload(self, wl: Workload):
        b = make_batch(wl)
        self._check(self.lib.song_los_upload_pairs(self.ctx, C.byref(b)))
        self.n_units = wl.n_units

    def compute(self, mode: int = MODE_EXACT) -> dict:
        self._check(self.lib.song_los_compute(self.ctx, mode, C.byref(self.timing)))
        return self.timing.as_dict()

    def download(self) -> np.ndarray:
        out = np.empty((self.tt2_size, self.n_units), dtype=np.float64)
        self._check(self.lib.song_los_download(self.ctx, _lib.dptr(out)))
        return out

    def transfer_pairs(self, wl: Workload, mode: int = MODE_EXACT, out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in, host buffer out: upload + compute + download (the e2e path)."""
        b = make_batch(wl)
        if out is None:
            out = np.empty((self.tt2_size, wl.n_units), dtype=np.float64)
        self._check(self.lib.song_los_transfer_pairs(self.ctx, C.byref(b), mode, _lib.dptr(out), C.byref(self.timing)))
        self.n_units = wl.n_units
        return out

    # -- asynchronous pipeline (song_los_submit / song_los_wait) ------------------------------
    def submit(self, wl: Workload, mode: int, out: np.ndarray):
        """Enqueue upload + kernels + download of `wl`; `wl`'s arrays and `out` must stay alive (and
        should be pinned, see :class:`PinnedArray`) until the matching :meth:`wait` returns."""
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self.tt2_size * wl.n_units
        b = make_batch(wl)
        self._inflight = getattr(self, "_inflight", [])
        self._inflight.append((b, wl, out))
        self._check(self.lib.song_los_submit(self.ctx, C.byref(b), mode, _lib.dptr(out)))
        self.n_units = wl.n_units

    def wait(self) -> dict:
        """Block until the oldest submitted batch is complete; returns its timing."""
        self._check(self.lib.song_los_wait(self.ctx, C.byref(self.timing)))
        if getattr(self, "_inflight", None):
            self._inflight.pop(0)
        return self.timing.as_dict()

    # -- k3 integral of the intrinsic bispectrum on the device-resident transfer functions ----------
    def set_bessel_table(self, j_l1, x_min_l1, ddj_l1=None):
        """pbs2->j_l1 rows (list of 1-D arrays), their first abscissa x_min_l1 and, for cubic interpolation, ddj_l1."""
        rows = [np.ascontiguousarray(r, dtype=np.float64) for r in j_l1]
        dd = None if ddj_l1 is None else [np.ascontiguousarray(r, dtype=np.float64) for r in ddj_l1]
        n = len(rows)
        rp = (_lib.c_double_p * n)(*[_lib.dptr(r) for r in rows])
        dp_ = None if dd is None else (_lib.c_double_p * n)(*[_lib.dptr(r) for r in dd])
        xmin = np.ascontiguousarray(x_min_l1, dtype=np.float64)
        size = np.asarray([r.size for r in rows], dtype=np.int32)
        self._check(self.lib.song_los_set_bessel_table(self.ctx, n, rp, dp_, _lib.dptr(xmin), _lib.iptr(size)))

    def integrate_k3(self, tt, index_l1, r, n_pairs: int, abs_M3: int = 0, M3_coefficient=None) -> np.ndarray:
        """bispectra2_intrinsic_integrate_over_k3 on the result of the last compute: [len(tt)][len(r)][n_pairs]."""
        tt = np.ascontiguousarray(tt, dtype=np.int32); index_l1 = np.ascontiguousarray(index_l1, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64)
        coef = None if M3_coefficient is None else np.ascontiguousarray(M3_coefficient, dtype=np.float64)
        rq = _lib.K3Request(n_requests=int(tt.size), tt=_lib.iptr(tt), index_l1=_lib.iptr(index_l1), abs_M3=int(abs_M3),
                            M3_coefficient=None if coef is None else _lib.dptr(coef), r_size=int(r.size), r=_lib.dptr(r))
        out = np.empty((tt.size, r.size, n_pairs), dtype=np.float64)
        self._check(self.lib.song_los_integrate_k3(self.ctx, C.byref(rq), _lib.dptr(out)))
        self.k3_ms = float(self.lib.song_los_last_k3_ms(self.ctx))
        return out

    def cl_sum(self, wl, tt1, tt2, m, primordial, rescale_cmb_sources: bool = True) -> np.ndarray:
        """spectra2_cls's (k1,k2,k3) sum over the pairs of `wl` (the batch of the last compute): one number per request."""
        tt1 = np.ascontiguousarray(tt1, dtype=np.int32); tt2 = np.ascontiguousarray(tt2, dtype=np.int32)
        m = np.ascontiguousarray(m, dtype=np.int32)
        k = np.ascontiguousarray(wl.k, dtype=np.float64); prim = np.ascontiguousarray(primordial, dtype=np.float64)
        k1 = np.ascontiguousarray(wl.pairs[:, 0], dtype=np.int32); k2 = np.ascontiguousarray(wl.pairs[:, 1], dtype=np.int32)
        rq = _lib.ClRequest(n_requests=int(tt1.size), tt1=_lib.iptr(tt1), tt2=_lib.iptr(tt2), m=_lib.iptr(m),
                            rescale_cmb_sources=int(rescale_cmb_sources), k_size=int(k.size), k=_lib.dptr(k),
                            primordial=_lib.dptr(prim), pair_k1=_lib.iptr(k1), pair_k2=_lib.iptr(k2))
        out = np.empty(tt1.size, dtype=np.float64)
        self._check(self.lib.song_los_cl_sum(self.ctx, C.byref(rq), _lib.dptr(out)))
        self.k3_ms = float(self.lib.song_los_last_k3_ms(self.ctx))
        return out

    def max_batch_units(self, mode: int, src_doubles_per_unit: float, n_slots: int = 3) -> int:
        """k3 values per batch so that `n_slots` batches (3: the pipeline of song_los_submit) fit in device memory."""
        n = C.c_int64(0)
        self._check(self.lib.song_los_max_resident_units(self.ctx, mode, float(src_doubles_per_unit), int(n_slots), C.byref(n)))
        return int(n.value)

    def reserve(self, mode: int, max_pairs: int, max_units: int, max_pt: int, max_src: int, n_in_flight: int = 3):
        self._check(self.lib.song_los_reserve(self.ctx, mode, int(n_in_flight), int(max_pairs), int(max_units), int(max_pt), int(max_src)))

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.song_los_set_stream(self.ctx, C.c_void_p(cuda_stream or 0)))

    def device_result(self):
        n = C.c_int64(0)
        ptr = self.lib.song_los_device_result(self.ctx, C.byref(n))
        return int(ptr), int(n.value)

    def copy_result_to(self, device_ptr: int, n_doubles: int):
        """Device-to-device copy of the last result into caller-owned device memory (e.g. a torch tensor)."""
        self._check(self.lib.song_los_copy_result(self.ctx, C.c_uint64(device_ptr), C.c_int64(n_doubles)))

    def measure_peaks(self, copy: bool = True):
        f, t, c = C.c_double(0), C.c_double(0), C.c_double(0)
        self._check(self.lib.song_los_measure_peaks(self.ctx, C.byref(f), C.byref(t), C.byref(c) if copy else None))
        return {"dfma_tflops": f.value, "dmma_tflops": t.value, "copy_gbs": c.value}

    def device_info(self):
        fr, tot, sm = C.c_int64(0), C.c_int64(0), C.c_int(0)
        name = C.create_string_buffer(128)
        self._check(self.lib.song_los_device_info(self.ctx, C.byref(fr), C.byref(tot), C.byref(sm), name, 128))
        return {"free": fr.value, "total": tot.value, "sm_count": sm.value, "name": name.value.decode()}
