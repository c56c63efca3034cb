"""Python mirror of SONG's bessel2 module (``bessel2_init`` / ``bessel2_free``,
reference include/bessel2.h:161-174) on top of the host C half and the device assembly.

``Bessel2(wl).init(engine)`` does what ``bessel2_init(ppr, ppr2, ppt2, pbs, pbs2)`` does:
x grid, l1 list, j_l1 table and 3j symbols on the host (``bessel2_host.c``), J_Llm(x) on the GPU
(``song_los_build_projection``), and leaves the ragged ``J_Llm_x`` / ``index_xmin_J`` /
``x_size_J`` views of ``struct bessels2`` available through ``tables()``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .workload import Workload

J_TT, J_EE, J_EB = 0, 1, 2


class Bessel2:
    def __init__(self, wl: Workload):
        self.wl = wl
        lay = wl.layout
        self.J_types = np.zeros(max(lay.J_size, 1), dtype=np.int32)
        if lay.index_J_TT >= 0:
            self.J_types[lay.index_J_TT] = J_TT
        if lay.index_J_EE >= 0:
            self.J_types[lay.index_J_EE] = J_EE
        if lay.index_J_EB >= 0:
            self.J_types[lay.index_J_EB] = J_EB
        self._host = None

    # -- bessel2_init --------------------------------------------------------------------------
    def prepare(self):
        """Host half: l1 list, j_l1(x), 3j symbols (bessel2.c:187-214, 1131-1167)."""
        wl, lay, p = self.wl, self.wl.layout, self.wl.params
        lib = _lib.host()
        self._l = np.ascontiguousarray(lay.l, dtype=np.int32)
        self._m = np.ascontiguousarray(lay.m, dtype=np.int32)
        ptr = C.POINTER(_lib.Bessel2Host)()
        err = C.create_string_buffer(512)
        rc = lib.song_host_bessel2_prepare(_lib.iptr(self._l), lay.l_size, _lib.iptr(self._m), lay.m_size, lay.L_size - 1,
                                           _lib.dptr(wl.xx), int(wl.xx.size), float(wl.xx_max), lay.J_size,
                                           _lib.iptr(self.J_types), float(p["bessel_j_cut_song"]),
                                           float(p["bessel_tol_x_min"]), float(p["smallest_allowed_variation"]),
                                           C.byref(ptr), err, 512)
        if rc:
            raise _lib.SongLosError(err.value.decode())
        self._host = ptr
        return self

    def init(self, engine):
        """bessel2_init: host half (if not done) + device assembly of J_Llm(x) into `engine`."""
        if self._host is None:
            self.prepare()
        err = C.create_string_buffer(1024)
        rc = _lib.host().song_host_bessel2_upload(engine.ctx, self._host, _lib.iptr(self.J_types),
                                                  float(self.wl.params["bessel_J_cut_song"]), float(self.wl.xx_max), err, 1024)
        if rc:
            raise _lib.SongLosError(err.value.decode())
        return self

    def free(self):
        if self._host is not None:
            _lib.host().song_host_bessel2_free(self._host)
            self._host = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # -- views ---------------------------------------------------------------------------------
    def host_tables(self) -> dict:
        h = self._host.contents
        n1 = h.l1_size
        out = {
            "l1": np.ctypeslib.as_array(h.l1, shape=(n1,)).copy(),
            "index_xmin_l1": np.ctypeslib.as_array(h.index_xmin_l1, shape=(n1,)).copy(),
            "x_size_l1": np.ctypeslib.as_array(h.x_size_l1, shape=(n1,)).copy(),
            "x_min_l1": np.ctypeslib.as_array(h.x_min_l1, shape=(n1,)).copy(),
            "first_3j": np.ctypeslib.as_array(h.first_3j, shape=(h.n_jidx, h.ndim)).copy(),
            "second_3j": np.ctypeslib.as_array(h.second_3j, shape=(h.n_jidx, h.ndim)).copy(),
            "l1_min": np.ctypeslib.as_array(h.l1_min, shape=(h.n_jidx,)).copy(),
            "l1_count": np.ctypeslib.as_array(h.l1_count, shape=(h.n_jidx,)).copy(),
            "index_l1_of_l1min": np.ctypeslib.as_array(h.index_l1_of_l1min, shape=(h.n_jidx,)).copy(),
        }
        out["j_l1"] = [np.ctypeslib.as_array(h.j_l1[i], shape=(int(out["x_size_l1"][i]),)).copy() for i in range(n1)]
        return out

    def tables(self, engine):
        """(rows, index_xmin_J, x_size_J) of struct bessels2 read back from the device."""
        lay = self.wl.layout
        n = lay.J_size * lay.L_size * lay.l_size * lay.m_size
        xmin = np.zeros(n, dtype=np.int32)
        xsize = np.zeros(n, dtype=np.int32)
        lib = _lib.los()
        if lib.song_los_get_projection(engine.ctx, _lib.iptr(xmin), _lib.iptr(xsize), None):
            raise _lib.SongLosError(lib.song_los_last_error(engine.ctx).decode())
        rows = [np.zeros(int(s)) if s > 0 else None for s in xsize]
        ptrs = (_lib.c_double_p * n)()
        for i, r in enumerate(rows):
            if r is not None:
                ptrs[i] = _lib.dptr(r)
        if lib.song_los_get_projection(engine.ctx, None, None, ptrs):
            raise _lib.SongLosError(lib.song_los_last_error(engine.ctx).decode())
        return rows, xmin, xsize

    def dd_tables(self, engine, xsize):
        """pbs2->ddJ_Llm_x (cubic Bessel interpolation, bessel2.c:395-435) read back from the device."""
        lib = _lib.los()
        rows = [np.zeros(int(s)) if s > 0 else None for s in xsize]
        ptrs = (_lib.c_double_p * len(rows))()
        for i, r in enumerate(rows):
            if r is not None:
                ptrs[i] = _lib.dptr(r)
        if lib.song_los_get_projection_dd(engine.ctx, ptrs):
            raise _lib.SongLosError(lib.song_los_last_error(engine.ctx).decode())
        return rows


def build_projection_direct(engine, wl, l1, j_l1_rows, index_xmin_l1, first_3j, second_3j, l1_min, l1_count,
                            index_l1_of_l1min, J_types):
    """song_los_build_projection with caller-provided j_l1 / 3j tables (used by the parity tests to
    feed the device and the oracle the very same inputs)."""
    lib = _lib.los()
    n1 = len(j_l1_rows)
    l1 = np.ascontiguousarray(l1, dtype=np.int32)
    xmin1 = np.ascontiguousarray(index_xmin_l1, dtype=np.int32)
    xsz1 = np.ascontiguousarray([r.size for r in j_l1_rows], dtype=np.int32)
    ptrs = (_lib.c_double_p * n1)()
    keep = [np.ascontiguousarray(r) for r in j_l1_rows]
    for i, r in enumerate(keep):
        ptrs[i] = _lib.dptr(r)
    f1 = np.ascontiguousarray(first_3j, dtype=np.float64)
    f2 = np.ascontiguousarray(second_3j, dtype=np.float64)
    a = np.ascontiguousarray(l1_min, dtype=np.int32)
    b = np.ascontiguousarray(l1_count, dtype=np.int32)
    c = np.ascontiguousarray(index_l1_of_l1min, dtype=np.int32)
    jt = np.ascontiguousarray(J_types, dtype=np.int32)
    rc = lib.song_los_build_projection(engine.ctx, n1, _lib.iptr(l1), ptrs, _lib.iptr(xmin1), _lib.iptr(xsz1),
                                       int(f1.shape[0]), _lib.dptr(f1), _lib.dptr(f2), _lib.iptr(a), _lib.iptr(b),
                                       _lib.iptr(c), _lib.iptr(jt), float(wl.params["bessel_J_cut_song"]), float(wl.xx_max))
    if rc:
        raise _lib.SongLosError(lib.song_los_last_error(engine.ctx).decode())
