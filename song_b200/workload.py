"""Synthetic inputs of the transfer2 stage for the BASELINE workloads.

SONG's own producers of these inputs (CLASS background/thermodynamics, perturbations2.c) cannot
run here, so the sampling logic that shapes the work is replayed from the `.pre` values and the
source functions are seeded synthetic fields (recipe: SURVEY.md section 8d).  What is replayed, with
the reference lines it follows:

* ``l_list``        CLASS transfer_get_l_list rule (log then linear steps; recalled from upstream)
* ``k_list``        perturb2_get_k_lists "smart" sampling, source/perturbations2.c:1641-1765
* ``k3_pt_grid``    smart k3 sampling of the sources, source/perturbations2.c:1963-2186
* ``xx_grid``       input2.c:1982-2006 + bessel2_get_xx_list, source/bessel2.c:960-985
* ``index tables``  transfer2_indices_of_transfers / transfer2_get_lm_lists, transfer2.c:866-1285
* fine k3 grids     transfer2_get_k3_size/_list through the host C library (los_host.c)
"""
from __future__ import annotations

import ctypes as C
import math
import os
import dataclasses
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .configs import COSMO

_MIN_K3_DISTANCE_ = 1e-10   # include/perturbations2.h:178
_MIN_K3_RATIO_ = 100        # include/perturbations2.h:187

INTERP = {"linear": 0, "cubic": 1}
TAU_SAMPLING = {"sources": 0, "custom": 1, "smart": 1, "bessel": 2}
K3_SAMPLING = {"bessel": 0, "smart": 1, "class": 1}
K3_EXTRAPOLATION = {"no_extrapolation": 0, "flat_extrapolation": 1}


# ----------------------------------------------------------------------------------------------
# samplings
# ----------------------------------------------------------------------------------------------
def l_list(l_max: int, l_logstep: float, l_linstep: int) -> np.ndarray:
    """Multipoles of pbs->l: logarithmic steps from l=2 until the step reaches l_linstep, then
    linear, last value exactly l_max (CLASS transfer_get_l_list; recalled from upstream)."""
    ls = [2]
    cur = 2
    inc = max(int(cur * (l_logstep - 1.0)), 1)
    while (cur + inc) < l_max and inc < l_linstep:
        cur += inc
        ls.append(cur)
        inc = max(int(cur * (l_logstep - 1.0)), 1)
    inc = l_linstep
    while (cur + inc) <= l_max:
        cur += inc
        ls.append(cur)
    if cur != l_max:
        ls.append(l_max)
    return np.asarray(ls, dtype=np.int32)


def k_list(p: dict, cosmo: dict = COSMO) -> np.ndarray:
    """ppt2->k, smart sampling (perturbations2.c:1659-1765)."""
    tau0, rs_rec = cosmo["tau0"], cosmo["rs_rec"]
    k_min = p["k_min_tau0_song"] / tau0
    k_rec = 2 * math.pi / rs_rec
    k_max = p["k_max_tau0_over_l_max_song"] * p["l_max_scalars"] / tau0
    ks = [k_min]
    k = k_min
    while k < k_max:
        prev = ks[-1]
        lin_step = p["k_step_super_song"] + 0.5 * (math.tanh((prev - k_rec) / k_rec / p["k_step_transition_song"]) + 1) \
            * (p["k_step_sub_song"] - p["k_step_super_song"])
        log_step = prev * (p["k_logstep_super_song"] - 1.0)
        if (log_step > (lin_step * k_rec)) or (prev > k_rec):
            k = prev + lin_step * k_rec
        else:
            k = prev * p["k_logstep_super_song"]
        ks.append(k)
    ks[-1] = k_max                                   # perturbations2.c:1765
    return np.asarray(ks, dtype=np.float64)


def k3_pt_grid(k: np.ndarray, i1: int, i2: int, k3_size_min: int) -> np.ndarray:
    """ppt2->k3[index_k1][index_k2], smart sampling (perturbations2.c:1966-2186)."""
    k1, k2 = k[i1], k[i2]
    k3_min = abs(k1 - k2) + abs(_MIN_K3_DISTANCE_)
    k3_max = k1 + k2 - abs(_MIN_K3_DISTANCE_)
    k3_min = max(k3_min, (k1 + k2) / _MIN_K3_RATIO_)
    k3_min = max(k3_min, k[0])
    k3_max = min(k3_max, k[-1])
    lo = int(np.searchsorted(k, k3_min, side="left"))       # first k >= k3_min  (:2109)
    hi = int(np.searchsorted(k, k3_max, side="right")) - 1  # last  k <= k3_max  (:2115)
    grid = np.union1d(k[lo:hi + 1], np.array([k3_min, k3_max]))   # merge_arrays_double (:2139)
    if grid.size < k3_size_min:                                   # :2156-2184
        n_extra = k3_size_min - grid.size + 2
        extra = k3_min + np.arange(n_extra) * (k3_max - k3_min) / (n_extra - 1)
        extra[-1] = k3_max
        grid = np.union1d(grid, extra)
    return grid


def tau_sampling(tau_size: int) -> np.ndarray:
    """Synthetic ppt2->tau_sampling on [230, 700] Mpc: the first half geometric across
    recombination (230 -> 330), the rest linear (stated assumption)."""
    n1 = tau_size // 2
    a = 230.0 * (330.0 / 230.0) ** (np.arange(n1) / n1)
    b = np.linspace(330.0, 700.0, tau_size - n1)
    return np.concatenate([a, b])


def xx_grid(p: dict):
    """pbs2->xx: xx_max = k_max_tau0_over_l_max * l_max (+ step, * 1.05), size truncated."""
    xx_step = p["bessel_x_step_song"]
    xx_max = p["k_max_tau0_over_l_max_song"] * p["l_max_scalars"]     # input2.c:1982
    xx_max += xx_step                                                 # input2.c:2005
    xx_max *= 1.05                                                    # input2.c:2006
    xx_size = int(xx_max / xx_step + 1)                               # bessel2.c:979
    xx = xx_max * np.arange(xx_size) / (xx_size - 1)                  # lin_space
    xx[-1] = xx_max
    return xx, xx_step, xx_max


# ----------------------------------------------------------------------------------------------
# index tables (transfer2_indices_of_transfers, transfer2.c:866-1285)
# ----------------------------------------------------------------------------------------------
@dataclass
class Layout:
    l: np.ndarray
    m: np.ndarray
    L_size: int
    l_max_los_t: int
    l_max_los_p: int
    has_source_T: bool
    has_source_E: bool
    has_source_B: bool
    n_transfers: int = 0
    lm_array: np.ndarray = None       # [l_size][m_size], -1 where m > l
    lm_src: np.ndarray = None         # [L_size][m_size], -1 where m > L
    n_sources_T: int = 0
    n_sources_P: int = 0
    index_tp2_T: int = -1
    index_tp2_E: int = -1
    index_tp2_B: int = -1
    tp2_size: int = 0
    index_tt2_T: int = -1
    index_tt2_E: int = -1
    index_tt2_B: int = -1
    tt2_size: int = 0
    tt_index_l: np.ndarray = None
    tt_index_m: np.ndarray = None
    tt_labels: list = field(default_factory=list)
    J_size: int = 0
    index_J_TT: int = -1
    index_J_EE: int = -1
    index_J_EB: int = -1

    @property
    def l_size(self):
        return int(self.l.size)

    @property
    def m_size(self):
        return int(self.m.size)


def make_layout(l, m, l_max_los_t, l_max_los_p, has_T=True, has_E=True, has_B=False) -> Layout:
    l = np.asarray(l, dtype=np.int32)
    m = np.asarray(m, dtype=np.int32)
    L_max = max(l_max_los_t if has_T else 0, l_max_los_p if (has_E or has_B) else 0, int(m.max()))
    lay = Layout(l=l, m=m, L_size=L_max + 1, l_max_los_t=l_max_los_t, l_max_los_p=l_max_los_p,
                 has_source_T=has_T, has_source_E=has_E, has_source_B=has_B)
    # (l,m) -> offset, l-major / m-minor, only m <= l  (multipole2offset_indexl_indexm)
    lm = -np.ones((l.size, m.size), dtype=np.int32)
    off = 0
    pairs = []
    for il, lv in enumerate(l):
        for im, mv in enumerate(m):
            if mv <= lv:
                lm[il, im] = off
                pairs.append((il, im))
                off += 1
    lay.n_transfers = off
    lay.lm_array = lm
    # lm(L,m) of the sources (multipole2offset_l_indexm)
    lms = -np.ones((lay.L_size, m.size), dtype=np.int32)
    off = 0
    for L in range(lay.L_size):
        for im, mv in enumerate(m):
            if mv <= L:
                lms[L, im] = off
                off += 1
    lay.lm_src = lms

    def n_sources(lmax):   # size_l_indexm
        return int(sum(1 for L in range(lmax + 1) for mv in m if mv <= L))
    lay.n_sources_T = n_sources(l_max_los_t)
    lay.n_sources_P = n_sources(l_max_los_p)
    # source types: T block, then E, then B (perturbations2.c:667-1015)
    tp = 0
    if has_T:
        lay.index_tp2_T = tp; tp += lay.n_sources_T
    if has_E:
        lay.index_tp2_E = tp; tp += lay.n_sources_P
    if has_B:
        lay.index_tp2_B = tp; tp += lay.n_sources_P
    lay.tp2_size = tp
    # transfer types (transfer2.c:917-940)
    tt = 0
    if has_T:
        lay.index_tt2_T = tt; tt += lay.n_transfers
    if has_E:
        lay.index_tt2_E = tt; tt += lay.n_transfers
    if has_B:
        lay.index_tt2_B = tt; tt += lay.n_transfers
    lay.tt2_size = tt
    lay.tt_index_l = np.zeros(tt, dtype=np.int32)
    lay.tt_index_m = np.zeros(tt, dtype=np.int32)
    for name, base, on in (("T", lay.index_tt2_T, has_T), ("E", lay.index_tt2_E, has_E), ("B", lay.index_tt2_B, has_B)):
        if not on:
            continue
        for o, (il, im) in enumerate(pairs):
            lay.tt_index_l[base + o] = il
            lay.tt_index_m[base + o] = im
    lay.tt_labels = [""] * tt
    for name, base, on in (("T", lay.index_tt2_T, has_T), ("E", lay.index_tt2_E, has_E), ("B", lay.index_tt2_B, has_B)):
        if on:
            for o, (il, im) in enumerate(pairs):
                lay.tt_labels[base + o] = f"{name}_{l[il]}_{m[im]}"      # transfer2.c:1211
    # projection function types (bessel2.c:145-173)
    j = 0
    if has_T:
        lay.index_J_TT = j; j += 1
    if has_E or has_B:
        lay.index_J_EE = j; j += 1
        if int(m.max()) > 0:
            lay.index_J_EB = j; j += 1
    lay.J_size = j
    return lay


# ----------------------------------------------------------------------------------------------
# the workload
# ----------------------------------------------------------------------------------------------
@dataclass
class Workload:
    name: str
    params: dict
    cosmo: dict
    layout: Layout
    k: np.ndarray
    tau: np.ndarray
    xx: np.ndarray
    xx_step: float
    xx_max: float
    pairs: np.ndarray            # [n_pairs, 2] (index_k1, index_k2), k1 >= k2, reference loop order
    k3_pt_size: np.ndarray
    k3_tr_size: np.ndarray
    phys_start: np.ndarray
    phys_size: np.ndarray
    pt_off: np.ndarray
    tr_off: np.ndarray
    src_off: np.ndarray
    k3_pt: np.ndarray
    k3_tr: np.ndarray
    k_min_tr: np.ndarray
    k_max_tr: np.ndarray
    src: np.ndarray = None       # packed [pair][tp][tau][k3_pt]
    # slices of fine grids (k3 bands, song_b200/shard.py): k3 of the first / last physical point of every pair's FULL grid
    # (phys_start is then relative to the slice and may lie outside it), the parent workload, and for every pair of this
    # workload its index in the parent and the first unit of the slice inside the parent pair's fine grid
    phys_k_first: np.ndarray = None
    phys_k_last: np.ndarray = None
    parent: "Workload" = None
    parent_pair: np.ndarray = None
    slice_start: np.ndarray = None

    @property
    def n_pairs(self):
        return int(self.pairs.shape[0])

    @property
    def n_units(self):
        return int(self.k3_tr_size.sum())

    @property
    def n_integrals(self):
        return self.n_units * self.layout.tt2_size

    def k3_params(self) -> _lib.K3Params:
        return k3_params_of(self.params, self.cosmo, self.k, self.xx_max, self.xx_step)

    def select(self, idx) -> "Workload":
        """The sub-workload made of pairs `idx` (packed arrays rebuilt)."""
        idx = np.asarray(idx, dtype=np.int64)
        if idx.size == self.n_pairs and np.array_equal(idx, np.arange(self.n_pairs)):      # all pairs: no copy of the big arrays
            return dataclasses.replace(self, pairs=self.pairs.copy(), k3_tr_size=self.k3_tr_size.copy(), tr_off=self.tr_off.copy(),
                                       phys_start=self.phys_start.copy(), phys_size=self.phys_size.copy())
        tp2, ntau = self.layout.tp2_size, self.tau.size
        np_ = self.k3_pt_size[idx]
        nt_ = self.k3_tr_size[idx]
        pt_off = np.concatenate([[0], np.cumsum(np_)[:-1]]).astype(np.int64) if idx.size else np.zeros(0, np.int64)
        tr_off = np.concatenate([[0], np.cumsum(nt_)[:-1]]).astype(np.int64) if idx.size else np.zeros(0, np.int64)
        ssz = np_.astype(np.int64) * tp2 * ntau
        src_off = np.concatenate([[0], np.cumsum(ssz)[:-1]]).astype(np.int64) if idx.size else np.zeros(0, np.int64)
        k3_pt = np.concatenate([self.k3_pt[self.pt_off[i]:self.pt_off[i] + self.k3_pt_size[i]] for i in idx]) if idx.size else np.zeros(0)
        k3_tr = np.concatenate([self.k3_tr[self.tr_off[i]:self.tr_off[i] + self.k3_tr_size[i]] for i in idx]) if idx.size else np.zeros(0)
        src = None
        if self.src is not None:
            src = np.concatenate([self.src[self.src_off[i]:self.src_off[i] + int(self.k3_pt_size[i]) * tp2 * ntau] for i in idx]) if idx.size else np.zeros(0)
        sl = {}
        if self.parent is not None:                   # pairs of a sliced workload keep their slices
            sl = dict(phys_k_first=np.ascontiguousarray(self.phys_k_first[idx]), phys_k_last=np.ascontiguousarray(self.phys_k_last[idx]),
                      parent=self.parent, parent_pair=self.parent_pair[idx], slice_start=self.slice_start[idx])
        return Workload(name=self.name, params=self.params, cosmo=self.cosmo, layout=self.layout, k=self.k,
                        tau=self.tau, xx=self.xx, xx_step=self.xx_step, xx_max=self.xx_max,
                        pairs=self.pairs[idx], k3_pt_size=np.ascontiguousarray(np_), k3_tr_size=np.ascontiguousarray(nt_),
                        phys_start=np.ascontiguousarray(self.phys_start[idx]), phys_size=np.ascontiguousarray(self.phys_size[idx]),
                        pt_off=pt_off, tr_off=tr_off, src_off=src_off, k3_pt=np.ascontiguousarray(k3_pt),
                        k3_tr=np.ascontiguousarray(k3_tr), k_min_tr=self.k_min_tr[idx], k_max_tr=self.k_max_tr[idx], src=src, **sl)

    def select_units(self, idx, first, last) -> "Workload":
        """The sub-workload made of the fine-grid slices [first[i], last[i]) of pairs `idx` (a k3 band of the multi-GPU
        schedule): coarse grids and sources stay whole -- the k3 spline needs the whole column -- the fine grid is cut and
        the physical range is re-expressed relative to the slice (include/song_los.h, song_los_batch)."""
        assert self.parent is None, "slices of slices are not needed"
        idx = np.asarray(idx, dtype=np.int64); first = np.asarray(first, dtype=np.int64); last = np.asarray(last, dtype=np.int64)
        assert np.all(first >= 0) and np.all(last <= self.k3_tr_size[idx]) and np.all(last > first)
        sub = self.select(idx)
        nt = (last - first).astype(np.int32)
        sub.tr_off = np.concatenate([[0], np.cumsum(nt)[:-1]]).astype(np.int64) if idx.size else np.zeros(0, np.int64)
        sub.k3_tr = np.ascontiguousarray(np.concatenate([self.k3_tr[self.tr_off[i] + a:self.tr_off[i] + b] for i, a, b in zip(idx, first, last)])
                                         if idx.size else np.zeros(0))
        ps, pz = self.phys_start[idx].astype(np.int64), self.phys_size[idx].astype(np.int64)
        sub.phys_k_first = np.ascontiguousarray(self.k3_tr[self.tr_off[idx] + ps])
        sub.phys_k_last = np.ascontiguousarray(self.k3_tr[self.tr_off[idx] + ps + pz - 1])
        sub.phys_start = np.ascontiguousarray((ps - first).astype(np.int32))
        sub.k3_tr_size = np.ascontiguousarray(nt)
        sub.k_min_tr = np.array([sub.k3_tr[o] for o in sub.tr_off]); sub.k_max_tr = np.array([sub.k3_tr[o + n - 1] for o, n in zip(sub.tr_off, nt)])
        sub.parent = self; sub.parent_pair = idx.copy(); sub.slice_start = first.copy()
        return sub


def k3_params_of(p, cosmo, k, xx_max, xx_step) -> _lib.K3Params:
    return _lib.K3Params(
        tau0=cosmo["tau0"], tau_rec=cosmo["tau_rec"], rs_rec=cosmo["rs_rec"],
        k_first=float(k[0]), k_last=float(k[-1]),
        k3_extrapolation=K3_EXTRAPOLATION[p["bispectra_k3_extrapolation"]],
        extra_k3_oscillations_left=float(p["extra_k3_oscillations_left"]),
        extra_k3_oscillations_right=float(p["extra_k3_oscillations_right"]),
        k_sampling=K3_SAMPLING[p["transfer2_k3_sampling"]],
        q_linstep_song=float(p["q_linstep_song"]), old_run_step=0,
        smallest_allowed_variation=float(p["smallest_allowed_variation"]),
        xx_max=float(xx_max), xx_step=float(xx_step))


def layout_of(p: dict) -> Layout:
    out = p.get("output", ["tBisp", "eBisp"])
    has_T = "tBisp" in out or "tCl" in out
    has_E = "eBisp" in out or "pCl" in out
    m = np.asarray(p.get("modes_song", [0]), dtype=np.int32)
    has_B = bool(p.get("has_source_B", has_E and int(m.max()) > 0))     # perturbations2.c:909
    l = p["l"] if "l" in p else l_list(p["l_max_scalars"], p["l_logstep"], p["l_linstep"])
    return make_layout(l, m, p["l_max_los_t"], p["l_max_los_p"], has_T, has_E, has_B)


def build_grids(p: dict, cosmo: dict = COSMO, max_pairs: int | None = None, pair_stride: int = 1) -> Workload:
    """All samplings of a workload (no sources yet).  `pair_stride`/`max_pairs` sub-sample the
    (k1,k2) pairs -- keeping the reference loop order -- for bounded CPU baselines."""
    lay = layout_of(p)
    k = np.asarray(p["k"], dtype=np.float64) if "k" in p else k_list(p, cosmo)
    tau = np.asarray(p["tau_sampling"], dtype=np.float64) if "tau_sampling" in p else tau_sampling(p["tau_size"])
    xx, xx_step, xx_max = xx_grid(p)
    pairs = np.asarray([(i1, i2) for i1 in range(k.size) for i2 in range(i1 + 1)], dtype=np.int32)
    if pair_stride > 1:
        pairs = pairs[::pair_stride]
    if max_pairs is not None:
        pairs = pairs[:max_pairs]
    grids = [k3_pt_grid(k, int(a), int(b), p["k3_size_min"]) for a, b in pairs]
    np_ = np.asarray([g.size for g in grids], dtype=np.int32)
    pt_off = np.concatenate([[0], np.cumsum(np_)[:-1]]).astype(np.int64)
    k3_pt = np.ascontiguousarray(np.concatenate(grids))
    # fine grids through the host C library (transfer2_get_k3_sizes / _list)
    host = _lib.host()
    kp = k3_params_of(p, cosmo, k, xx_max, xx_step)
    n = pairs.shape[0]
    ranges = (_lib.K3Range * n)()
    err = C.create_string_buffer(512)
    kmax = C.c_int(0)
    ksum = C.c_int64(0)
    if host.song_host_k3_sizes(C.byref(kp), n, _lib.iptr(np_), _lib.lptr(pt_off), _lib.dptr(k3_pt), ranges,
                               C.byref(kmax), C.byref(ksum), err, 512):
        raise _lib.SongLosError(err.value.decode())
    nt_ = np.asarray([r.k_size for r in ranges], dtype=np.int32)
    tr_off = np.concatenate([[0], np.cumsum(nt_)[:-1]]).astype(np.int64)
    k3_tr = np.zeros(int(ksum.value), dtype=np.float64)
    if host.song_host_k3_lists(C.byref(kp), n, _lib.iptr(np_), _lib.lptr(pt_off), _lib.dptr(k3_pt), ranges,
                               _lib.lptr(tr_off), _lib.dptr(k3_tr), err, 512):
        raise _lib.SongLosError(err.value.decode())
    ssz = np_.astype(np.int64) * lay.tp2_size * tau.size
    src_off = np.concatenate([[0], np.cumsum(ssz)[:-1]]).astype(np.int64)
    return Workload(
        name=p.get("name", "custom"), params=p, cosmo=cosmo, layout=lay, k=k, tau=tau, xx=xx, xx_step=xx_step,
        xx_max=xx_max, pairs=pairs, k3_pt_size=np_, k3_tr_size=nt_,
        phys_start=np.asarray([r.k_physical_start for r in ranges], dtype=np.int32),
        phys_size=np.asarray([r.k_physical_size for r in ranges], dtype=np.int32),
        pt_off=pt_off, tr_off=tr_off, src_off=src_off, k3_pt=k3_pt, k3_tr=k3_tr,
        k_min_tr=np.asarray([r.k_min for r in ranges]), k_max_tr=np.asarray([r.k_max for r in ranges]))


def fill_sources(wl: Workload, kind: str = "smooth", seed: int = 20261017) -> Workload:
    """ppt2->sources for every pair, layout [tp][tau][k3_pt] (transfer2.h:47-51).

    smooth : S_L0 = A_L g(tau) cos(k3 r_s(tau) + phi_L) (k3/k_rec)^min(L,2) exp(-(k3/0.15)^2)
             (SURVEY.md 8d); E-type L<2 are zero (perturbations2.c:12442-12448)
    noise  : i.i.d. N(0,1) -- adversarial for summation-order sensitivity
    """
    lay, tau = wl.layout, wl.tau
    ntau, tp2 = tau.size, lay.tp2_size
    k_rec = 2 * math.pi / wl.cosmo["rs_rec"]
    g = np.exp(-0.5 * ((tau - wl.cosmo["tau_rec"]) / 18.0) ** 2)
    rs = tau / math.sqrt(3.0)
    src = np.empty(int((wl.k3_pt_size.astype(np.int64) * tp2 * ntau).sum()), dtype=np.float64)
    # source index -> (field, L, m)
    tp_info = []
    for field_name, base, lmax, on in (("T", lay.index_tp2_T, lay.l_max_los_t, lay.has_source_T),
                                       ("E", lay.index_tp2_E, lay.l_max_los_p, lay.has_source_E),
                                       ("B", lay.index_tp2_B, lay.l_max_los_p, lay.has_source_B)):
        if not on:
            continue
        for L in range(lmax + 1):
            for im, mv in enumerate(lay.m):
                if mv <= L:
                    tp_info.append((base + lay.lm_src[L, im], field_name, L, int(mv)))
    tp_info.sort()
    # per source type: index of the k3 power (min(L,2)) and whether the type is identically zero (E, B with L < 2)
    tp_index = np.array([t[0] for t in tp_info], dtype=np.int64)
    tp_pow = np.array([min(t[2], 2) for t in tp_info], dtype=np.int64)
    tp_live = np.array([not (t[1] != "T" and t[2] < 2) for t in tp_info])
    tp_norm = np.array([1.0 / (t[2] + 1) ** 2 for t in tp_info])
    def fill_pair(ip):
        i1, i2 = int(wl.pairs[ip, 0]), int(wl.pairs[ip, 1])
        rng = np.random.Generator(np.random.PCG64(seed + 1000 * i1 + i2))
        n = int(wl.k3_pt_size[ip])
        k3 = wl.k3_pt[wl.pt_off[ip]:wl.pt_off[ip] + n]
        blk = src[wl.src_off[ip]:wl.src_off[ip] + tp2 * ntau * n].reshape(tp2, ntau, n)
        if kind == "noise":
            blk[...] = rng.standard_normal(blk.shape)
            return
        amp = rng.uniform(0.5, 1.5, size=tp2)
        phi = rng.uniform(0.0, 2 * math.pi, size=tp2)
        env = np.exp(-(k3 / 0.15) ** 2)
        # cos(k3 r_s + phi_tp) = cos(k3 r_s) cos(phi_tp) - sin(k3 r_s) sin(phi_tp): the two tau x k3 tables are evaluated once
        # per pair and every source type is a combination of them (the generator, not the GPU, dominated the set-up of the
        # big workloads); one broadcast expression per pair
        arg = k3[None, :] * rs[:, None]
        cg, sg = np.cos(arg) * g[:, None], np.sin(arg) * g[:, None]
        kpow = np.stack([env, (k3 / k_rec) * env, (k3 / k_rec) ** 2 * env])               # [3][n]
        a = np.where(tp_live, amp[tp_index] * tp_norm, 0.0)
        ca, sa = a * np.cos(phi[tp_index]), a * np.sin(phi[tp_index])
        blk[...] = 0.0
        blk[tp_index] = (ca[:, None, None] * cg[None, :, :] - sa[:, None, None] * sg[None, :, :]) * kpow[tp_pow][:, None, :]

    # pairs are independent (own generator each): a few threads, numpy releases the GIL inside the array expressions
    n_threads = max(1, min(16, int(os.environ.get("SONG_FILL_THREADS", "0")) or
                           len(os.sched_getaffinity(0)) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
    if n_threads == 1 or wl.n_pairs < 64:
        for ip in range(wl.n_pairs):
            fill_pair(ip)
    else:
        from concurrent.futures import ThreadPoolExecutor
        chunks = [range(a, min(a + 32, wl.n_pairs)) for a in range(0, wl.n_pairs, 32)]
        with ThreadPoolExecutor(n_threads) as pool:
            list(pool.map(lambda r: [fill_pair(ip) for ip in r], chunks))
    wl.src = src
    return wl


def make_batch(wl: Workload) -> _lib.LosBatch:
    """The C-ABI view of a workload (arrays stay owned by `wl`)."""
    assert wl.src is not None, "fill_sources first"
    return _lib.LosBatch(
        n_pairs=wl.n_pairs, k3_pt_size=_lib.iptr(wl.k3_pt_size), k3_tr_size=_lib.iptr(wl.k3_tr_size),
        phys_start=_lib.iptr(wl.phys_start), phys_size=_lib.iptr(wl.phys_size),
        pt_off=_lib.lptr(wl.pt_off), tr_off=_lib.lptr(wl.tr_off), src_off=_lib.lptr(wl.src_off),
        n_pt_total=int(wl.k3_pt.size), n_units=int(wl.k3_tr.size), n_src_total=int(wl.src.size),
        k3_pt=_lib.dptr(wl.k3_pt), k3_tr=_lib.dptr(wl.k3_tr), src=_lib.dptr(wl.src),
        phys_k_first=None if wl.phys_k_first is None else _lib.dptr(wl.phys_k_first),
        phys_k_last=None if wl.phys_k_last is None else _lib.dptr(wl.phys_k_last))


def describe(wl: Workload) -> dict:
    lay = wl.layout
    return {
        "workload": wl.name, "k_size": int(wl.k.size), "pairs": wl.n_pairs, "k3_units": wl.n_units,
        "k3_tr_avg": float(wl.k3_tr_size.mean()), "k3_tr_max": int(wl.k3_tr_size.max()),
        "k3_pt_avg": float(wl.k3_pt_size.mean()), "l_size": lay.l_size, "tt2_size": lay.tt2_size,
        "tp2_size": lay.tp2_size, "L_max": lay.L_size - 1, "tau_size": int(wl.tau.size),
        "xx_size": int(wl.xx.size), "xx_step": wl.xx_step, "integrals": wl.n_integrals,
        "tau0": wl.cosmo["tau0"], "tau_rec": wl.cosmo["tau_rec"], "rs_rec": wl.cosmo["rs_rec"],
    }
