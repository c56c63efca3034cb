"""Multi-GPU partition of the (k1,k2) work and the gather of the per-GPU transfer arrays.

The (k1,k2) pairs are independent (reference source/transfer2.c:549-551): every rank computes the
pairs the static longest-processing-time schedule gives it (`song_host_lpt_schedule`, the same C
function the drop-in module uses for its k1 batches) and the results are gathered on rank 0 in the
reference's [tt][k1][k2][k3] order.  The gather is the only collective of the path
(``torch.distributed``: NCCL between GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def lpt_schedule(cost, n_ranks: int):
    """device[i] of every work item and the per-rank loads (C: song_host_lpt_schedule)."""
    cost = np.ascontiguousarray(cost, dtype=np.int64)
    dev = np.zeros(cost.size, dtype=np.int32)
    load = np.zeros(n_ranks, dtype=np.int64)
    if _lib.host().song_host_lpt_schedule(_lib.lptr(cost), int(cost.size), int(n_ranks), _lib.iptr(dev), _lib.lptr(load)):
        raise _lib.SongLosError("song_host_lpt_schedule: bad arguments")
    return dev, load


def pair_cost(wl) -> np.ndarray:
    """Cost of every (k1,k2) pair as transfer2_init computes it (C: song_host_pair_costs): k3 values x (fixed part +
    fraction of the multipole columns active at the pair's k3 range); the factors common to all pairs of a run
    (tt2_size, tau_size, L_max+1) drop out of the schedule."""
    cost = np.zeros(wl.n_pairs, dtype=np.int64)
    kmin = np.ascontiguousarray(wl.k_min_tr, dtype=np.float64); kmax = np.ascontiguousarray(wl.k_max_tr, dtype=np.float64)
    if _lib.host().song_host_pair_costs(wl.n_pairs, _lib.iptr(wl.k3_tr_size), _lib.dptr(kmin), _lib.dptr(kmax),
                                        float(wl.cosmo["tau0"]), int(wl.layout.l.max()), _lib.lptr(cost)):
        raise _lib.SongLosError("song_host_pair_costs: bad arguments")
    return cost


def chunks_per_device(wl, n_dev: int) -> int:
    """Contiguous pair chunks per GPU in the static schedule, the rule of transfer2_init: as many as keep about 40 k3
    values per table row in every chunk (the tensor-core kernel's efficiency falls with the k3 density), at most 3."""
    if n_dev <= 1:
        return 1
    dense_units = 40 * int(wl.xx.size)
    return int(max(1, min(3, (wl.n_units // n_dev) // max(dense_units, 1))))


def schedule_pairs(cost, n_ranks: int, chunks_per_rank: int = 3):
    """The static multi-GPU schedule of the drop-in (C: song_host_schedule_pairs, the function transfer2_init calls):
    the pair sequence in reference order is cut into n_ranks * chunks_per_rank contiguous chunks of about equal cost,
    the chunks are assigned by LPT.  Returns (rank of every pair, per-rank load)."""
    cost = np.ascontiguousarray(cost, dtype=np.int64)
    dev = np.zeros(cost.size, dtype=np.int32)
    load = np.zeros(n_ranks, dtype=np.int64)
    if _lib.host().song_host_schedule_pairs(_lib.lptr(cost), int(cost.size), int(n_ranks), int(chunks_per_rank),
                                            _lib.iptr(dev), _lib.lptr(load)):
        raise _lib.SongLosError("song_host_schedule_pairs: bad arguments")
    return dev, load


DENSE_K3_PER_ROW = 80      # fine-grid values per projection-table row a GPU should keep (the whole lmax3000 run in 2 batches: 85)


def band_count(wl, n_dev: int, dense_per_row: int = DENSE_K3_PER_ROW) -> int:
    """k3 bands of the (pair groups x k3 bands) schedule for n_dev GPUs (C: song_host_band_count)."""
    return int(_lib.host().song_host_band_count(int(wl.n_units), int(wl.xx.size), int(n_dev), int(dense_per_row)))


def rows_per_k(wl) -> float:
    """Rows of the projection table per unit of k3 at a mid-run time: x = k3 (tau0 - tau), row = x / xx_step."""
    return float((wl.cosmo["tau0"] - float(np.median(wl.tau))) / wl.xx_step)


# the cost model of los_host.c (microseconds): C_INTERP + (C0 + C1 a)(1 + RHO0 / rho) per k3 value, C_POINT per coarse point
MODEL = dict(c_interp=0.0167, c0=0.070, c1=0.098, rho0=31.0, c_point=0.245, density_rows=50)


def predicted_ms(wl, own, first, last) -> float:
    """Time the model of los_host.c predicts for a GPU that computes the fine-grid slices [first, last) of pairs `own`
    (numpy restatement of model_unit_cost / song_host_band_edges, used for reporting and by the tests)."""
    own = np.asarray(own, dtype=np.int64)
    if own.size == 0:
        return 0.0
    ks = np.concatenate([wl.k3_tr[int(wl.tr_off[ip]) + int(a):int(wl.tr_off[ip]) + int(b)] for ip, a, b in zip(own, first, last)])
    win = (ks * rows_per_k(wl) / MODEL["density_rows"]).astype(np.int64)
    rho = (np.bincount(win) / MODEL["density_rows"])[win]
    act = np.minimum(1.0, ks * wl.cosmo["tau0"] / float(wl.layout.l.max()))
    us = np.sum(MODEL["c_interp"] + (MODEL["c0"] + MODEL["c1"] * act) * (1.0 + MODEL["rho0"] / rho)) \
        + MODEL["c_point"] * float(wl.k3_pt_size[own].sum())
    return float(us) * 1e-3


def pair_cost_model(wl, n_groups: int) -> np.ndarray:
    """Predicted cost of every pair under the measured model of los_host.c (C: song_host_pair_costs_model), 1/1024 us."""
    cost = np.zeros(wl.n_pairs, dtype=np.int64)
    tr_off = np.ascontiguousarray(wl.tr_off, dtype=np.int64); nt = np.ascontiguousarray(wl.k3_tr_size, dtype=np.int32)
    npt = np.ascontiguousarray(wl.k3_pt_size, dtype=np.int32)
    if _lib.host().song_host_pair_costs_model(wl.n_pairs, _lib.iptr(nt), _lib.lptr(tr_off), _lib.dptr(wl.k3_tr), _lib.iptr(npt),
                                              float(wl.cosmo["tau0"]), int(wl.layout.l.max()), rows_per_k(wl), int(n_groups), _lib.lptr(cost)):
        raise _lib.SongLosError("song_host_pair_costs_model: bad arguments")
    return cost


def schedule_bands(wl, n_dev: int, n_bands: int | None = None, chunks_per_group: int | None = None):
    """The static schedule of a run on n_dev GPUs as n_groups pair groups x n_bands k3 bands (los_host.h).  Returns
    (n_groups, n_bands, group_of_pair, cut) with cut[p, r]..cut[p, r+1] the fine-grid slice of pair p in band r.  Pairs
    and band edges are balanced under the measured cost model of los_host.c (k3 density, active multipole share, coarse
    points touched); the edges are those of each group's own k3 values.  GPU d = group * n_bands + band."""
    n_bands = band_count(wl, n_dev) if n_bands is None else int(n_bands)
    assert n_dev % n_bands == 0, "the number of GPUs must be a multiple of the number of k3 bands"
    n_groups = n_dev // n_bands
    cost = pair_cost_model(wl, n_groups)
    chunks = chunks_per_group if chunks_per_group else (chunks_per_device(wl, n_groups) if n_groups > 1 else 1)
    group_of_pair, _ = schedule_pairs(cost, n_groups, chunks)
    cut = np.zeros((wl.n_pairs, n_bands + 1), dtype=np.int32)
    tr_off = np.ascontiguousarray(wl.tr_off, dtype=np.int64); nt = np.ascontiguousarray(wl.k3_tr_size, dtype=np.int32)
    npt = np.ascontiguousarray(wl.k3_pt_size, dtype=np.int32)
    for g in range(n_groups):
        mask = np.ascontiguousarray(group_of_pair == g, dtype=np.uint8)
        edges = np.zeros(max(n_bands - 1, 1))
        cg = np.zeros_like(cut)
        h = _lib.host()
        if h.song_host_band_edges(wl.n_pairs, _lib.iptr(nt), _lib.lptr(tr_off), _lib.dptr(wl.k3_tr), _lib.iptr(npt),
                                  mask.ctypes.data_as(_lib.C.c_void_p), float(wl.cosmo["tau0"]), int(wl.layout.l.max()), rows_per_k(wl),
                                  n_bands, _lib.dptr(edges)) or \
           h.song_host_band_cuts(wl.n_pairs, _lib.iptr(nt), _lib.lptr(tr_off), _lib.dptr(wl.k3_tr), n_bands, _lib.dptr(edges), _lib.iptr(cg)):
            raise _lib.SongLosError("song_host_band_edges / song_host_band_cuts: bad arguments")
        cut[mask != 0] = cg[mask != 0]
    return n_groups, n_bands, group_of_pair, cut


def shard_units(wl, rank: int, world: int, n_bands: int | None = None, schedule=None):
    """The part of `wl` GPU `rank` computes under the (pair groups x k3 bands) schedule: (sub-workload, pair indices in
    wl, first, last) -- sub is wl.select(pairs) when the run has a single band, else the slices [first, last) of the
    pairs' fine grids (pairs the band does not touch are left out)."""
    n_groups, nb, group_of_pair, cut = schedule if schedule is not None else schedule_bands(wl, world, n_bands)
    g, r = rank // nb, rank % nb
    own = np.nonzero(group_of_pair == g)[0]
    first, last = cut[own, r].astype(np.int64), cut[own, r + 1].astype(np.int64)
    keep = last > first
    own, first, last = own[keep], first[keep], last[keep]
    if nb == 1:
        return wl.select(own), own, first, last
    return wl.select_units(own, first, last), own, first, last


def shard_pairs(wl, rank: int, world: int, chunks_per_rank: int = 3) -> np.ndarray:
    """Indices (ascending, i.e. reference loop order) of the pairs rank `rank` owns: a few contiguous runs."""
    dev, _ = schedule_pairs(pair_cost(wl), world, chunks_per_rank if world > 1 else 1)
    return np.nonzero(dev == rank)[0]


def gather_transfers(local: np.ndarray, wl, owned: np.ndarray, rank: int, world: int, dist=None, device="cpu"):
    """Gather the [tt2_size][units of the owned pairs] arrays of all ranks on rank 0 and assemble
    ptr2->transfer as [tt2_size][n_units] in the reference pair order.  Returns None on other ranks."""
    import torch
    tt2 = wl.layout.tt2_size
    if world == 1:
        full = np.empty((tt2, wl.n_units))
        _scatter_into(full, local, wl, owned)
        return full
    sizes = [int(wl.k3_tr_size[shard_pairs(wl, r, world)].sum()) * tt2 for r in range(world)]   # known to every rank
    n_max = max(sizes)
    buf = torch.zeros(n_max, dtype=torch.float64, device=device)                 # ragged -> padded (no gatherv)
    buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1)).to(device)
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, bufs, dst=0)
    if rank != 0:
        return None
    full = np.empty((tt2, wl.n_units))
    for r in range(world):
        own_r = shard_pairs(wl, r, world)
        part = bufs[r][:sizes[r]].cpu().numpy().reshape(tt2, -1)
        _scatter_into(full, part, wl, own_r)
    return full


def gather_band_transfers(local: np.ndarray, wl, rank: int, world: int, schedule, dist=None, device="cpu"):
    """gather_transfers for the (pair groups x k3 bands) schedule: `local` is [tt2_size][units of this rank's slices] as
    shard_units lays them out; rank 0 assembles ptr2->transfer as [tt2_size][n_units] in the reference order."""
    import torch
    tt2 = wl.layout.tt2_size
    parts = [shard_units_index(wl, r, world, schedule) for r in range(world)]       # known to every rank
    sizes = [int((last - first).sum()) * tt2 for _, first, last in parts]
    n_max = max(sizes)
    buf = torch.zeros(n_max, dtype=torch.float64, device=device)
    buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1)).to(device)
    if world > 1:
        bufs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
        dist.gather(buf, bufs, dst=0)
    else:
        bufs = [buf]
    if rank != 0:
        return None
    full = np.full((tt2, wl.n_units), np.nan)
    for r in range(world):
        own, first, last = parts[r]
        part = bufs[r][:sizes[r]].cpu().numpy().reshape(tt2, -1)
        o = 0
        for ip, a, b in zip(own, first, last):
            t = int(wl.tr_off[ip])
            full[:, t + a:t + b] = part[:, o:o + (b - a)]
            o += b - a
    return full


def shard_units_index(wl, rank: int, world: int, schedule):
    """(pair indices, first, last) of shard_units without building the sub-workload."""
    n_groups, nb, group_of_pair, cut = schedule
    g, r = rank // nb, rank % nb
    own = np.nonzero(group_of_pair == g)[0]
    first, last = cut[own, r].astype(np.int64), cut[own, r + 1].astype(np.int64)
    keep = last > first
    return own[keep], first[keep], last[keep]


def _scatter_into(full, part, wl, owned):
    o = 0
    for ip in owned:
        n = int(wl.k3_tr_size[ip]); a = int(wl.tr_off[ip])
        full[:, a:a + n] = part[:, o:o + n]
        o += n
