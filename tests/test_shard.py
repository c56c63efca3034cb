"""Multi-GPU host logic on CPU: the LPT partition of the (k1,k2) pairs (the C function the drop-in
uses) and the gather of the per-rank transfer arrays over torch.distributed (gloo, world_size 2).
Each rank runs the CPU oracle on its shard here -- the test is about partition + gather order, the
GPU kernels are covered by the -m gpu tests."""
import os
import sys

import numpy as np
import pytest

from song_b200 import configs, shard, workload

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_schedule_properties():
    rng = np.random.default_rng(3)
    for n, nd in ((1, 1), (7, 3), (120, 8), (5000, 8), (3, 8)):
        cost = rng.integers(1, 10_000, size=n)
        dev, load = shard.lpt_schedule(cost, nd)
        assert dev.min() >= 0 and dev.max() < nd
        assert np.array_equal(load, np.bincount(dev, weights=cost, minlength=nd).astype(np.int64))
        # Graham's bound for LPT: makespan <= (4/3 - 1/(3m)) OPT, and OPT >= max(mean load, largest item)
        opt_lb = max(cost.sum() / nd, cost.max())
        assert load.max() <= (4 / 3 - 1 / (3 * nd)) * opt_lb + cost.max() * 1e-9 + (0 if n > nd else cost.max())
        # deterministic, and every rank gets work when there is enough of it
        dev2, _ = shard.lpt_schedule(cost, nd)
        assert np.array_equal(dev, dev2)
        if n >= nd:
            assert np.unique(dev).size == nd


def test_pair_schedule_properties():
    """song_host_schedule_pairs: contiguous chunks of about equal cost, LPT over the chunks."""
    rng = np.random.default_rng(11)
    for n, nd, cpd in ((1, 1, 1), (7, 3, 2), (253, 8, 3), (7260, 8, 3), (7260, 8, 1), (3, 8, 3), (40, 4, 16)):
        cost = rng.integers(1, 2000, size=n)
        dev, load = shard.schedule_pairs(cost, nd, cpd)
        assert dev.min() >= 0 and dev.max() < nd
        assert np.array_equal(load, np.bincount(dev, weights=cost, minlength=nd).astype(np.int64))
        runs = 1 + int(np.count_nonzero(np.diff(dev)))          # contiguous runs of one rank
        assert runs <= min(n, nd * cpd)
        if n >= nd * cpd * 20:
            assert load.max() / load.mean() < 1.05
        dev2, _ = shard.schedule_pairs(cost, nd, cpd)
        assert np.array_equal(dev, dev2)


def test_lpt_balance_on_baseline_workload():
    """configs[1] split over 8 GPUs by pair: the north_star asks for >= 85 % scaling efficiency; the static
    schedule itself must not cost more than a per cent."""
    wl = workload.build_grids(configs.get("sn_t_10percent_lmax2000"))
    dev, load = shard.lpt_schedule(shard.pair_cost(wl), 8)
    assert load.max() / load.mean() < 1.01
    dev, load = shard.schedule_pairs(shard.pair_cost(wl), 8, 3)     # the schedule transfer2_init and bench.py use
    assert load.max() / load.mean() < 1.01
    own = [shard.shard_pairs(wl, r, 8) for r in range(8)]
    assert sum(o.size for o in own) == wl.n_pairs and np.array_equal(np.sort(np.concatenate(own)), np.arange(wl.n_pairs))


def test_band_schedule_properties():
    """(pair groups x k3 bands): every fine-grid value belongs to exactly one GPU, bands are k3 intervals, the slices of a
    pair tile its grid, the cost is balanced, and the band count follows the density rule of song_host_band_count."""
    wl = workload.build_grids(configs.get("sn_t_10percent_lmax2000"))
    assert shard.band_count(wl, 1) == 1
    assert shard.band_count(wl, 8, dense_per_row=1) == 1            # pairs alone are dense enough
    assert shard.band_count(wl, 8, dense_per_row=10 ** 9) == 8      # never more bands than GPUs
    for n_dev, nb in ((2, 2), (4, 2), (8, 4), (8, 8), (6, 3)):
        sched = shard.schedule_bands(wl, n_dev, n_bands=nb)
        n_groups, n_bands, group_of_pair, cut = sched
        assert n_groups * n_bands == n_dev and n_bands == nb
        assert np.all(cut[:, 0] == 0) and np.all(cut[:, -1] == wl.k3_tr_size) and np.all(np.diff(cut, axis=1) >= 0)
        seen = np.zeros(wl.n_units, dtype=np.int32)
        cost = np.zeros(n_dev)
        for rank in range(n_dev):
            own, first, last = shard.shard_units_index(wl, rank, n_dev, sched)
            kmin, kmax = np.inf, -np.inf
            for ip, a, b in zip(own, first, last):
                t = int(wl.tr_off[ip])
                seen[t + a:t + b] += 1
                k = wl.k3_tr[t + a:t + b]
                kmin, kmax = min(kmin, k[0]), max(kmax, k[-1])
            cost[rank] = shard.predicted_ms(wl, own, first, last)     # numpy restatement of the C cost model
            if rank % n_bands:                                       # bands of a group do not overlap in k3
                assert kmin >= prev_kmax
            prev_kmax = kmax
        assert np.all(seen == 1)
        for g in range(n_groups):                                    # the bands of a group take the same predicted time
            c = cost[g * n_bands:(g + 1) * n_bands]
            assert c.max() / c.mean() < 1.02, (n_dev, nb, c)
        assert cost.max() / cost.mean() < 1.06, (n_dev, nb, cost)    # and so do the groups, up to their density difference


@pytest.mark.parametrize("name,n_dev,expect", [("sn_pol_5percent_lmax3000", 2, (2, 1)), ("sn_pol_5percent_lmax3000", 4, (2, 2)),
                                               ("sn_pol_5percent_lmax3000", 8, (2, 4)), ("sn_pol_5percent_lmax4000", 8, (2, 4))])
def test_default_schedules_of_the_scaling_runs(name, n_dev, expect):
    """The schedules `bench.py --gpus N` runs by default (configs[3] at N = 2, 4, 8; configs[4] at N = 8), at full size:
    groups x bands as DESIGN.md states, every fine-grid value on exactly one GPU, no idle GPU, predicted times balanced."""
    wl = workload.build_grids(configs.get(name))
    sched = shard.schedule_bands(wl, n_dev)
    assert (sched[0], sched[1]) == expect
    seen = np.zeros(wl.n_units, dtype=np.int8)
    ms = []
    for rank in range(n_dev):
        own, first, last = shard.shard_units_index(wl, rank, n_dev, sched)
        assert own.size > 0 and int(np.sum(np.asarray(last) - np.asarray(first))) > 0
        for ip, a, b in zip(own, first, last):
            seen[int(wl.tr_off[ip]) + int(a):int(wl.tr_off[ip]) + int(b)] += 1
        ms.append(shard.predicted_ms(wl, own, first, last))
    assert np.all(seen == 1)
    assert max(ms) / (sum(ms) / len(ms)) < 1.04, ms


def test_select_units_metadata():
    """Slices of fine grids: whole coarse grids and sources, cut fine grids, physical range relative to the slice and the
    k3 of the first / last physical point carried along (include/song_los.h, song_los_batch)."""
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")))
    sched = shard.schedule_bands(wl, 3, n_bands=3)
    total = 0
    for rank in range(3):
        sub, own, first, last = shard.shard_units(wl, rank, 3, schedule=sched)
        total += sub.n_units
        assert np.array_equal(sub.k3_pt_size, wl.k3_pt_size[own]) and sub.src.size == sub.src_off[-1] + sub.k3_pt_size[-1] * wl.layout.tp2_size * wl.tau.size
        for q, (ip, a, b) in enumerate(zip(own, first, last)):
            t = int(wl.tr_off[ip])
            assert np.array_equal(sub.k3_tr[sub.tr_off[q]:sub.tr_off[q] + sub.k3_tr_size[q]], wl.k3_tr[t + a:t + b])
            assert sub.phys_start[q] == wl.phys_start[ip] - a and sub.phys_size[q] == wl.phys_size[ip]
            assert sub.phys_k_first[q] == wl.k3_tr[t + wl.phys_start[ip]]
            assert sub.phys_k_last[q] == wl.k3_tr[t + wl.phys_start[ip] + wl.phys_size[ip] - 1]
    assert total == wl.n_units


def _band_worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl)
    proj = O.build_projection(wl)
    sched = shard.schedule_bands(wl, world, n_bands=world)
    own, first, last = shard.shard_units_index(wl, rank, world, sched)
    whole, _ = O.OracleRun(wl.select(own), proj).run()              # the oracle works on whole pairs: cut its result
    sub = wl.select(own)
    local = np.concatenate([whole[:, int(sub.tr_off[q]) + a:int(sub.tr_off[q]) + b] for q, (a, b) in enumerate(zip(first, last))], axis=1)
    full = shard.gather_band_transfers(local, wl, rank, world, sched, dist=dist)
    if rank == 0:
        ref, _ = O.OracleRun(wl, proj).run()
        np.save(out_path, np.stack([full, ref]))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_band_gather(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered_bands.npy")
    port = 29800 + (os.getpid() % 200)
    mp.spawn(_band_worker, args=(2, port, out), nprocs=2, join=True)
    full, ref = np.load(out)
    assert np.array_equal(full, ref)          # band slices are disjoint, complete and land in the reference order


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import oracle_lib as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl)
    proj = O.build_projection(wl)
    owned = shard.shard_pairs(wl, rank, world)
    sub = wl.select(owned)
    local, _ = O.OracleRun(sub, proj).run()
    full = shard.gather_transfers(local, wl, owned, rank, world, dist=dist)
    if rank == 0:
        ref, _ = O.OracleRun(wl, proj).run()
        np.save(out_path, np.stack([full, ref]))
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "gathered.npy")
    port = 29600 + (os.getpid() % 200)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    full, ref = np.load(out)
    assert np.array_equal(full, ref)          # shards are disjoint, complete and land in the reference order


def test_oracle_view_of_slices_equals_whole_pairs_cut():
    """bench.py checks band shards against the oracle on a VIEW of the slices (slice + the physical end its wings copy):
    the view must give what the oracle gives on the whole pairs, cut -- bit for bit."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import bench
    import oracle_lib as O
    wl = workload.fill_sources(workload.build_grids(configs.tiny(sources_k3_interpolation="cubic")), kind="noise")
    proj = O.build_projection(wl)
    ref, _ = O.OracleRun(wl, proj).run()
    sched = shard.schedule_bands(wl, 4, n_bands=4)
    cases = {"left": 0, "right": 0, "inside": 0}
    for rank in range(4):
        sub, own, first, last = shard.shard_units(wl, rank, 4, schedule=sched)
        idx = np.arange(sub.n_pairs)
        view, keep = bench.oracle_view_of_slices(sub, idx)
        res, _ = O.OracleRun(view, proj).run()
        for q, (lo, hi) in enumerate(keep):
            ip, a, b = int(own[q]), int(first[q]), int(last[q])
            got = res[:, int(view.tr_off[q]) + lo:int(view.tr_off[q]) + hi]
            assert np.array_equal(got, ref[:, int(wl.tr_off[ip]) + a:int(wl.tr_off[ip]) + b])
            cases["left" if view.k3_tr_size[q] > b - a and lo == 0 else "right" if lo == 1 else "inside"] += 1
    assert cases["left"] > 0 and cases["right"] > 0 and cases["inside"] > 0, cases
