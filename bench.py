#!/usr/bin/env python
"""bench.py -- LOS transfer integrals per second of the transfer2 stage on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (k3 spline, k3 sort, source interpolation, DMMA LOS kernel) over ONE
workload of synthetic input: BASELINE.json configs[3] (ini/intrinsic.ini + pre/sn_pol_5percent_lmax3000.pre, T+E,
2.5e9 integrals), the largest configuration that fits one B200.  With N ranks the SAME workload is split with the
static schedule of song_b200/shard.py (C: los_host.c) into (k1,k2)-pair groups x k3 bands: pairs as transfer2_init
splits them (song_host_schedule_pairs), and -- when pairs alone would leave fewer than 80 k3 values per projection-table
row on a GPU, which is what the tensor-core kernel's efficiency follows -- bands of the fine k3 axis
(song_host_band_edges; N=4: 2 x 2, N=8: 2 x 4 at configs[3]).  Strong scaling, no data-path collective,
value = integrals of the workload / max-over-ranks time; the per-GPU results are then gathered on rank 0 (NCCL, timed
separately) and spot-checked against the CPU oracle in the reference pair order.

`value`  : inputs resident in HBM when the timed region starts (the shard's batches, one per engine slot).
`e2e`    : the same metric through the reference-facing call -- SONG's transfer2_init() (the drop-in of
           song_b200/host/transfer2_gpu.c, called on SONG structures exactly as main/song.c:106 does) with the sources
           in pageable host memory (ppt2->sources) and the result in ptr2->transfer: packing, H2D, kernels, D2H all
           inside the timed region, on all N GPUs of the box (one host thread per GPU inside the module).
`configs`: the other single-GPU BASELINE configurations measured the same way with fewer steps (N=1), and
           configs[4] (sn_pol_5percent_lmax4000) split over the ranks when N=8.  These legs run after the main workload is
           measured and under a deadline (--extra-deadline): a leg that would start late is skipped, one that is still
           running at the deadline is abandoned, and the main line goes out either way with a note in `configs`.
`--impl reference` : the reference's own source/transfer2.c, compiled where it lies into oracle/_ref/libsong_ref.so
           (oracle/Makefile target `ref`; stand-ins only for its absent CLASS fork), run through its module API
           transfer2_init on SONG structures with all host threads, on a bounded sample of the same workload
           (kind = "reference"); where that library is absent, the CPU restatement of the same loops (oracle/, kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

T_PROCESS_START = time.time()                      # the deadline of the `configs` legs counts from here (see --extra-deadline)
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

METRIC = "2nd-order LOS transfer integrals/sec (k1,k2,k3,l,m)"
UNIT = "integrals/s"
DEFAULT_WORKLOAD = "sn_pol_5percent_lmax3000"      # BASELINE.json configs[3]


# ----------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int, period_ms: int = 50):
        self.idx = device_index
        self.period_ms = period_ms
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                 "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pin this process to the CPUs NVML reports as local to its GPU, so that the page-locked staging
    buffers (allocated afterwards) live on the socket the GPU's PCIe root port hangs off.  With 8 ranks on a
    two-socket host the copies otherwise cross the inter-socket link.  Best effort: reports what it did."""
    info = {"bound": False}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(local_rank)
        handle = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                handle = None
        if handle is None:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
            handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            info = {"bound": True, "n_cpus": len(cpus), "first_cpu": cpus[0], "last_cpu": cpus[-1]}
    except Exception as exc:                          # no NVML, restricted cpuset, ...: run unbound
        info = {"bound": False, "why": str(exc)[:120]}
    return info


# ----------------------------------------------------------------------------------------------
# work counters
# ----------------------------------------------------------------------------------------------
def count_terms(wl, xmin, xsize) -> int:
    """Exact number of (tau, L) terms that pass the reference's skip tests (transfer2.c:2217,2232),
    summed over every stored integral: sum over (k3, tau) of #{(tt, L): index_x >= index_xmin_J}."""
    lay = wl.layout
    nx = wl.xx.size
    hist = np.zeros(nx + 1, dtype=np.int64)
    for tt in range(lay.tt2_size):
        il, im = int(lay.tt_index_l[tt]), int(lay.tt_index_m[tt])
        segs = []
        if lay.has_source_T and lay.index_tt2_T <= tt < lay.index_tt2_T + lay.n_transfers:
            segs.append((lay.index_J_TT, lay.l_max_los_t))
        elif lay.has_source_E and lay.index_tt2_E <= tt < lay.index_tt2_E + lay.n_transfers:
            segs.append((lay.index_J_EE, lay.l_max_los_p))
            if lay.m[im] != 0:
                segs.append((lay.index_J_EB, lay.l_max_los_p))
        else:
            segs.append((lay.index_J_EE, lay.l_max_los_p)); segs.append((lay.index_J_EB, lay.l_max_los_p))
        for J, lmax in segs:
            for L in range(lmax + 1):
                if lay.m[im] > min(L, int(lay.l[il])):
                    continue
                jidx = ((J * lay.L_size + L) * lay.l_size + il) * lay.m_size + im
                if xsize[jidx] <= 0:
                    continue
                hist[int(xmin[jidx])] += 1
    cnt = np.cumsum(hist)            # cnt[ix] = number of (tt,L) with xmin <= ix
    d = wl.cosmo["tau0"] - wl.tau
    total = 0
    step = 200_000
    for s in range(0, wl.k3_tr.size, step):
        x = wl.k3_tr[s:s + step, None] * d[None, :]
        ix = (x / wl.xx_step).astype(np.int64)
        total += int(cnt[np.minimum(ix, nx)].sum())
    return total


# ----------------------------------------------------------------------------------------------
# the reference arm: CPU restatement of transfer2 (oracle/) on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample(name: str, target_seconds: float, max_pairs=None):
    """(sub-workload, projection) sized for about `target_seconds` of oracle time per pass."""
    import oracle_lib as O
    from song_b200 import configs, workload
    p = configs.get(name)
    grids = workload.build_grids(p, max_pairs=max_pairs)
    # calibrate on a thin slice, then size the sample
    probe_idx = np.unique(np.linspace(0, grids.n_pairs - 1, 24).astype(int))
    probe = workload.fill_sources(grids.select(probe_idx))
    proj = O.build_projection(grids)
    run = O.OracleRun(probe, proj)
    t = time.perf_counter(); run.run(); dt = time.perf_counter() - t
    rate = probe.n_integrals / max(dt, 1e-6)
    want = rate * target_seconds
    stride = max(1, int(round(grids.n_integrals / max(want, 1.0))))
    idx = np.arange(0, grids.n_pairs, stride)
    sub = workload.fill_sources(grids.select(idx))
    return grids, sub, proj, stride


def reference_triangle(name: str, n_k: int, max_pairs=None):
    """The workload restricted to `n_k` of its k values: the reference's transfer2_init loops over the whole (k1,k2)
    triangle of ppt2->k, so a bounded sample of a run is the whole triangle of a thinned k list (every pair of it is a pair
    of the full run, with the same coarse and fine k3 grids).  The k values kept are the midpoints of n_k equal strata of
    the k list, so that the kept pairs form a regular lattice over the full triangle and cheap low-k and expensive high-k
    pairs appear in the run's own proportions (keeping both ends of the list instead over-weights the most expensive
    row: 1.15e5 instead of ~4e5 integrals/s at configs[3])."""
    from song_b200 import configs, workload
    p = configs.get(name)
    k_full = workload.build_grids(p, max_pairs=1).k
    n_k = max(2, min(n_k, k_full.size))
    idx = np.unique(np.floor((np.arange(n_k) + 0.5) * k_full.size / n_k).astype(int))
    return workload.build_grids(dict(p, k=k_full[idx]), max_pairs=max_pairs), int(k_full.size)


class c_stdout_to_stderr:
    """The reference prints to the C stdout (e.g. its warning about fewer than 4 k values, transfer2.c:120-123); the one JSON
    line of this script must stay alone there: file descriptor 1 points to stderr while the library runs."""

    def __enter__(self):
        import ctypes
        self.libc = ctypes.CDLL(None)
        sys.stdout.flush(); self.libc.fflush(None)
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        self.libc.fflush(None)
        os.dup2(self.saved, 1); os.close(self.saved)
        return False


def time_reference_library(name: str, target_seconds: float, steps: int, warmup: int):
    with c_stdout_to_stderr():
        return _time_reference_library(name, target_seconds, steps, warmup)


def _time_reference_library(name: str, target_seconds: float, steps: int, warmup: int):
    """oracle/_ref/libsong_ref.so -- the reference's own source/transfer2.c, compiled where it lies (oracle/Makefile) --
    timed on a bounded sample of workload `name`: transfer2_init on SONG structures, all host threads (the reference
    parallelises the k3 loop of every pair with OpenMP, transfer2.c:450).  Returns None when the library is not there."""
    import dropin_lib
    import oracle_lib as O
    from song_b200 import workload
    if not os.path.exists(dropin_lib.REF_PATH):
        return None
    probe, k_size = reference_triangle(name, 3)
    workload.fill_sources(probe)
    proj = O.build_projection(probe)                   # the tables depend on the multipole lists and the x grid only
    with dropin_lib.Harness(probe, proj, lib_path=dropin_lib.REF_PATH, keep_sources=True) as h:
        t = time.perf_counter(); h.transfer2_init(); dt = time.perf_counter() - t
        got, count, _ = h.read()
        h.transfer2_free()
    ref, _ = O.OracleRun(probe, proj).run()
    same = bool(np.array_equal(got, ref))              # the arm being timed computes what the oracle computes, bit for bit
    # a triangle of n_k of the k_size values holds about (n_k / k_size)^2 of the run's integrals
    from song_b200 import configs
    full_integrals = workload.build_grids(configs.get(name)).n_integrals
    rate = probe.n_integrals / max(dt, 1e-9)
    n_k = int(max(2, min(k_size, round(k_size * (target_seconds * rate / full_integrals) ** 0.5))))   # 2 values = 3 pairs at least
    sub, _ = reference_triangle(name, n_k)
    workload.fill_sources(sub)
    each = []
    with dropin_lib.Harness(sub, proj, lib_path=dropin_lib.REF_PATH, keep_sources=True) as h:
        for i in range(warmup + steps):
            t = time.perf_counter(); h.transfer2_init(); dt = time.perf_counter() - t
            if i >= warmup:
                each.append(dt)
            h.transfer2_free(); h.fill_sources()
    dt = sum(each) / max(len(each), 1)
    return {"value": sub.n_integrals / dt, "seconds_per_step": dt, "integrals": sub.n_integrals, "pairs": sub.n_pairs,
            "k_values": int(sub.k.size), "k_size_full": k_size, "probe_bitwise_equal_to_oracle": same,
            "cores": O.lib().oracle_num_threads()}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 when it is unset; the CPU arm uses every host thread
    if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle_lib as O
    from song_b200 import configs, workload
    set_omp_threads(os.cpu_count() or 1)
    try:
        # a step is a bounded sample: about --cpu-seconds of host time, less when K + W calls would not end within a few minutes
        per_step = max(1.0, min(args.cpu_seconds, 240.0 / max(args.steps + args.warmup, 1)))
        lib_run = time_reference_library(args.workload, per_step, max(args.steps, 1), args.warmup)
    except Exception as exc:                         # the oracle port below is always there
        sys.stderr.write(f"bench.py: the reference library could not be timed ({str(exc)[:200]}); using the oracle port\n")
        lib_run = None
    if lib_run is not None:
        grids = workload.build_grids(configs.get(args.workload))
        value = lib_run["value"]
        sample = (f"transfer2_init of the reference on the whole (k1,k2) triangle of {lib_run['k_values']} of the {lib_run['k_size_full']} k values of "
                  f"{args.workload} (midpoints of equal strata of the k list): {lib_run['pairs']} pairs, {lib_run['integrals']} of {grids.n_integrals} integrals per step")
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": lib_run["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload.describe(grids), workload=args.workload,
                           note="the reference's own source/transfer2.c (coccoinomane/song @ 2ea44ecf), compiled where it lies into oracle/_ref/"
                                "libsong_ref.so with stand-ins for the declarations and helpers of its absent CLASS fork (oracle/ref_build/), "
                                "run through its module API transfer2_init on SONG structures; synthetic inputs from the same generator as "
                                "the GPU arm; projection tables from the oracle's bessel2 restatement; the GPU is idle"),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": lib_run["cores"], "cpu_model": cpu_model(), "kind": "reference", "sample": sample,
                             "probe_bitwise_equal_to_oracle": lib_run["probe_bitwise_equal_to_oracle"],
                             "build": "gcc -O3 -g -DDEBUG -std=c99 -w (the reference Makefile's own defaults) -fopenmp (its optional parallel flags) "
                                      "-ffp-contract=off on /root/reference/source/transfer2.c, oracle/Makefile target ref"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return
    grids, sub, proj, stride = cpu_sample(args.workload, args.cpu_seconds)
    run = O.OracleRun(sub, proj)
    for _ in range(args.warmup):
        run.run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run.run()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    value = sub.n_integrals / dt
    cores = O.lib().oracle_num_threads()
    sample = (f"every {stride}-th (k1,k2) pair of {args.workload} in the reference loop order: {sub.n_pairs} of "
              f"{grids.n_pairs} pairs, {sub.n_integrals} of {grids.n_integrals} integrals per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload.describe(grids), workload=args.workload,
                       note="CPU restatement of reference transfer2 (reference not buildable here: class.git is an empty submodule); "
                            "the synthetic inputs (grids, sources) come from the same generator as the GPU arm "
                            "(song_b200.workload, libsong_host.so for the k3 grids) -- the timed region is the oracle alone, the GPU is idle"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "cpu_model": cpu_model(), "kind": "port", "sample": sample,
                         "build": "gcc -O3 -std=c99 -fopenmp -ffp-contract=off, reference loop structure (serial pair loop, omp for over k3); "
                                  "no -DDEBUG variant: the restatement does not carry the reference's debug-only checks"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing of one bench process (a no-op at world size 1)."""

    def __init__(self, rank, world, local_rank):
        import torch
        self.torch = torch
        self.rank, self.world, self.local_rank = rank, world, local_rank
        self.dist = None
        if world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            import datetime
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(minutes=60))
            self.dist = dist
            self.cpu_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=60))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        """Wait on the host only: the GPUs of the waiting ranks stay free (rank 0 drives all of them through transfer2_init)."""
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier(group=self.cpu_group)

    def reduce(self, x: float, op: str = "max") -> float:
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return float(t.item())

    def all_gather_floats(self, x: float):
        if self.dist is None:
            return [x]
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


def split_bounds(k3_tr_size, cap_units: int, n_slots: int = 3):
    """Pair indices at which a shard (pairs in reference order) is cut into at most n_slots batches of at most cap_units k3 values."""
    cum = np.cumsum(np.asarray(k3_tr_size, dtype=np.int64))
    n_b = max(1, -(-int(cum[-1]) // max(cap_units, 1)))
    if n_b > n_slots:
        raise SystemExit(f"bench.py: the shard has {int(cum[-1])} k3 values, the GPU holds {n_slots} batches of {cap_units}")
    cuts = [0]
    for i in range(1, n_b):
        c = int(np.searchsorted(cum, cum[-1] * i / n_b)) + 1
        if c > cuts[-1]:
            cuts.append(c)
    if cuts[-1] < cum.size:
        cuts.append(int(cum.size))
    return cuts


def split_resident(sub, cap_units: int, n_slots: int = 3):
    cuts = split_bounds(sub.k3_tr_size, cap_units, n_slots)
    if len(cuts) == 2:
        return [sub]
    return [sub.select(np.arange(cuts[i], cuts[i + 1])) for i in range(len(cuts) - 1)]


def set_omp_threads(n: int):
    """Team size of the OpenMP runtime the oracle and the host libraries share in this process (torchrun starts its
    workers with OMP_NUM_THREADS=1; the runtime was loaded with that value)."""
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def oracle_view_of_slices(sub, idx):
    """A workload the oracle can run as it is for pairs `idx` of a sliced workload (k3 bands), and per pair the columns of
    its result that belong to the slice.  The oracle, like the reference, takes the fine k3 grid of a pair as given and
    evaluates every physical point on its own; only the wings of the flat extrapolation (transfer2.c:2743-2757) look at
    another point, the first / last PHYSICAL one.  The view is therefore the slice itself, plus that one point when the
    slice holds wing values but not the physical end they copy (it then sorts before / after the whole slice)."""
    par = sub.parent
    whole = par.select(sub.parent_pair[idx])          # coarse grids + sources of the whole pairs; fine grids replaced below
    grids, starts, sizes, keep = [], [], [], []
    for q, i in enumerate(idx):
        ip = int(sub.parent_pair[i]); a = int(sub.slice_start[i]); n = int(sub.k3_tr_size[i]); b = a + n
        K = par.k3_tr[int(par.tr_off[ip]):int(par.tr_off[ip]) + int(par.k3_tr_size[ip])]
        f = int(par.phys_start[ip]); l = f + int(par.phys_size[ip]) - 1
        if f >= b:                                    # the whole slice is left wing: copies the first physical point
            grids.append(np.concatenate([K[a:b], K[f:f + 1]])); starts.append(n); sizes.append(1); keep.append((0, n))
        elif l < a:                                   # the whole slice is right wing: copies the last physical point
            grids.append(np.concatenate([K[l:l + 1], K[a:b]])); starts.append(0); sizes.append(1); keep.append((1, n + 1))
        else:                                         # the slice holds the physical ends its wings copy
            grids.append(K[a:b]); starts.append(max(f, a) - a); sizes.append(min(l, b - 1) - max(f, a) + 1); keep.append((0, n))
    nt = np.array([g.size for g in grids], dtype=np.int32)
    whole.k3_tr_size = nt
    whole.tr_off = np.concatenate([[0], np.cumsum(nt)[:-1]]).astype(np.int64)
    whole.k3_tr = np.ascontiguousarray(np.concatenate(grids))
    whole.phys_start = np.array(starts, dtype=np.int32); whole.phys_size = np.array(sizes, dtype=np.int32)
    return whole, keep


def spot_check(O, sub, got_of_pair, rows, xmin, xsize, n_check: int, run_scale=None):
    """Oracle on a few pairs of `sub`; got_of_pair(i) -> [tt2][k3 of pair i].  Max relative error above 1e-6 of the
    per-(l,m) maximum (BASELINE.json criterion; the maximum is taken over the checked pairs unless run_scale is given).
    When `sub` holds slices of fine grids (k3 bands) the oracle runs on a view of the slices (oracle_view_of_slices)."""
    idx = np.unique(np.linspace(0, sub.n_pairs - 1, min(n_check, sub.n_pairs)).astype(int))
    oproj = O.OracleProjection(rows, xmin, xsize, None)
    if sub.parent is not None:
        view, keep = oracle_view_of_slices(sub, idx)
        res, _ = O.OracleRun(view, oproj).run()
        ref = np.concatenate([res[:, int(view.tr_off[q]) + lo:int(view.tr_off[q]) + hi] for q, (lo, hi) in enumerate(keep)], axis=1)
    else:
        ref, _ = O.OracleRun(sub.select(idx), oproj).run()
    got = np.concatenate([got_of_pair(int(i)) for i in idx], axis=1)
    scale = run_scale if run_scale is not None else np.abs(ref).max(axis=1, keepdims=True)
    m = np.abs(ref) > 1e-6 * scale
    err = float(np.max(np.abs(got - ref)[m] / np.abs(ref)[m])) if m.any() else 0.0
    return {"pairs_checked": int(idx.size), "max_rel_err_above_1e-6_of_max": err, "tolerance": 1e-9, "ok": bool(err <= 1e-9)}


def measure_workload(name: str, args, D: Dist, steps: int, warmup: int, e2e_steps: int, with_cpu: bool, sampler=None):
    """One BASELINE workload on the ranks of this run.  Returns the result dict on rank 0 (None elsewhere)."""
    import torch
    import dropin_lib
    from song_b200 import _lib, configs, shard, workload
    from song_b200.bessel2 import Bessel2
    from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine

    rank, world = D.rank, D.world
    mode = MODE_FAST if args.mode == "fast" else MODE_EXACT
    t_setup = time.perf_counter()
    p = configs.get(name)
    grids = workload.build_grids(p, max_pairs=args.max_pairs)
    full_pairs = grids.n_pairs == grids.k.size * (grids.k.size + 1) // 2
    # static schedule of the run: pair groups x k3 bands (song_b200/shard.py, los_host.h); GPU = group * n_bands + band
    sched = shard.schedule_bands(grids, world, n_bands=(args.bands or None), chunks_per_group=(args.chunks_per_rank or None))
    n_groups, n_bands, group_of_pair, band_cut = sched
    chunks = (args.chunks_per_rank or shard.chunks_per_device(grids, n_groups)) if n_groups > 1 else 1
    owned, own_first, own_last = shard.shard_units_index(grids, rank, world, sched)
    load = np.array([shard.predicted_ms(grids, *shard.shard_units_index(grids, r, world, sched)) for r in range(world)])   # model of los_host.c
    want_e2e = e2e_steps > 0 and full_pairs and os.path.exists(dropin_lib.PATH)
    if rank == 0 and want_e2e:
        workload.fill_sources(grids)                 # rank 0 also needs the whole run for transfer2_init
        whole = grids.select(owned) if world > 1 else grids
    else:
        whole = workload.fill_sources(grids.select(owned))
    # the pairs this GPU touches (whole: coarse grids and sources) and, with several bands, the slices of their fine grids
    sub = whole if n_bands == 1 else whole.select_units(np.arange(whole.n_pairs), own_first, own_last)
    eng = LosEngine(sub, device=D.local_rank)
    b2 = Bessel2(sub).init(eng)                      # bessel2_init: host j_l1 + 3j, device J_Llm assembly
    # resident batches of this GPU: as few as fit (1, 2 or 3 engine slots), the same number and bound on every rank
    units_max = int(D.reduce(float(sub.n_units), "max"))
    for n_slots in (1, 2, 3):
        cap = int(D.reduce(float(eng.max_batch_units(mode, sub.src.size / max(sub.n_units, 1), n_slots)), "min"))
        if -(-units_max // n_slots) <= int(0.88 * cap) or n_slots == 3:        # slack: batches are cut at pair boundaries, and the
                                                                                # parity / gather buffers of this script need room too
            break
    if int(D.reduce(float(-(-sub.n_units // max(cap, 1))), "max")) > 3:      # decided together: no rank may leave the others in a collective
        eng.close()
        raise RuntimeError(f"{name}: a GPU's shard does not fit three resident batches of {cap} k3 values on {world} GPU(s)")
    batches = split_resident(sub, cap)
    peaks = eng.measure_peaks()
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    for i, b in enumerate(batches):                  # inputs resident in HBM: one slot per batch
        eng.select_slot(i)
        eng.upload(b)
    t_setup = time.perf_counter() - t_setup

    # ---- value: inputs resident in HBM ----
    def one_step(acc=None):
        n = 0
        for i in range(len(batches)):
            eng.select_slot(i)
            tm = eng.compute(mode)
            n += tm["n_launches"]
            if acc is not None:
                for k in acc:
                    acc[k] += tm[k]
        return n
    for _ in range(warmup):
        one_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {"ms_integrate": 0.0, "ms_spline": 0.0, "ms_sort": 0.0, "ms_interp": 0.0, "ms_total": 0.0}
    launches = 0
    D.barrier()
    if sampler is not None:
        sampler.start()
    ev0.record(stream)
    for _ in range(steps):
        launches += one_step(kern)
    ev1.record(stream)
    D.barrier()
    clocks = sampler.stop() if sampler is not None else None
    my_ms = ev0.elapsed_time(ev1) / steps
    rank_ms = D.all_gather_floats(my_ms)
    ms_step = max(rank_ms)
    for k in kern:
        kern[k] /= steps
    n_int_rank = sub.n_integrals
    n_int_total = grids.n_integrals
    value = n_int_total / (ms_step * 1e-3)
    launches_total = int(D.reduce(float(launches), "sum"))

    # ---- parity of what was just computed + the one collective of the path: the gather of the per-GPU results ----
    import oracle_lib as O
    rows, xmin, xsize = b2.tables(eng)
    results = []
    for i, b in enumerate(batches):
        eng.select_slot(i)
        ptr, n = eng.device_result()
        res = torch.empty(n, dtype=torch.float64, device="cuda")
        eng.copy_result_to(res.data_ptr(), n)
        results.append(res)
    pair_batch = np.concatenate([np.full(b.n_pairs, i) for i, b in enumerate(batches)])
    pair_local = np.concatenate([np.arange(b.n_pairs) for b in batches])

    def local_pair(i):
        b = batches[pair_batch[i]]; j = int(pair_local[i])
        a, n = int(b.tr_off[j]), int(b.k3_tr_size[j])
        return results[pair_batch[i]].view(sub.layout.tt2_size, b.n_units)[:, a:a + n].cpu().numpy()
    parity = spot_check(O, sub, local_pair, rows, xmin, xsize, args.check_pairs)
    # ---- the k3 integral of the intrinsic bispectrum on the transfer functions still in HBM ----
    k3int = None
    if args.k3_integral and rank == 0 and n_bands > 1:
        k3int = {"skipped": "the k3 integral of a pair needs its whole fine grid (trapezoidal weights of bispectra2.c:1279-1326); with k3 bands "
                            "a GPU holds slices -- run with --bands 1 for this leg"}
    elif args.k3_integral and rank == 0:
        try:                                         # rank 0 alone is in here: an exception must not leave the other ranks in a collective
            ht = b2.host_tables()
            eng.set_bessel_table(ht["j_l1"], ht["x_min_l1"])
            lay = sub.layout
            index_l1 = {int(l1): i for i, l1 in enumerate(ht["l1"])}
            c0 = sub.cosmo["tau0"] - sub.cosmo["tau_rec"]
            r_size = {"sn_t_10percent_lmax2000": 75, "sn_pol_10percent_lmax2000": 75, "quick_song_run": 50}.get(name, 125)   # r_size of the .pre files
            r_arr = np.linspace(c0 - 400.0, c0 + 250.0, r_size)
            bases = ([lay.index_tt2_T] if lay.has_source_T else []) + ([lay.index_tt2_E] if lay.has_source_E else [])
            tts = np.asarray([bs + int(lay.lm_array[il, 0]) for bs in bases for il in range(lay.l_size)])
            rws = np.asarray([index_l1[int(l3)] for _ in bases for l3 in lay.l])
            ms_k3 = 0.0; out_bytes = 0; ok_k3 = True; n_k3 = 0
            for i, b in enumerate(batches):
                eng.select_slot(i)
                I = eng.integrate_k3(tts, rws, r_arr, b.n_pairs)            # [tt][r][pair] of this batch
                ms_k3 += eng.k3_ms; out_bytes += I.nbytes; n_k3 += I.size
                if i == 0:                                                    # bitwise against the oracle loop on the same T
                    T_dev = results[0].view(lay.tt2_size, b.n_units)
                    for q in (0, len(tts) // 2, len(tts) - 1):
                        ref = O.integral_over_k3(b, T_dev[int(tts[q])].cpu().numpy(), ht["j_l1"][int(rws[q])], None, ht["x_min_l1"][int(rws[q])], r_arr)
                        ok_k3 = ok_k3 and bool(np.array_equal(I[q], ref))
                del I
            k3int = {"kernel_ms": ms_k3, "integrals": int(n_k3), "r_size": int(r_size), "requests": int(tts.size),
                     "k3_terms": int(sub.n_units) * int(tts.size) * int(r_size),
                     "result_bytes": int(out_bytes), "transfer_bytes_replaced": int(n_int_rank * 8),
                     "bitwise_equal_to_oracle": ok_k3,
                     "api": "song_los_integrate_k3 on the resident result of every batch (bispectra2.c:1107-1439, bessel2.c:499-605); "
                            "rank 0's shard; the result a consumer copies to the host shrinks from transfer_bytes_replaced to result_bytes"}
        except Exception as exc:
            k3int = {"error": str(exc)[:400]}
    # everything below works on the copies in `results`: release the engine's device memory before the gather buffers
    rank_batches = [int(b.n_units) for b in batches]
    eng.close()
    del eng
    gather = None
    if world > 1:
        # ragged -> padded (no gatherv): one gather per resident batch slot
        n_slots = int(D.reduce(float(len(batches)), "max"))
        sizes = [int(x) for x in D.all_gather_floats(float(sum(r.numel() for r in results)))]
        n_max = int(D.reduce(float(max(r.numel() for r in results)), "max"))
        pads = []
        for i in range(n_slots):
            t = torch.zeros(n_max, dtype=torch.float64, device="cuda")
            if i < len(results):
                t[:results[i].numel()] = results[i]
            pads.append(t)
        bufs = [[torch.empty(n_max, dtype=torch.float64, device="cuda") for _ in range(world)] if rank == 0 else None for _ in range(n_slots)]
        D.dist.gather(pads[0], bufs[0], dst=0)        # untimed: the first collective builds the NCCL communicator
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        D.barrier()
        g0.record()
        for i in range(n_slots):
            D.dist.gather(pads[i], bufs[i], dst=0)
        g1.record()
        D.barrier()
        gather_ms = D.reduce(g0.elapsed_time(g1), "max")
        gather = {"ms": gather_ms, "bytes": int(sum(sizes) * 8), "gbs": sum(sizes) * 8 / 1e9 / max(gather_ms * 1e-3, 1e-9),
                  "api": "torch.distributed.gather over NCCL, one per resident batch slot, padded to the largest batch"}
        if rank == 0:
            # the gathered array in the reference pair order: check pairs owned by every rank against the oracle
            tt2 = grids.layout.tt2_size
            def gathered_pair(ip):
                parts = []                                          # the pair's slices, band by band, from the GPUs that own them
                for band in range(n_bands):
                    r = int(group_of_pair[ip]) * n_bands + band
                    own_r, first_r, last_r = shard.shard_units_index(grids, r, world, sched)
                    j = int(np.searchsorted(own_r, ip))
                    if j >= own_r.size or own_r[j] != ip:
                        continue                                    # this band does not touch the pair
                    sizes_r = (last_r - first_r).astype(np.int64)
                    cuts = split_bounds(sizes_r, cap)               # every rank cut its shard with the same rule
                    ib = int(np.searchsorted(cuts, j, side="right")) - 1
                    units_b = int(sizes_r[cuts[ib]:cuts[ib + 1]].sum())
                    a = int(sizes_r[cuts[ib]:j].sum()); n = int(sizes_r[j])
                    parts.append(bufs[ib][r][:units_b * tt2].view(tt2, units_b)[:, a:a + n].cpu().numpy())
                return np.concatenate(parts, axis=1)
            # the other ranks are idle here: the oracle gets the whole host; sources are generated for the checked pairs only
            # (the generator is deterministic per pair)
            n_chk = max(args.check_pairs, world)
            idx_chk = np.unique(np.linspace(0, grids.n_pairs - 1, min(n_chk, grids.n_pairs)).astype(int))
            chk = grids.select(idx_chk)
            if chk.src is None:
                workload.fill_sources(chk)
            set_omp_threads(os.cpu_count() or 1)
            gather["parity_of_gathered_array"] = spot_check(O, chk, lambda j: gathered_pair(int(idx_chk[j])), rows, xmin, xsize, idx_chk.size)
            set_omp_threads(max(1, (os.cpu_count() or 1) // world))
        del bufs, pads
    del results

    # ---- roofline of the dominant kernel (the DMMA LOS kernel), this rank's shard ----
    n_terms = count_terms(sub, xmin, xsize)
    flops = 5.0 * n_terms + 2.0 * n_int_rank * sub.tau.size          # F = 5 N_terms + 2 N_int n_tau
    t_kernel = kern["ms_integrate"] * 1e-3
    achieved = flops / t_kernel / 1e12
    peak = peaks["dmma_tflops"] if mode == MODE_FAST else peaks["dfma_tflops"]
    alg_bytes = 8.0 * (n_int_rank + sub.src.size + sub.k3_pt.size + sub.k3_tr.size)
    traffic = None; tinfo = {}
    tfile = os.path.join(ROOT, "profiles", "traffic_bytes.json")
    if os.path.exists(tfile):
        try:
            tinfo = json.load(open(tfile)).get(name, {})
            traffic = tinfo.get("los_tc_kernel")
        except Exception:
            traffic = None
    fracs = D.all_gather_floats(achieved / peak)
    per_rank = {k: D.all_gather_floats(float(v)) for k, v in (("ms_integrate", kern["ms_integrate"]), ("ms_interp", kern["ms_interp"]),
                                                              ("ms_spline", kern["ms_spline"]), ("k3_values", float(sub.n_units)),
                                                              ("pairs_touched", float(sub.n_pairs)))}
    roofline = {
        "bound": "tensor", "kernel": "los_tc_kernel (FP64 tensor core, DMMA m8n8k4, TMA-fed)" if mode == MODE_FAST else "los_exact_kernel",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_note": (None if traffic is None else
                         "DRAM bytes of ONE captured launch (ncu --set full, profiles/traffic_bytes.json): %s ms, one of %s launches per step; "
                         "algorithmic flops of a launch = algorithmic_flops_per_step / launches" % (tinfo.get("launch_ms", "the step's only"), tinfo.get("launches_per_step", 1))),
        "peak_source": ("DMMA m8n8k4 loop" if mode == MODE_FAST else "DFMA loop") + " measured in this run with song_los_measure_peaks "
                       "(MEASURED_PEAKS.json has no FP64 entry and the profiling guide no FP64 fallback)",
        "algorithmic_flops_per_step": flops, "n_terms": n_terms, "kernel_ms_per_step": kern["ms_integrate"],
        "kernel_share_of_step": kern["ms_integrate"] / max(kern["ms_total"], 1e-9),
        "step_frac": flops / (my_ms * 1e-3) / 1e12 / peak,
        "frac_per_rank": fracs, "per_rank": per_rank,
        "dfma_peak_tflops": peaks["dfma_tflops"], "dmma_peak_tflops": peaks["dmma_tflops"],
        "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (kern["ms_total"] * 1e-3) / 1e9,
                "peak_gbs": peaks["copy_gbs"], "frac": alg_bytes / (kern["ms_total"] * 1e-3) / 1e9 / peaks["copy_gbs"]},
        "step_breakdown_ms": kern, "resident_batches": len(batches), "batch_k3_values": [int(b.n_units) for b in batches],
    }
    torch.cuda.empty_cache()
    D.cpu_barrier()

    # ---- e2e: SONG's transfer2_init on SONG structures, all N GPUs, host buffers in and out (rank 0 drives it) ----
    e2e = None
    def run_e2e():
        try:                                         # transfer2_init drives every GPU of the box from one process: all host cores
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except Exception:
            pass
        os.environ["SONG_LOS_DEVICES"] = str(world)
        os.environ["SONG_LOS_MODE"] = args.mode
        if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
            # torchrun exports OMP_NUM_THREADS=1 to its workers; a SONG run uses the machine (the drop-in sizes its own
            # packing teams explicitly, bessel2_init and the harness use the default team)
            os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
            set_omp_threads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        h = dropin_lib.Harness(grids, None, verbose=0, keep_sources=True)      # bessel2_init (drop-in) leaves the engines on the GPUs
        if args.dropin_verbose:
            import ctypes
            dropin_lib.lib().harness_set_verbose.argtypes = [ctypes.c_void_p, ctypes.c_int]
        t_b2 = time.perf_counter() - t0
        first_ms = None
        for i in range(2):                                                       # warm: CUDA buffers reserved, page-locked pool filled
            t0 = time.perf_counter(); h.transfer2_init(); dt = time.perf_counter() - t0
            if first_ms is None:
                first_ms = dt * 1e3
            h.transfer2_free(); h.fill_sources()
        e2e_sampler = ClockSampler(D.local_rank, period_ms=1000)       # slow polling: see the note at the 50 ms sampler
        e2e_sampler.start()
        t_all = 0.0; t_read = 0.0; cs = 0.0; each = []
        for i in range(e2e_steps):
            if args.dropin_verbose and i + 1 == e2e_steps:
                sys.stdout.flush()
                dropin_lib.lib().harness_set_verbose(h.h, args.dropin_verbose)      # laps of the last call on stdout (before the JSON line)
            t0 = time.perf_counter()
            h.transfer2_init()
            t1 = time.perf_counter()
            cs = h.checksum()                                                    # the step's result, read through ptr2->transfer
            t2 = time.perf_counter()
            t_all += t1 - t0; t_read += t2 - t1; each.append((t1 - t0) * 1e3)
            if i + 1 < e2e_steps:
                h.transfer2_free(); h.fill_sources()
        e2e_clocks = e2e_sampler.stop()
        got, count, labels_ok = h.read()
        tables = h.bessel2_tables()
        run_scale = np.abs(got).max(axis=1, keepdims=True)
        e2e_parity = spot_check(O, grids, lambda i: got[:, int(grids.tr_off[i]):int(grids.tr_off[i]) + int(grids.k3_tr_size[i])],
                                tables[0], tables[1], tables[2], max(args.check_pairs, 2 * world), run_scale=run_scale)
        del got
        h.close()
        if world > 1:                                 # back to this rank's share of the host (the other ranks were idle so far)
            set_omp_threads(max(1, (os.cpu_count() or 1) // world))
        e2e_ms = t_all / e2e_steps * 1e3
        h2d = int(grids.src.nbytes + grids.k3_pt.nbytes + grids.k3_tr.nbytes + (3 * 8 + 5 * 4) * grids.n_pairs)
        d2h = int(n_int_total * 8)
        e2e = {"value": n_int_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms, "ms_each_step": each, "steps": e2e_steps, "n_gpus": world,
               "api": "transfer2_init(ppr, ppr2, pba, pth, ppt, ppt2, pbs, pbs2, ptr, ptr2) of the drop-in module on SONG structures "
                      "(tests/harness/dropin_harness.c plays main/song.c:100-106): sources in pageable ppt2->sources, result in "
                      "ptr2->transfer; packing, H2D, kernels and D2H of every batch inside the call; host clock around the call; "
                      "schedule of the drop-in: (k1,k2) pairs only (every source byte crosses PCIe once), SONG_LOS_DEVICES GPUs from one process",
               "first_call_ms": first_ms, "first_call_note": "includes page-locking the staging and the ptr2->transfer slabs (kept in a pool afterwards)",
               "bessel2_init_s": t_b2, "result_read_ms_per_step": t_read / e2e_steps * 1e3, "checksum_abs_sum": cs,
               "count_memorised_transfers": count, "labels_ok": labels_ok, "parity": e2e_parity, "clocks_gpu0_1s_period": e2e_clocks}
        return e2e

    if want_e2e and rank == 0:
        try:                                         # rank 0 alone is in here: an exception must not leave the other ranks in a barrier
            e2e = run_e2e()
        except Exception as exc:
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "unavailable": "transfer2_init leg failed: " + str(exc)[:400]}
            set_omp_threads(max(1, (os.cpu_count() or 1) // world))
    D.cpu_barrier()

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    lib_run = None
    if rank == 0 and with_cpu and args.max_pairs is None:
        try:
            lib_run = time_reference_library(name, args.cpu_seconds, 1, 0)   # the reference's own transfer2.c when oracle/_ref is there
        except Exception as exc:                     # falls back to the oracle port below
            sys.stderr.write(f"bench.py: the reference library could not be timed ({str(exc)[:200]}); using the oracle port\n")
    if lib_run is not None:
        cpu = {"value": lib_run["value"], "unit": UNIT, "cores": lib_run["cores"], "cpu_model": cpu_model(), "kind": "reference",
               "sample": f"transfer2_init of the reference's own source/transfer2.c (oracle/_ref/libsong_ref.so) on the whole (k1,k2) triangle of "
                         f"{lib_run['k_values']} of the {lib_run['k_size_full']} k values (midpoints of equal strata): {lib_run['pairs']} pairs, {lib_run['integrals']} "
                         f"integrals, one call of {lib_run['seconds_per_step']:.1f} s",
               "probe_bitwise_equal_to_oracle": lib_run["probe_bitwise_equal_to_oracle"]}
    elif rank == 0 and with_cpu:
        grids_c, sub_c, proj_c, stride = cpu_sample(name, args.cpu_seconds, max_pairs=args.max_pairs)
        run = O.OracleRun(sub_c, proj_c)
        t0 = time.perf_counter(); run.run(); dt = time.perf_counter() - t0
        cpu = {"value": sub_c.n_integrals / dt, "unit": UNIT, "cores": O.lib().oracle_num_threads(), "cpu_model": cpu_model(), "kind": "port",
               "sample": f"every {stride}-th (k1,k2) pair, {sub_c.n_pairs} of {grids_c.n_pairs} pairs, "
                         f"{sub_c.n_integrals} integrals, one pass of {dt:.1f} s; CPU restatement of reference transfer2 "
                         f"(gcc -O3 -fopenmp, reference loop structure; reference not buildable here)"}

    if rank != 0:
        return None
    return {
        "workload": name, "value": value, "ms_per_step": ms_step, "steps": steps, "warmup": warmup,
        "chunks_per_gpu": chunks, "n_groups": n_groups, "n_bands": n_bands, "per_rank_ms": rank_ms,
        "source_replication": float(sum(shard.shard_units_index(grids, r, world, sched)[0].size for r in range(world)) / max(grids.n_pairs, 1)), "imbalance_max_over_mean": max(rank_ms) / (sum(rank_ms) / len(rank_ms)),
        "schedule_load_max_over_mean": float(load.max() / load.mean()),
        "gpu_launches": launches_total, "describe": workload.describe(grids), "parity": parity, "gather": gather,
        "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "setup_s": t_setup, "k3_integral": k3int, "clocks": clocks,
        "l2_gb_streamed": (3 * sub.src.nbytes + n_int_rank * 8) / 1e9,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--max-pairs", type=int, default=None, help="debug: truncate the pair list (disables the transfer2_init leg)")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="CPU time per oracle pass of the baselines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed transfer2_init calls (at most --steps)")
    ap.add_argument("--chunks-per-rank", type=int, default=0, help="contiguous pair chunks per GPU in the static schedule (0: the rule of transfer2_init)")
    ap.add_argument("--bands", type=int, default=0, help="k3 bands of the (pair groups x k3 bands) schedule (0: song_host_band_count, the "
                                                        "fewest that keep 80 k3 values per projection-table row on every GPU)")
    ap.add_argument("--check-pairs", type=int, default=8, help="pairs checked against the CPU oracle per workload")
    ap.add_argument("--no-k3-integral", dest="k3_integral", action="store_false", help="skip the bispectrum k3-integral leg (SURVEY 8f rank 1)")
    ap.add_argument("--dropin-verbose", type=int, default=0, help="debug: ptr2->transfer2_verbose of the last timed transfer2_init")
    ap.add_argument("--extra", default="auto", help="comma list of further workloads for the `configs` array; auto: the other "
                                                  "single-GPU BASELINE configs at N=1, configs[4] at N=8; none: skip")
    ap.add_argument("--extra-deadline", type=float, default=660.0,
                    help="seconds after the start of the process by which the `configs` legs must be over: no further workload is started "
                         "in the last quarter of it, and one still running at the deadline is abandoned -- the main line, measured before, "
                         "goes out with a note in `configs` (0: no deadline)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)         # before any pinned allocation: page-locked buffers are first-touch local
    if world > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        # torchrun exports OMP_NUM_THREADS=1 to its workers: give every rank its share of the host for the host-side
        # work of the set-up (bessel2's j_l1 / 3j tables, the oracle of the parity spot checks)
        set_omp_threads(max(1, (os.cpu_count() or 1) // world))
    D = Dist(rank, world, local_rank)

    sampler = ClockSampler(local_rank, period_ms=100) # over the timed `value` region only: nvidia-smi queries take driver
                                                     # locks and stretched single transfer2_init calls from 34 to 120-480 ms
    main_res = measure_workload(args.workload, args, D, args.steps, args.warmup, min(args.e2e_steps, args.steps),
                                with_cpu=(world == 1 and not args.no_cpu_baseline), sampler=sampler)
    clocks = main_res.get("clocks") if rank == 0 else None

    extras = []
    if args.extra == "auto":
        names = []
        if args.workload == DEFAULT_WORKLOAD and args.max_pairs is None:
            if world == 1:
                names = ["quick_song_run", "sn_t_10percent_lmax2000", "sn_pol_10percent_lmax2000"]
            elif world == 8:
                names = ["sn_pol_5percent_lmax4000"]
    elif args.extra in ("none", ""):
        names = []
    else:
        names = [n for n in args.extra.split(",") if n]
    # The main workload is measured; the `configs` legs that follow must not be able to lose it.  A leg is not started in the
    # last quarter of --extra-deadline, and a watchdog on every rank abandons one that is still running at the deadline
    # (rank 0 prints the line first): a slow host or a stuck leg costs that leg, not the run.
    emitted = threading.Lock()
    state = {"current": None}
    has_deadline = args.extra_deadline > 0

    def emit_line():
        """The one JSON line of the run, once (rank 0): the main workload and whatever `configs` legs are complete."""
        if not emitted.acquire(blocking=False):
            return
        if rank != 0:
            return
        r = main_res
        line = {
            "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(r["describe"], workload=args.workload, mode=args.mode,
                           parallelism=f"ONE workload over {world} GPU(s) = {r['n_groups']} (k1,k2)-pair group(s) x {r['n_bands']} k3 band(s): pairs by "
                                       f"song_host_schedule_pairs ({r['chunks_per_gpu']} contiguous chunk(s) per group, LPT under the cost model of los_host.c: "
                                       f"k3 density per projection-table row, active multipole share, coarse points touched), "
                                       f"bands = cost quantiles of a group's fine k3 values (song_host_band_edges; a band keeps the k3 density per "
                                       f"projection-table row, the coarse sources of a pair are resident on every band it crosses: x{r['source_replication']:.2f}); "
                                       f"no data-path collective; results gathered on rank 0",
                           l2_policy="inputs larger than L2: each step streams ~%.1f GB per GPU (sources, transposed sources, "
                                     "interpolated slab, projection table, output) against 126 MB of L2" % r["l2_gb_streamed"],
                           tau_sampling="synthetic, geometric 230->330 then linear ->700 Mpc",
                           cpu_affinity=numa),
            "per_rank_ms": r["per_rank_ms"], "imbalance": r["imbalance_max_over_mean"],
            "schedule_load_max_over_mean": r["schedule_load_max_over_mean"],
            "parity": r["parity"],
            "clocks": clocks,
            "e2e": r["e2e"] if r["e2e"] is not None else {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                          "unavailable": "transfer2_init leg needs libsong_transfer2.so and the full pair list"},
            "gpu_launches": r["gpu_launches"],
            "roofline": r["roofline"],
            "cpu_baseline": r["cpu_baseline"],
            "k3_integral": r["k3_integral"],
            "configs": list(extras),
        }
        if r["gather"] is not None:
            line["nccl_gather"] = r["gather"]
            line["nccl_gather_ms"] = r["gather"]["ms"]
        print(json.dumps(line), flush=True)

    legs_done = threading.Event()

    def watchdog():
        """One per process, from the first `configs` leg to the end of the run: whatever a rank is stuck in at the deadline --
        a leg, or a collective the others never entered because a leg failed on one rank only -- it leaves with status 0,
        rank 0 after printing the line."""
        while not legs_done.wait(0.5):
            if time.time() >= T_PROCESS_START + args.extra_deadline + (0.0 if rank == 0 else 3.0):   # rank 0 prints before the others leave
                if state["current"] is not None:
                    extras.append({"workload": state["current"],
                                   "error": f"leg abandoned: still running {args.extra_deadline:.0f} s after the start of the process (--extra-deadline)"})
                emit_line()
                sys.stdout.flush()
                os._exit(0)                          # every rank leaves at its own copy of the deadline; nothing of the main line is pending
                return

    if names and has_deadline:
        threading.Thread(target=watchdog, daemon=True).start()
    for n in names:
        start_leg = 1.0
        if has_deadline:
            start_leg = 0.0 if time.time() > T_PROCESS_START + 0.75 * args.extra_deadline else 1.0
            start_leg = D.reduce(start_leg, "min")   # decided together: no rank may enter a leg the others skip
        if start_leg < 0.5:
            if rank == 0:
                extras.append({"workload": n, "skipped": f"not started: more than {0.75 * args.extra_deadline:.0f} s of the {args.extra_deadline:.0f} s "
                                                         "of --extra-deadline were used when its turn came"})
            continue
        big = n.startswith("sn_pol_5percent")
        state["current"] = n
        try:
            r = measure_workload(n, args, D, steps=3 if big else 5, warmup=3, e2e_steps=0 if n == "sn_pol_5percent_lmax4000" else 3,
                                 with_cpu=False)
        except RuntimeError as exc:                  # raised on every rank alike (see measure_workload): the main line still goes out
            state["current"] = None
            if rank == 0:
                extras.append({"workload": n, "error": str(exc)[:300]})
            continue
        state["current"] = None
        if rank == 0:
            extras.append({"workload": n, "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "steps": r["steps"],
                           "n_gpus": world, "integrals": r["describe"]["integrals"], "parity": r["parity"],
                           "gather": r["gather"], "imbalance_max_over_mean": r["imbalance_max_over_mean"],
                           "roofline": {k: r["roofline"][k] for k in ("achieved", "peak", "unit", "frac", "step_frac", "kernel_ms_per_step",
                                                                      "kernel_share_of_step", "resident_batches", "frac_per_rank")},
                           "e2e": None if r["e2e"] is None else {k: r["e2e"].get(k) for k in ("value", "ms_per_step", "ms_each_step", "steps", "first_call_ms", "parity",
                                                                                              "h2d_bytes_per_step", "d2h_bytes_per_step", "unavailable")}})

    state["current"] = None
    emit_line()
    sys.stdout.flush()
    D.close()                                      # the watchdog, if any, still bounds this: a process group that cannot shut down
    legs_done.set()                                # cleanly (a rank that left a leg early) ends at the deadline, the line being out


if __name__ == "__main__":
    main()
