#!/usr/bin/env python
"""bench.py -- LOS transfer integrals per second of the transfer2 stage on N B200s.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (song_los_compute: k3 spline, k3 sort, source interpolation,
DMMA LOS kernel) over
one batch of synthetic input.  At N=1 the workload is BASELINE.json configs[1]
(ini/intrinsic.ini + pre/sn_t_10percent_lmax2000.pre, temperature only: 8.6e7 integrals).  With N
ranks every rank processes its own (k1,k2) shard of the same size (independent sources, no
data-path collective): weak scaling, value = all ranks' integrals / max-over-ranks time.

`value`  : inputs resident in HBM when the timed region starts.
`e2e`    : the same metric through the C-ABI call with HOST (pinned) buffers: every step copies the
           sources and grids host->device and the transfer functions device->host.
`--impl reference` : the CPU restatement of the reference's transfer2 loops (oracle/, OpenMP on all
           host threads) on a bounded sample of the same workload.  The reference itself cannot be
           built here (its CLASS fork is an empty submodule), so kind = "port".
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

METRIC = "2nd-order LOS transfer integrals/sec (k1,k2,k3,l,m)"
UNIT = "integrals/s"
DEFAULT_WORKLOAD = "sn_t_10percent_lmax2000"


# ----------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
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


def pinned_copy(wl):
    """The workload with its large host arrays in pinned memory (torch allocator)."""
    import copy
    import torch
    out = copy.copy(wl)
    keep = []
    for name in ("src", "k3_pt", "k3_tr"):
        t = torch.from_numpy(getattr(wl, name)).pin_memory()
        keep.append(t)
        setattr(out, name, t.numpy())
    out._pinned = keep
    return out


# ----------------------------------------------------------------------------------------------
# the reference arm: CPU restatement of transfer2 (oracle/) on the host cores
# ----------------------------------------------------------------------------------------------
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


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    # torch.distributed.run exports OMP_NUM_THREADS=1 when it is unset; the CPU arm uses every host thread
    if "LOCAL_RANK" in os.environ and os.environ.get("OMP_NUM_THREADS") == "1":
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import oracle_lib as O
    from song_b200 import workload
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
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload.describe(grids), workload=args.workload,
                       note="CPU restatement of reference transfer2 (reference not buildable here: class.git is an empty submodule)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--max-pairs", type=int, default=None, help="debug: truncate the pair list")
    ap.add_argument("--cpu-seconds", type=float, default=6.0, help="CPU time per oracle pass of the baselines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="fast", choices=["fast", "exact"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from song_b200 import configs, workload
    from song_b200.bessel2 import Bessel2
    from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine

    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank)         # before any pinned allocation: page-locked buffers are first-touch local
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    mode = MODE_FAST if args.mode == "fast" else MODE_EXACT

    # ---- inputs: every rank owns a shard of the same shape with its own sources (weak scaling) ----
    p = configs.get(args.workload)
    wl = workload.build_grids(p, max_pairs=args.max_pairs)
    workload.fill_sources(wl, seed=20261017 + 7919 * rank)
    eng = LosEngine(wl, device=local_rank)
    b2 = Bessel2(wl).init(eng)                      # bessel2_init: host j_l1 + 3j, device J_Llm assembly
    cap = eng.max_batch_units(mode, wl.src.size / max(wl.n_units, 1))
    if wl.n_units > cap:
        raise SystemExit(f"bench.py: {args.workload} has {wl.n_units} k3 values, one resident batch holds {cap} on this GPU; "
                         f"bench.py times configs[1]-sized workloads in one batch -- use tests/big/run_config.py {args.workload} "
                         "for the batched end-to-end run of the larger BASELINE workloads")
    peaks = eng.measure_peaks()
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    wl_pin = pinned_copy(wl)
    out_pin = torch.empty((wl.layout.tt2_size, wl.n_units), dtype=torch.float64).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value: inputs resident in HBM ----
    eng.upload(wl_pin)
    for _ in range(args.warmup):
        eng.compute(mode)
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern = {"ms_integrate": 0.0, "ms_spline": 0.0, "ms_sort": 0.0, "ms_interp": 0.0, "ms_total": 0.0}
    launches = 0
    barrier()
    sampler.start()
    ev0.record(stream)
    for _ in range(args.steps):
        tm = eng.compute(mode)
        for k in kern:
            kern[k] += tm[k]
        launches += tm["n_launches"]
    ev1.record(stream)
    barrier()
    ms_step = max_over_ranks(ev0.elapsed_time(ev1) / args.steps)
    for k in kern:
        kern[k] /= args.steps
    n_int_rank = wl.n_integrals
    n_int_total = n_int_rank * world
    value = n_int_total / (ms_step * 1e-3)

    # ---- e2e: host (pinned) buffers through the C-ABI, H2D + D2H inside the timed region ----
    # (a) the synchronous call, one batch at a time: song_los_transfer_pairs
    # (b) the pipeline a SONG run uses for its k1 batches: song_los_submit / song_los_wait with three
    #     batches in flight, so that the copies of neighbouring steps overlap the kernels.  Every step
    #     still copies all of its inputs to the device and its whole result back; the region is
    #     bracketed by device-wide synchronisation and timed on the host clock (three streams).
    from song_b200.engine import PinnedArray
    DEPTH = max(1, min(3, int(os.environ.get("SONG_BENCH_DEPTH", "3"))))     # batches in flight (SONG_LOS_MAX_IN_FLIGHT = 3)
    outs = [PinnedArray((wl.layout.tt2_size, wl.n_units)) for _ in range(DEPTH)]
    out_np = outs[0].array
    eng.transfer_pairs(wl_pin, mode=mode, out=out_np)          # warm
    barrier()
    t_host0 = time.perf_counter()
    n_sync = min(args.steps, 3)
    for _ in range(n_sync):
        eng.transfer_pairs(wl_pin, mode=mode, out=out_np)
    torch.cuda.synchronize()
    sync_ms = max_over_ranks((time.perf_counter() - t_host0) * 1e3 / n_sync)
    checksum = float(np.abs(out_np).sum())
    for i in range(DEPTH):                                      # warm every slot
        eng.submit(wl_pin, mode, outs[i].array)
    for i in range(DEPTH):
        eng.wait()
    barrier()
    t_host0 = time.perf_counter()
    copy_ms = {"h2d": 0.0, "d2h": 0.0}
    for i in range(args.steps):
        eng.submit(wl_pin, mode, outs[i % DEPTH].array)
        if i >= DEPTH - 1:
            tm = eng.wait(); copy_ms["h2d"] += tm["ms_h2d"]; copy_ms["d2h"] += tm["ms_d2h"]
    for i in range(min(DEPTH - 1, args.steps)):
        tm = eng.wait(); copy_ms["h2d"] += tm["ms_h2d"]; copy_ms["d2h"] += tm["ms_d2h"]
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t_host0) * 1e3 / args.steps)
    clocks = sampler.stop()          # sampled over both timed regions (value and e2e), 50 ms period
    if world > 1:
        dist.barrier()
    h2d = int(wl.src.nbytes + wl.k3_pt.nbytes + wl.k3_tr.nbytes + (3 * 8 + 5 * 4) * wl.n_pairs)
    d2h = int(out_np.nbytes)
    e2e_value = n_int_total / (e2e_ms * 1e-3)
    checksum_pipe = float(np.abs(outs[(args.steps - 1) % DEPTH].array).sum())
    if checksum_pipe != checksum:
        raise SystemExit(f"bench.py: pipelined result differs from the synchronous one ({checksum_pipe} vs {checksum})")

    # ---- multi-GPU: the one collective of the path, the gather of the per-GPU transfer arrays ----
    gather_ms = None
    if world > 1:
        res = torch.empty(wl.layout.tt2_size * wl.n_units, dtype=torch.float64, device="cuda")
        eng.copy_result_to(res.data_ptr(), res.numel())
        bufs = [torch.empty_like(res) for _ in range(world)] if rank == 0 else None
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.gather(res, bufs, dst=0)                 # untimed: the first collective builds the NCCL communicator
        barrier()
        g0.record()
        dist.gather(res, bufs, dst=0)
        g1.record()
        barrier()
        gather_ms = max_over_ranks(g0.elapsed_time(g1))

    # ---- roofline of the dominant kernel (the DMMA LOS kernel) ----
    rows, xmin, xsize = None, None, None
    lay = wl.layout
    n_j = lay.J_size * lay.L_size * lay.l_size * lay.m_size
    xmin = np.zeros(n_j, dtype=np.int32); xsize = np.zeros(n_j, dtype=np.int32)
    from song_b200 import _lib
    _lib.los().song_los_get_projection(eng.ctx, _lib.iptr(xmin), _lib.iptr(xsize), None)
    n_terms = count_terms(wl, xmin, xsize)
    flops = 5.0 * n_terms + 2.0 * n_int_rank * wl.tau.size          # F = 5 N_terms + 2 N_int n_tau
    t_kernel = kern["ms_integrate"] * 1e-3
    achieved = flops / t_kernel / 1e12
    peak = max(peaks["dmma_tflops"], peaks["dfma_tflops"])
    alg_bytes = 8.0 * (n_int_rank + wl.src.size + wl.k3_pt.size + wl.k3_tr.size)
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "traffic_bytes.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(args.workload, {}).get("los_tc_kernel")
        except Exception:
            traffic = None
    roofline = {
        "bound": "tensor", "kernel": "los_tc_kernel (FP64 tensor core, DMMA m8n8k4, TMA-fed)" if mode == MODE_FAST else "los_exact_kernel",
        "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": "measured in this run with song_los_measure_peaks (DMMA m8n8k4 loop; MEASURED_PEAKS.json has no FP64 entry)",
        "algorithmic_flops_per_launch": flops, "n_terms": n_terms, "kernel_ms": kern["ms_integrate"],
        "kernel_share_of_step": kern["ms_integrate"] / max(kern["ms_total"], 1e-9),
        "dfma_peak_tflops": peaks["dfma_tflops"], "dmma_peak_tflops": peaks["dmma_tflops"],
        "hbm": {"algorithmic_bytes_per_step": alg_bytes, "achieved_gbs": alg_bytes / (kern["ms_total"] * 1e-3) / 1e9,
                "peak_gbs": peaks["copy_gbs"], "frac": alg_bytes / (kern["ms_total"] * 1e-3) / 1e9 / peaks["copy_gbs"]},
        "step_breakdown_ms": kern,
    }

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as O
        grids, sub, proj, stride = cpu_sample(args.workload, args.cpu_seconds, max_pairs=args.max_pairs)
        run = O.OracleRun(sub, proj)
        t0 = time.perf_counter(); ref, _ = run.run(); dt = time.perf_counter() - t0
        cpu = {"value": sub.n_integrals / dt, "unit": UNIT, "cores": O.lib().oracle_num_threads(), "kind": "port",
               "sample": f"every {stride}-th (k1,k2) pair, {sub.n_pairs} of {grids.n_pairs} pairs, "
                         f"{sub.n_integrals} integrals, one pass of {dt:.1f} s; CPU restatement of reference transfer2 "
                         f"(gcc -O3 -fopenmp, reference loop structure; reference not buildable here)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(workload.describe(wl), workload=args.workload, mode=args.mode,
                           per_gpu_integrals=n_int_rank, parallelism=f"(k1,k2)-pair shards x{world}, no data-path collective",
                           l2_policy="inputs larger than L2: each step streams ~%.1f GB (sources, transposed sources, "
                                     "projection table, output) against 126 MB of L2" %
                                     ((3 * wl.src.nbytes + out_np.nbytes) / 1e9),
                           tau_sampling="synthetic, geometric 230->330 then linear ->700 Mpc",
                           cpu_affinity=numa),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "checksum_abs_sum": checksum,
                    "api": f"song_los_submit/song_los_wait, up to {DEPTH} batches in flight (copy-in, kernels, copy-out), pinned host buffers; every step copies "
                           "its inputs H2D and its result D2H; host clock between device-wide synchronisations",
                    "h2d_ms_per_step": copy_ms["h2d"] / args.steps, "d2h_ms_per_step": copy_ms["d2h"] / args.steps,
                    "h2d_gbs": h2d / 1e9 / max(copy_ms["h2d"] / args.steps * 1e-3, 1e-12),
                    "d2h_gbs": d2h / 1e9 / max(copy_ms["d2h"] / args.steps * 1e-3, 1e-12),
                    "sync_call": {"value": n_int_total / (sync_ms * 1e-3), "ms_per_step": sync_ms,
                                  "api": "song_los_transfer_pairs (upload, kernels, download back to back)"}},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
        if gather_ms is not None:
            line["nccl_gather_ms"] = gather_ms
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
