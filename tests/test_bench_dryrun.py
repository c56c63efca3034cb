"""bench.py end to end on the CPU with the engine replaced by the oracle: checks the script's own logic -- schedule,
resident batches, timing bookkeeping, parity spot checks, the k3-integral leg, the `configs` array, the CPU baseline and
the shape of the ONE JSON line the driver parses -- where no GPU is available.  The numbers it prints mean nothing (the
"kernels" are the CPU oracle); the GPU runs of the same script are the measurements.

The fakes live here, in tests/: nothing under song_b200/ or in bench.py's product path knows about them.
"""
import ctypes
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from song_b200.bessel2 import Bessel2 as RealBessel2  # noqa: E402  (before the fixture swaps the module attribute)


class FakeEngine:
    """LosEngine's interface as bench.py uses it; results come from the oracle."""

    def __init__(self, wl, device=0):
        import oracle_lib as O
        self.O = O
        self.wl0 = wl
        self.slots = {}
        self.cur = 0
        self.results = {}
        self.k3_ms = 0.0
        self.proj = O.build_projection(wl if wl.parent is None else wl.parent)

    def max_batch_units(self, mode, src_per_unit, n_slots=3):
        return 10 ** 9

    def measure_peaks(self, copy=True):
        return {"dfma_tflops": 30.0, "dmma_tflops": 37.0, "copy_gbs": 6500.0}

    def set_stream(self, s):
        pass

    def select_slot(self, i):
        self.cur = i

    def upload(self, b):
        self.slots[self.cur] = b

    def compute(self, mode):
        b = self.slots[self.cur]
        if self.cur not in self.results:
            if b.parent is not None:                 # slices of fine grids (k3 bands): the oracle on the whole pairs, cut
                par = b.parent.select(b.parent_pair)
                whole, _ = self.O.OracleRun(par, self.proj).run()
                self.results[self.cur] = np.ascontiguousarray(np.concatenate(
                    [whole[:, int(par.tr_off[q]) + int(a):int(par.tr_off[q]) + int(a) + int(n)]
                     for q, (a, n) in enumerate(zip(b.slice_start, b.k3_tr_size))], axis=1))
            else:
                self.results[self.cur], _ = self.O.OracleRun(b, self.proj).run()
        return {"ms_total": 1.0, "ms_spline": 0.1, "ms_sort": 0.0, "ms_interp": 0.2, "ms_integrate": 0.7, "n_launches": 12}

    def device_result(self):
        return 0, int(self.results[self.cur].size)

    def copy_result_to(self, ptr, n):
        src = np.ascontiguousarray(self.results[self.cur])
        assert n == src.size
        ctypes.memmove(ptr, src.ctypes.data, src.nbytes)

    def set_bessel_table(self, j_l1, x_min_l1, ddj_l1=None):
        self.jl, self.jl_xmin = j_l1, x_min_l1

    def integrate_k3(self, tt, index_l1, r, n_pairs, abs_M3=0, M3_coefficient=None):
        b = self.slots[self.cur]
        T = self.results[self.cur]
        out = np.stack([self.O.integral_over_k3(b, T[int(t)], self.jl[int(row)], None, self.jl_xmin[int(row)], np.asarray(r))
                        for t, row in zip(tt, index_l1)])
        self.k3_ms = 0.5
        return out

    def close(self):
        pass


class FakeBessel2:
    def __init__(self, wl):
        self.real = RealBessel2(wl)

    def init(self, eng):
        self.eng = eng
        self.real.prepare()                          # the host half runs without a GPU
        return self

    def tables(self, eng):
        p = eng.proj
        return p.rows, p.index_xmin, p.x_size

    def host_tables(self):
        return self.real.host_tables()


class FakeEvent:
    def __init__(self, enable_timing=True):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 5.0


class FakeStream:
    cuda_stream = 0


@pytest.fixture
def fake_gpu(monkeypatch):
    import torch
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import song_b200.bessel2 as b2mod
    import song_b200.engine as engmod
    monkeypatch.setattr(engmod, "LosEngine", FakeEngine)
    monkeypatch.setattr(b2mod, "Bessel2", FakeBessel2)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    real_empty, real_zeros = torch.empty, torch.zeros
    monkeypatch.setattr(torch, "empty", lambda *a, device=None, **k: real_empty(*a, **k))
    monkeypatch.setattr(torch, "zeros", lambda *a, device=None, **k: real_zeros(*a, **k))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return torch


def run_bench(monkeypatch, argv):
    import importlib
    import bench
    importlib.reload(bench)
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "bench.py prints ONE JSON line"
    return json.loads(lines[0])


def test_bench_line_single_gpu(fake_gpu, monkeypatch):
    line = run_bench(monkeypatch, ["--workload", "quick_song_run", "--max-pairs", "40", "--steps", "2", "--warmup", "3",
                                   "--cpu-seconds", "0.3", "--extra", "quick_song_run"])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "k3_integral", "configs", "parity"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["higher_is_better"] is True
    assert line["scaling"] == "strong" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["config"]["workload"] == "quick_song_run" and "model" not in line["config"]
    assert line["parity"]["ok"] and line["parity"]["pairs_checked"] >= 2          # the fake engine IS the oracle
    r = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic", "kernel_ms_per_step", "step_frac", "frac_per_rank", "per_rank"):
        assert key in r, key
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert line["gpu_launches"] == 2 * 12
    c = line["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    assert line["k3_integral"]["bitwise_equal_to_oracle"] is True
    assert line["e2e"]["value"] is None and "unavailable" in line["e2e"]           # --max-pairs: no transfer2_init leg
    assert len(line["configs"]) == 1 and line["configs"][0]["workload"] == "quick_song_run" and line["configs"][0]["parity"]["ok"]


def test_bench_refuses_to_run_without_a_gpu(monkeypatch):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import bench
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "quick_song_run", "--max-pairs", "8"])
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    with pytest.raises(SystemExit, match="no CUDA device"):
        bench.main()


# ---- two ranks (gloo): pair groups x k3 bands, gather and the parity of the gathered array ----
def _two_rank_worker(rank, world, port, bands, out_dir, extra="none"):
    import torch
    import torch.distributed as dist
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import song_b200.bessel2 as b2mod
    import song_b200.engine as engmod
    engmod.LosEngine = FakeEngine
    b2mod.Bessel2 = FakeBessel2
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.current_stream = lambda *a, **k: FakeStream()
    torch.cuda.Event = FakeEvent
    real_empty, real_zeros, real_tensor = torch.empty, torch.zeros, torch.tensor
    torch.empty = lambda *a, device=None, **k: real_empty(*a, **k)
    torch.zeros = lambda *a, device=None, **k: real_zeros(*a, **k)
    torch.tensor = lambda *a, device=None, **k: real_tensor(*a, **k)
    import bench

    def gloo_init(self, rank, world, local_rank):               # Dist.__init__ with gloo in place of NCCL
        self.torch = torch
        self.rank, self.world, self.local_rank = rank, world, local_rank
        dist.init_process_group("gloo", rank=rank, world_size=world)
        self.dist = dist
        self.cpu_group = dist.new_group(backend="gloo")
    bench.Dist.__init__ = gloo_init
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", "quick_song_run", "--max-pairs", "60", "--steps", "2", "--warmup", "3",
                "--bands", str(bands), "--extra", extra, "--no-cpu-baseline"]
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    if rank == 0:
        open(os.path.join(out_dir, "line.json"), "w").write([l for l in buf.getvalue().splitlines() if l.startswith("{")][0])
    else:
        assert not [l for l in buf.getvalue().splitlines() if l.startswith("{")]     # rank 0 alone prints the line


@pytest.mark.parametrize("world,bands", [(2, 1), (2, 2), (4, 2)], ids=["2 groups", "2 bands", "2 groups x 2 bands"])
def test_bench_line_several_ranks_gloo(tmp_path, world, bands):
    import torch.multiprocessing as mp
    port = 29900 + (os.getpid() % 80) + 3 * world + bands
    mp.spawn(_two_rank_worker, args=(world, port, bands, str(tmp_path)), nprocs=world, join=True)
    line = json.loads(open(tmp_path / "line.json").read())
    assert line["n_gpus"] == world and line["scaling"] == "strong" and len(line["per_rank_ms"]) == world
    assert f"{world // bands} (k1,k2)-pair group(s) x {bands} k3 band(s)" in line["config"]["parallelism"]
    assert f"x {bands} k3 band(s)" in line["config"]["parallelism"]
    assert line["parity"]["ok"]
    g = line["nccl_gather"]
    assert g["parity_of_gathered_array"]["ok"] and g["parity_of_gathered_array"]["pairs_checked"] >= 2
    assert g["parity_of_gathered_array"]["max_rel_err_above_1e-6_of_max"] == 0.0     # the fake engine IS the oracle
    assert len(line["roofline"]["per_rank"]["k3_values"]) == world
    assert sum(line["roofline"]["per_rank"]["k3_values"]) == line["config"]["k3_units"]      # every k3 value on exactly one GPU
    if bands > 1:
        assert "skipped" in line["k3_integral"]


def test_bench_configs_leg_on_several_ranks_gloo(tmp_path):
    """The `configs` legs of a multi-rank run: every rank takes the start / skip decision of a leg together (a collective)
    and runs it under its own watchdog; rank 0 alone prints the line, with the leg in `configs`."""
    import torch.multiprocessing as mp
    port = 29990 + (os.getpid() % 7)
    mp.spawn(_two_rank_worker, args=(2, port, 1, str(tmp_path), "quick_song_run"), nprocs=2, join=True)
    line = json.loads(open(tmp_path / "line.json").read())
    assert line["n_gpus"] == 2 and line["parity"]["ok"]
    assert len(line["configs"]) == 1 and line["configs"][0]["workload"] == "quick_song_run"
    assert line["configs"][0]["n_gpus"] == 2 and line["configs"][0]["parity"]["ok"] and line["configs"][0]["gather"] is not None


def test_reference_arm_times_the_reference_library(monkeypatch):
    """bench.py --impl reference: the CPU arm is the reference's own transfer2.c (oracle/_ref) when that library is there,
    the oracle port otherwise; either way one JSON line with impl = reference and no GPU work."""
    import dropin_lib
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    line = run_bench(monkeypatch, ["--impl", "reference", "--workload", "quick_song_run", "--steps", "1", "--warmup", "0", "--cpu-seconds", "0.5"])
    assert line["impl"] == "reference" and line["gpu_launches"] == 0 and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    c = line["cpu_baseline"]
    assert c["value"] == line["value"] and c["cores"] >= 1 and "sample" in c
    if os.path.exists(dropin_lib.REF_PATH):
        assert c["kind"] == "reference" and c["probe_bitwise_equal_to_oracle"] is True
    else:
        assert c["kind"] == "port"


def test_bench_cpu_baseline_leg_uses_the_reference_library(fake_gpu, monkeypatch):
    """The cpu_baseline object of the main line: the reference's own transfer2.c (kind "reference") when oracle/_ref is
    there.  The transfer2_init leg needs a GPU: the drop-in library is hidden for this run."""
    import dropin_lib
    if not os.path.exists(dropin_lib.REF_PATH):
        pytest.skip("oracle/_ref/libsong_ref.so not built")
    monkeypatch.setattr(dropin_lib, "PATH", os.path.join(ROOT, "song_b200", "no_such_library.so"))
    line = run_bench(monkeypatch, ["--workload", "quick_song_run", "--steps", "1", "--warmup", "3", "--cpu-seconds", "0.5",
                                   "--extra", "none", "--no-k3-integral", "--check-pairs", "2"])
    c = line["cpu_baseline"]
    assert c["kind"] == "reference" and c["probe_bitwise_equal_to_oracle"] is True and c["value"] > 0 and c["cores"] >= 1
    assert "transfer2_init of the reference" in c["sample"]
    assert line["e2e"]["value"] is None


def test_bench_skips_a_configs_leg_that_would_start_late(fake_gpu, monkeypatch):
    """--extra-deadline: the `configs` legs cannot cost the main line.  With (almost) no time left the leg is not started."""
    line = run_bench(monkeypatch, ["--workload", "quick_song_run", "--max-pairs", "20", "--steps", "1", "--warmup", "3", "--no-cpu-baseline",
                                   "--no-k3-integral", "--check-pairs", "2", "--extra", "quick_song_run", "--extra-deadline", "0.001"])
    assert line["value"] > 0 and line["parity"]["ok"]
    assert line["configs"] == [{"workload": "quick_song_run", "skipped": line["configs"][0]["skipped"]}]
    assert "not started" in line["configs"][0]["skipped"]


def test_bench_abandons_a_configs_leg_at_the_deadline(fake_gpu, monkeypatch):
    """A `configs` leg still running at the deadline: the watchdog prints the main line (measured before) with a note and
    leaves.  os._exit is replaced by a flag here, and the stuck leg is a sleep."""
    import importlib
    import time
    import bench
    importlib.reload(bench)
    real = bench.measure_workload
    calls, exits = [], []

    def measure(name, args, D, steps, warmup, e2e_steps, with_cpu, sampler=None):
        calls.append(name)
        if len(calls) == 1:
            return real(name, args, D, steps, warmup, e2e_steps, with_cpu, sampler=sampler)
        bench.T_PROCESS_START = time.time() - 10000.0          # the deadline is long past while this leg "runs"
        t0 = time.time()
        while not exits and time.time() - t0 < 20:
            time.sleep(0.1)
        raise RuntimeError("the process would be gone by now")
    monkeypatch.setattr(bench, "measure_workload", measure)
    monkeypatch.setattr(bench.os, "_exit", lambda code: exits.append(code))
    monkeypatch.setattr(importlib, "reload", lambda m: m)         # run_bench reloads bench: keep the patched module
    line = run_bench(monkeypatch, ["--workload", "quick_song_run", "--max-pairs", "20", "--steps", "1", "--warmup", "3", "--no-cpu-baseline",
                                   "--no-k3-integral", "--check-pairs", "2", "--extra", "sn_t_10percent_lmax2000", "--extra-deadline", "5000"])
    assert exits == [0] and calls == ["quick_song_run", "sn_t_10percent_lmax2000"]
    assert line["value"] > 0 and line["parity"]["ok"]
    assert len(line["configs"]) == 1 and line["configs"][0]["workload"] == "sn_t_10percent_lmax2000"
    assert "abandoned" in line["configs"][0]["error"]
