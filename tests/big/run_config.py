"""Runs one BASELINE workload end to end through the pipelined C-ABI in memory-sized k1 batches
(the way transfer2_gpu.c does), generating the synthetic sources batch by batch, and spot-checks
parity against the CPU oracle on sampled (k1,k2) pairs.  Prints one JSON line.

usage: python tests/big/run_config.py NAME [--oracle-pairs N] [--mode fast|exact] [--batch-units N] [--max-pairs N]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O  # noqa: E402
from song_b200 import configs, workload  # noqa: E402
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, PinnedArray, Projection  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name")
ap.add_argument("--mode", default="fast")
ap.add_argument("--oracle-pairs", type=int, default=8)
ap.add_argument("--batch-units", type=int, default=0)
ap.add_argument("--max-pairs", type=int, default=None)
ap.add_argument("--device", type=int, default=0)
ap.add_argument("--cold", action="store_true", help="no warm-up pass: the first batch then includes allocations and table build")
ap.add_argument("--pair-stride", type=int, default=1, help="every n-th pair only (multi-GPU shard emulation)")
args = ap.parse_args()
mode = MODE_FAST if args.mode == "fast" else MODE_EXACT

t0 = time.time()
grids = workload.build_grids(configs.get(args.name), max_pairs=args.max_pairs, pair_stride=args.pair_stride)
desc = workload.describe(grids)
print(args.name, desc, f"grids {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
t0 = time.time()
proj = O.build_projection(grids)
print(f"projection (CPU restatement of bessel2) {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
eng = LosEngine(grids, args.device)
t0 = time.time()
eng.set_projection(Projection(proj.rows, proj.index_xmin, proj.x_size))
print(f"set_projection {time.time() - t0:.1f}s", file=sys.stderr, flush=True)

lay = grids.layout
tp2, ntau = lay.tp2_size, grids.tau.size
ratio = float(grids.k3_pt_size.sum()) * tp2 * ntau / float(grids.n_units)
cap = eng.max_batch_units(mode, ratio)
target = args.batch_units or grids.n_units          # as large as device memory allows: many k3 per table row
batch_units = min(cap, target)
# batches of consecutive pairs (pairs are in k1-major order: consecutive k1 levels)
bounds, start, acc = [], 0, 0
for ip in range(grids.n_pairs):
    n = int(grids.k3_tr_size[ip])
    if acc and acc + n > batch_units:
        bounds.append((start, ip)); start, acc = ip, 0
    acc += n
bounds.append((start, grids.n_pairs))
max_units = max(int(grids.k3_tr_size[a:b].sum()) for a, b in bounds)
max_pt = max(int(grids.k3_pt_size[a:b].sum()) for a, b in bounds)
eng.reserve(mode, max(b - a for a, b in bounds), max_units, max_pt, max_pt * tp2 * ntau)
print(f"{len(bounds)} batches of <= {batch_units} k3 values (memory cap {cap}); {ratio:.1f} source doubles per k3 value",
      file=sys.stderr, flush=True)

rng = np.random.default_rng(7)
check = set(int(i) for i in rng.choice(grids.n_pairs, size=min(args.oracle_pairs, grids.n_pairs), replace=False))
DEPTH = 3
outs = [None] * DEPTH
inflight = []
kern = {"ms_total": 0.0, "ms_integrate": 0.0, "ms_interp": 0.0, "ms_spline": 0.0, "ms_sort": 0.0, "ms_h2d": 0.0, "ms_d2h": 0.0}
launches = 0
worst = 0.0
exact_equal = True
checked = []
tt_max = np.zeros(lay.tt2_size)
t_gen = 0.0
t_wall0 = time.time()
t_gpu_wall = 0.0


def finish():
    global launches, worst, exact_equal, tt_max
    sub, out, lo = inflight.pop(0)
    tm = eng.wait()
    for k in kern:
        kern[k] += tm[k]
    launches += tm["n_launches"]
    tt_max = np.maximum(tt_max, np.abs(out).max(axis=1))       # per-(l,m) maximum over every (k1,k2,k3) of the run
    for ip in sorted(check):
        if lo <= ip < lo + sub.n_pairs:
            one = sub.select(np.array([ip - lo]))
            ref, _ = O.OracleRun(one, proj).run()
            a = int(sub.tr_off[ip - lo]); got = out[:, a:a + int(sub.k3_tr_size[ip - lo])].copy()
            checked.append((got, ref))
            exact_equal = exact_equal and bool(np.array_equal(got, ref))


def parity(scale_of):
    """max |got-ref|/|ref| over the checked pairs, restricted to |ref| > 1e-6 x scale."""
    w = 0.0
    for got, ref in checked:
        scale = scale_of(ref)
        sel = np.abs(ref) > 1e-6 * np.maximum(scale, 1e-300)
        if sel.any():
            w = max(w, float(np.max(np.abs(got - ref)[sel] / np.abs(ref)[sel])))
    return w


for ib, (lo, hi) in enumerate(bounds):
    t0 = time.time()
    sub = workload.fill_sources(grids.select(np.arange(lo, hi)))
    t_gen += time.time() - t0
    if ib == 0 and not args.cold:
        # untimed warm-up of the first batch: device allocations and the tensor-core table build
        warm = np.empty((lay.tt2_size, sub.n_units))
        eng.submit(sub, mode, warm); tmw = eng.wait(); del warm
        print(f"warm-up pass: {tmw['ms_total']:.1f} ms (allocations + table build included)", file=sys.stderr, flush=True)
    if len(inflight) == DEPTH:
        finish()
    slot = ib % DEPTH
    if outs[slot] is None:
        outs[slot] = PinnedArray((lay.tt2_size * max_units,))
    out = outs[slot].array[:lay.tt2_size * sub.n_units].reshape(lay.tt2_size, sub.n_units)
    eng.submit(sub, mode, out)
    inflight.append((sub, out, lo))
    print(f"batch {ib}: pairs [{lo},{hi}) {sub.n_units} k3 values, sources generated in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
while inflight:
    finish()
wall = time.time() - t_wall0
# BASELINE.json criterion: floor = 1e-6 of the per-(l,m) maximum over the whole run; the per-pair floor is stricter
worst = parity(lambda ref: tt_max[:, None])
worst_pair_floor = parity(lambda ref: np.abs(ref).max(axis=1, keepdims=True))

n_int = grids.n_integrals
line = {
    "workload": args.name, "mode": args.mode, **{k: desc[k] for k in ("pairs", "k3_units", "tt2_size", "tp2_size", "tau_size", "L_max", "xx_size")},
    "integrals": n_int, "batches": len(bounds), "batch_units": batch_units, "memory_cap_units": cap,
    "gpu_ms": kern, "gpu_launches": launches,
    "integrals_per_s_kernels": n_int / (kern["ms_total"] * 1e-3),
    "wall_s_including_source_generation": wall, "source_generation_s": t_gen,
    "parity": {"pairs_checked": len(check), "max_rel_err_above_1e-6_of_max": worst,
               "max_rel_err_with_per_pair_floor": worst_pair_floor, "bitwise_equal": exact_equal,
               "tolerance": 1e-9 if mode == MODE_FAST else 0.0},
}
print(json.dumps(line), flush=True)
if mode == MODE_FAST and not (worst < 1e-9):
    raise SystemExit(f"parity failure: {worst}")
if mode == MODE_EXACT and not exact_equal:
    raise SystemExit("parity failure: EXACT mode is not bit-identical")
eng.close()
