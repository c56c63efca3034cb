"""A/B runs of the FAST path's tuning knobs on one BASELINE workload, built once: for every setting (a ';'-separated list
of 'ENV=VAL,ENV=VAL' groups) a fresh engine computes the resident batches and a few pairs are checked against the oracle.
usage: python tests/big/ab_knobs.py NAME "SONG_LOS_TC_RMAX=2;SONG_LOS_TC_SHAPE=12,1,7" [--reps 2]
(a value with commas is written with '/' : SONG_LOS_TC_SHAPE=12/1/7)"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O  # noqa: E402
from song_b200 import configs, workload  # noqa: E402
from song_b200.bessel2 import Bessel2  # noqa: E402
from song_b200.engine import MODE_FAST, LosEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name"); ap.add_argument("settings"); ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--pair-stride", type=int, default=1); ap.add_argument("--check-pairs", type=int, default=3)
a = ap.parse_args()
t = time.time()
wl = workload.build_grids(configs.get(a.name), pair_stride=a.pair_stride); workload.fill_sources(wl)
print(f"{a.name}: {wl.n_pairs} pairs, {wl.n_units} k3 values ({time.time()-t:.0f} s)", flush=True)
sys.path.insert(0, ROOT)
from bench import split_resident  # noqa: E402
rows = None
for setting in ["BASELINE=1"] + [s for s in a.settings.split(";") if s]:
    env = dict(kv.split("=", 1) for kv in setting.split(","))
    for k, v in env.items():
        os.environ[k] = v.replace("/", ",")
    try:
        eng = LosEngine(wl, 0)
        b2 = Bessel2(wl).init(eng)
        cap = eng.max_batch_units(MODE_FAST, wl.src.size / wl.n_units)
        batches = split_resident(wl, cap)
        for i, b in enumerate(batches):
            eng.select_slot(i); eng.upload(b)
        best = None
        for rep in range(a.reps):
            tot = {"ms_total": 0., "ms_integrate": 0., "ms_interp": 0., "ms_spline": 0.}
            for i in range(len(batches)):
                eng.select_slot(i); tm = eng.compute(MODE_FAST)
                for k in tot: tot[k] += tm[k]
            if best is None or tot["ms_total"] < best["ms_total"]: best = tot
        if rows is None:
            rows = b2.tables(eng)
        # parity on a few pairs of the last batch
        b = batches[-1]; eng.select_slot(len(batches) - 1)
        out = eng.download()
        idx = np.unique(np.linspace(0, b.n_pairs - 1, a.check_pairs).astype(int))
        ref, _ = O.OracleRun(b.select(idx), O.OracleProjection(rows[0], rows[1], rows[2], None)).run()
        got = np.concatenate([out[:, b.tr_off[i]:b.tr_off[i] + b.k3_tr_size[i]] for i in idx], axis=1)
        scale = np.abs(ref).max(axis=1, keepdims=True); m = np.abs(ref) > 1e-6 * scale
        err = float(np.max(np.abs(got - ref)[m] / np.abs(ref)[m]))
        print(f"RESULT {setting:50s} batches={len(batches)} total={best['ms_total']:9.2f} ms  los_tc={best['ms_integrate']:9.2f}  interp={best['ms_interp']:8.2f}  spline={best['ms_spline']:7.2f}  err={err:.1e}", flush=True)
        eng.close(); del eng
    except Exception as exc:
        print(f"RESULT {setting:50s} FAILED: {str(exc)[:150]}", flush=True)
    for k in env:
        os.environ.pop(k, None)
