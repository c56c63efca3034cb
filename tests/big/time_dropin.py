"""Wall time of transfer2_init (the drop-in module, song_b200/host/transfer2_gpu.c) on a BASELINE workload, as a SONG
run sees it: SONG structs filled by the test harness, bessel2_init (drop-in) leaves the engines on the GPUs, sources in
ppt2->sources, result in ptr2->transfer.  Steps: [fill sources (untimed: perturbations2's job)] -> transfer2_init (timed)
-> checksum through ptr2->transfer -> transfer2_free.
usage: python tests/big/time_dropin.py [NAME] [--reps N] [--ref-bessel2] [--oracle-pairs M]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import dropin_lib, oracle_lib as O  # noqa: E402
from song_b200 import configs, workload  # noqa: E402
from song_b200.engine import Projection  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name", nargs="?", default="sn_t_10percent_lmax2000")
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--oracle-pairs", type=int, default=6)
ap.add_argument("--ref-bessel2", action="store_true", help="struct bessels2 as the reference's bessel2.c leaves it (tables from the oracle): transfer2_init builds its own engines")
ap.add_argument("--verbose", type=int, default=2)
a = ap.parse_args()
t = time.perf_counter()
wl = workload.build_grids(configs.get(a.name)); workload.fill_sources(wl)
print(f"workload {a.name}: {wl.n_pairs} pairs, {wl.n_units} k3 values, {wl.n_integrals} integrals, sources {wl.src.nbytes/1e9:.2f} GB "
      f"({time.perf_counter()-t:.1f} s to generate)", flush=True)
proj = None
if a.ref_bessel2:
    op = O.build_projection(wl)
    proj = Projection(op.rows, op.index_xmin, op.x_size)
t = time.perf_counter()
h = dropin_lib.Harness(wl, proj, verbose=0, keep_sources=True)
print(f"harness + {'reference-style bessels2' if a.ref_bessel2 else 'bessel2_init (drop-in)'}: {time.perf_counter()-t:.2f} s", flush=True)
for i in range(a.reps):
    h.inp.verbose = 0
    if i == a.reps - 1 and a.verbose:
        import ctypes
        # verbosity is read from ptr2 at call time: flip it through the harness for the last repetition
        dropin_lib.lib().harness_set_verbose.argtypes = [ctypes.c_void_p, ctypes.c_int]
        dropin_lib.lib().harness_set_verbose(h.h, a.verbose)
    h.fill_sources()
    t = time.perf_counter(); h.transfer2_init(); dt = time.perf_counter() - t
    t = time.perf_counter(); cs = h.checksum(); dc = time.perf_counter() - t
    print(f"transfer2_init #{i}: {dt*1e3:.1f} ms, {wl.n_integrals/dt:.3e} integrals/s; checksum {cs:.12e} ({dc*1e3:.1f} ms)", flush=True)
    if i + 1 < a.reps:
        h.transfer2_free()
got, count, ok = h.read()
assert count == wl.n_integrals and ok
if a.oracle_pairs > 0:
    rows, xmin, xsize, dd, info = h.bessel2_tables() if proj is None else (op.rows, op.index_xmin, op.x_size, None, None)
    oproj = O.OracleProjection(rows, xmin, xsize, None)
    idx = np.unique(np.linspace(0, wl.n_pairs - 1, a.oracle_pairs).astype(int))
    ref, _ = O.OracleRun(wl.select(idx), oproj).run()
    sel = np.concatenate([got[:, wl.tr_off[i]:wl.tr_off[i] + wl.k3_tr_size[i]] for i in idx], axis=1)
    scale = np.abs(got).max(axis=1, keepdims=True); m = np.abs(ref) > 1e-6 * scale
    print("max rel err above floor:", float(np.max(np.abs(sel - ref)[m] / np.abs(ref)[m])))
h.close()
