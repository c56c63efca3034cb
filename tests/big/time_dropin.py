"""Wall time of transfer2_init (the drop-in module, song_b200/host/transfer2_gpu.c) on a BASELINE workload:
SONG structs filled by the test harness, sources in ppt2->sources, result in ptr2->transfer.  Includes what the
C-ABI bench does not: packing the sources into page-locked staging and scattering the result rows.
usage: python tests/big/time_dropin.py [NAME] [--reps N]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import dropin_lib, oracle_lib as O  # noqa: E402
from song_b200 import configs, workload  # noqa: E402
from song_b200.engine import Projection  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("name", nargs="?", default="sn_t_10percent_lmax2000")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--oracle-pairs", type=int, default=6)
a = ap.parse_args()
wl = workload.build_grids(configs.get(a.name)); workload.fill_sources(wl)
proj = O.build_projection(wl)
p = Projection(proj.rows, proj.index_xmin, proj.x_size)
for i in range(a.reps):
    t = time.perf_counter(); got, count, ok = dropin_lib.run_transfer2(wl, p, verbose=2 if i == a.reps - 1 else 0); dt = time.perf_counter() - t
    print(f"transfer2_init (harness incl. struct set-up and read-back): {dt*1e3:.1f} ms, {count/dt:.3e} integrals/s, "
          f"threads={os.environ.get('SONG_LOS_HOST_THREADS','auto')}", flush=True)
idx = np.unique(np.linspace(0, wl.n_pairs - 1, a.oracle_pairs).astype(int))
ref, _ = O.OracleRun(wl.select(idx), proj).run()
sel = np.concatenate([got[:, wl.tr_off[i]:wl.tr_off[i] + wl.k3_tr_size[i]] for i in idx], axis=1)
scale = np.abs(got).max(axis=1, keepdims=True); m = np.abs(ref) > 1e-6 * scale
print("max rel err above floor:", float(np.max(np.abs(sel - ref)[m] / np.abs(ref)[m])))
