"""Times song_los_compute on BASELINE workloads and spot-checks parity on sampled pairs.
usage: python tests/big/time_configs.py [--mode fast|exact|both] [--oracle-pairs N] name [name...]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import oracle_lib as O  # noqa: E402
from song_b200 import configs, workload  # noqa: E402
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="+")
ap.add_argument("--mode", default="both")
ap.add_argument("--oracle-pairs", type=int, default=40)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-pairs", type=int, default=None)
args = ap.parse_args()

for name in args.names:
    p = configs.get(name)
    t = time.time(); wl = workload.build_grids(p, max_pairs=args.max_pairs); workload.fill_sources(wl)
    print(name, workload.describe(wl), f"workload {time.time()-t:.1f}s", flush=True)
    t = time.time(); proj = O.build_projection(wl); print(f"oracle projection {time.time()-t:.1f}s", flush=True)
    eng = LosEngine(wl, 0)
    t = time.time(); eng.set_projection(Projection(proj.rows, proj.index_xmin, proj.x_size)); print(f"set_projection {time.time()-t:.2f}s", flush=True)
    t = time.time(); eng.upload(wl); print(f"upload {time.time()-t:.2f}s", flush=True)
    idx = np.unique(np.linspace(0, wl.n_pairs - 1, args.oracle_pairs).astype(int))
    sub = wl.select(idx)
    t = time.time(); ref, nterms = O.OracleRun(sub, proj).run()
    dt = time.time() - t
    print(f"oracle on {idx.size} pairs: {dt:.2f}s, {sub.n_integrals/dt:.3e} integrals/s ({O.lib().oracle_num_threads()} threads), terms/integral {nterms/sub.n_integrals:.1f}", flush=True)
    modes = {"fast": [MODE_FAST], "exact": [MODE_EXACT], "both": [MODE_EXACT, MODE_FAST]}[args.mode]
    for mode in modes:
        for i in range(args.reps):
            tm = eng.compute(mode)
            print("FAST " if mode else "EXACT", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in tm.items()},
                  f"{tm['n_integrals']/(tm['ms_total']*1e-3):.3e} integrals/s", flush=True)
        out = eng.download()
        got = np.concatenate([out[:, wl.tr_off[i]:wl.tr_off[i] + wl.k3_tr_size[i]] for i in idx], axis=1)
        scale = np.abs(ref).max(axis=1, keepdims=True)
        sel = np.abs(ref) > 1e-6 * scale
        rel = np.max(np.abs(got - ref)[sel] / np.abs(ref)[sel])
        print(("FAST " if mode else "EXACT"), "bitwise equal:", bool(np.array_equal(got, ref)),
              f"max rel err above floor {rel:.3e}  max abs diff/scale {np.max(np.abs(got-ref)/scale):.3e}", flush=True)
    eng.close()
