set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python - <<'PY' 2>&1 | tail -20
import sys, time, numpy as np
sys.path.insert(0,'tests')
import oracle_lib as O
from song_b200 import configs, workload
from song_b200.engine import LosEngine, Projection, MODE_EXACT
for name in ["quick_song_run", "sn_t_10percent_lmax2000"]:
    p = configs.get(name)
    t=time.time(); wl = workload.build_grids(p); workload.fill_sources(wl); print(name, "workload", time.time()-t, flush=True)
    t=time.time(); proj = O.build_projection(wl); print("oracle proj", time.time()-t, flush=True)
    eng = LosEngine(wl, 0)
    print(eng.device_info(), eng.measure_peaks(), flush=True)
    t=time.time(); eng.set_projection(Projection(proj.rows, proj.index_xmin, proj.x_size)); print("set_proj", time.time()-t, flush=True)
    eng.upload(wl)
    for i in range(3):
        tm = eng.compute(MODE_EXACT); print(tm, "integrals/s", tm['n_integrals']/(tm['ms_total']*1e-3), flush=True)
    out = eng.download()
    print("out absmax", np.abs(out).max())
    # sampled parity on the big config
    idx = np.arange(0, wl.n_pairs, max(1, wl.n_pairs//40))
    sub = wl.select(idx)
    ref,_ = O.OracleRun(sub, proj).run()
    got = np.concatenate([out[:, wl.tr_off[i]:wl.tr_off[i]+wl.k3_tr_size[i]] for i in idx], axis=1)
    print("sampled pairs", idx.size, "bitwise equal:", np.array_equal(got, ref), "max abs diff", np.abs(got-ref).max())
    eng.close()
PY
