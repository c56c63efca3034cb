"""GPU parity on the BASELINE.json workloads beyond configs[0] (which test_parity_gpu.py runs at full size): sampled
(k1,k2) pairs of configs[1]-[4] through the C-ABI, FAST <= 1e-9 above 1e-6 of the per-(l,m) maximum and EXACT bit for
bit on a few pairs; plus the row-index boundary case of the tensor-core path (x an exact multiple of xx_step).

The projection tables of the big workloads are built by the product (Bessel2: host j_l1 / 3j, device J_Llm assembly)
and handed to the oracle, so that both sides integrate against the same numbers (parity "on the same inputs")."""
import numpy as np
import pytest

from song_b200 import configs, workload
from song_b200.bessel2 import Bessel2
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9
FLOOR = 1e-6


def rel_err(got, ref):
    scale = np.abs(ref).max(axis=1, keepdims=True)
    sel = np.abs(ref) > FLOOR * scale
    return float(np.max(np.abs(got - ref)[sel] / np.abs(ref)[sel])) if sel.any() else 0.0


# (workload, pairs for FAST, pairs for EXACT): configs[1] and [2] with >= 32 sampled pairs, [3] and [4] strided
CASES = [("sn_t_10percent_lmax2000", 48, 4), ("sn_pol_10percent_lmax2000", 40, 4),
         ("sn_pol_5percent_lmax3000", 20, 2), ("sn_pol_5percent_lmax4000", 12, 2)]


@pytest.mark.parametrize("name,n_fast,n_exact", CASES, ids=[c[0] for c in CASES])
def test_baseline_workload_sampled_pairs(oracle, name, n_fast, n_exact):
    grids = workload.build_grids(configs.get(name))
    idx = np.unique(np.linspace(0, grids.n_pairs - 1, n_fast).astype(int))      # strided over the whole (k1,k2) triangle
    sub = workload.fill_sources(grids.select(idx))
    eng = LosEngine(sub, device=0)
    b2 = Bessel2(sub).init(eng)
    rows, xmin, xsize = b2.tables(eng)
    oproj = oracle.OracleProjection(rows, xmin, xsize, None)
    ref, _ = oracle.OracleRun(sub, oproj).run()
    got = eng.transfer_pairs(sub, mode=MODE_FAST)
    err = rel_err(got, ref)
    assert err < REL_TOL, (name, err)
    # reference statement order on a few of the same pairs: bit for bit
    few = np.unique(np.linspace(0, sub.n_pairs - 1, n_exact).astype(int))
    sub2 = sub.select(few)
    got2 = eng.transfer_pairs(sub2, mode=MODE_EXACT)
    ref2 = np.concatenate([ref[:, sub.tr_off[i]:sub.tr_off[i] + sub.k3_tr_size[i]] for i in few], axis=1)
    assert np.array_equal(got2, ref2), (name, float(np.abs(got2 - ref2).max()))
    eng.close()


def test_row_index_on_grid_points(oracle):
    """transfer2.c:2199: index_x = (int)(x/xx_step).  The tensor-core path multiplies by 1/xx_step and must fall back to
    the division wherever the two can disagree.  Sources that are an impulse in time at a tau with tau0 - tau exactly
    representable, and fine k3 values moved onto x = m xx_step: a row index off by one would put a unit on the other
    side of index_xmin_J for some l, i.e. change WHICH terms pass the skip test of transfer2.c:2232 -- the zero pattern
    of the result must therefore be the oracle's, and the values agree to the FAST tolerance."""
    tau = 230.0 + 20.0 * np.arange(24)
    p = configs.tiny(output=["tBisp"], sources_k3_interpolation="linear", tau_sampling=list(tau))
    wl = workload.build_grids(p)
    workload.fill_sources(wl)
    t_star = 3
    d = wl.cosmo["tau0"] - wl.tau[t_star]
    tp2, ntau = wl.layout.tp2_size, wl.tau.size
    moved = 0
    for ip in range(wl.n_pairs):
        n = int(wl.k3_pt_size[ip])
        blk = wl.src[wl.src_off[ip]:wl.src_off[ip] + tp2 * ntau * n].reshape(tp2, ntau, n)
        blk[...] = 0.0
        blk[:, t_star, :] = 1.0 / (1.0 + np.arange(tp2))[:, None]              # impulse at t_star, every L
        a, m = int(wl.tr_off[ip]), int(wl.k3_tr_size[ip])
        g = wl.k3_tr[a:a + m]
        lo, hi = int(wl.phys_start[ip]) + 1, int(wl.phys_start[ip]) + int(wl.phys_size[ip]) - 1   # interior physical points only
        for i in range(lo, hi):
            knew = (np.round(g[i] * d / wl.xx_step) * wl.xx_step) / d
            if g[i - 1] < knew < g[i + 1]:
                g[i] = knew; moved += 1
    assert moved > 100
    x = wl.k3_tr * d
    on_grid = np.abs(x / wl.xx_step - np.round(x / wl.xx_step)) < 1e-9
    assert on_grid.sum() > 100
    oproj = oracle.build_projection(wl)
    ref, _ = oracle.OracleRun(wl, oproj).run()
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size))
    got = eng.transfer_pairs(wl, mode=MODE_FAST)
    exact = eng.transfer_pairs(wl, mode=MODE_EXACT)
    eng.close()
    assert np.array_equal(exact, ref)
    assert np.array_equal(got == 0.0, ref == 0.0), int(((got == 0.0) != (ref == 0.0)).sum())
    nz = ref != 0.0
    assert np.max(np.abs(got - ref)[nz] / np.abs(ref)[nz]) < 1e-9             # single-term sums: no floor needed
