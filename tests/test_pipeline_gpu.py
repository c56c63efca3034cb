"""GPU: the asynchronous, double-buffered C-ABI (song_los_submit / song_los_wait) against the
synchronous one and the oracle.

The pipeline overlaps the host->device copy of batch n+1 and the device->host copy of batch n-1
with the kernels of batch n (the k1 levels of transfer2_init are independent, transfer2.c:549-551);
it must not change a single bit of the result.
"""
import numpy as np
import pytest

from song_b200 import configs, workload
from song_b200._lib import SongLosError
from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, PinnedArray, Projection

pytestmark = pytest.mark.gpu


def _engine(oracle, wl):
    oproj = oracle.build_projection(wl)
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(oproj.rows, oproj.index_xmin, oproj.x_size))
    return eng, oproj


def _shards(wl, n):
    """n disjoint (k1,k2) shards of different sizes (pairs dealt round-robin with a skew)."""
    idx = np.arange(wl.n_pairs)
    cuts = np.linspace(0, wl.n_pairs, n + 1).astype(int)
    cuts[1:-1] += np.arange(1, n) % 3            # ragged on purpose
    return [wl.select(idx[cuts[i]:cuts[i + 1]]) for i in range(n)]


@pytest.mark.parametrize("mode", [MODE_EXACT, MODE_FAST])
def test_pipeline_is_bitwise_equal_to_sync_calls(oracle, mode):
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl)
    eng, _ = _engine(oracle, wl)
    shards = _shards(wl, 5)
    ref = [eng.transfer_pairs(s, mode=mode).copy() for s in shards]
    outs = [PinnedArray((wl.layout.tt2_size, s.n_units)) for s in shards]
    for o in outs:
        o.array[:] = np.nan
    # up to three batches in flight, every slot reused
    shards, ref, outs = shards * 2, ref * 2, outs + [PinnedArray(o.array.shape) for o in outs]
    done = 0
    for i in range(len(shards)):
        eng.submit(shards[i], mode, outs[i].array)
        if i >= 2:
            tm = eng.wait()
            assert tm["n_integrals"] == shards[done].n_integrals
            assert np.array_equal(outs[done].array, ref[done])
            done += 1
    while done < len(shards):
        eng.wait()
        assert np.array_equal(outs[done].array, ref[done])
        done += 1
    # and the sync API still works on the same context afterwards
    again = eng.transfer_pairs(shards[2], mode=mode)
    assert np.array_equal(again, ref[2])
    eng.close()


def test_pipeline_matches_oracle_exactly(oracle):
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl, kind="noise")
    eng, oproj = _engine(oracle, wl)
    shards = _shards(wl, 3)
    outs = [np.empty((wl.layout.tt2_size, s.n_units)) for s in shards]      # pageable memory is allowed
    for s, o in zip(shards[:2], outs[:2]):
        eng.submit(s, MODE_EXACT, o)
    eng.wait()
    eng.submit(shards[2], MODE_EXACT, outs[2])
    eng.wait(); eng.wait()
    for s, o in zip(shards, outs):
        ref, _ = oracle.OracleRun(s, oproj).run()
        assert np.array_equal(o, ref)
    eng.close()


def test_pipeline_depth_and_ordering_errors(oracle):
    wl = workload.build_grids(configs.tiny())
    workload.fill_sources(wl)
    eng, _ = _engine(oracle, wl)
    with pytest.raises(SongLosError, match="nothing submitted"):
        eng.wait()
    outs = [np.empty((wl.layout.tt2_size, wl.n_units)) for _ in range(4)]
    for i in range(3):
        eng.submit(wl, MODE_FAST, outs[i])
    with pytest.raises(SongLosError, match="already in flight"):
        eng.submit(wl, MODE_FAST, outs[3])
    with pytest.raises(SongLosError, match="in flight"):
        eng.upload(wl)
    with pytest.raises(SongLosError, match="in flight"):
        eng.transfer_pairs(wl, mode=MODE_FAST)
    eng.wait(); eng.wait(); eng.wait()
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    eng.close()


def test_pipeline_reports_batch_errors(oracle):
    """An inconsistent batch is refused at submit time; a device-side error (x beyond xx_max,
    transfer2.c:2172) surfaces at wait time in the reference's words."""
    wl = workload.build_grids(configs.tiny())
    workload.fill_sources(wl)
    eng, _ = _engine(oracle, wl)
    out = np.empty((wl.layout.tt2_size, wl.n_units))
    bad = wl.select(np.arange(wl.n_pairs))
    bad.src = wl.src.copy()
    bad.k3_tr = wl.k3_tr * 50.0
    eng.submit(bad, MODE_FAST, out)
    with pytest.raises(SongLosError, match="not enough J"):
        eng.wait()
    # the context stays usable
    good = eng.transfer_pairs(wl, mode=MODE_FAST)
    assert np.isfinite(good).all()
    eng.close()


def test_max_batch_units_and_timing_fields(oracle):
    wl = workload.build_grids(configs.tiny(sources_k3_interpolation="cubic"))
    workload.fill_sources(wl)
    eng, _ = _engine(oracle, wl)
    ratio = wl.src.size / wl.n_units
    nf = eng.max_batch_units(MODE_FAST, ratio)
    ne = eng.max_batch_units(MODE_EXACT, ratio)
    info = eng.device_info()
    assert 1000 < nf < ne
    per_unit_fast = info["free"] / nf
    assert per_unit_fast > 8 * wl.layout.tt2_size            # at least the result rows
    out = PinnedArray((wl.layout.tt2_size, wl.n_units))
    eng.submit(wl, MODE_FAST, out.array)
    tm = eng.wait()
    assert tm["ms_h2d"] > 0 and tm["ms_d2h"] > 0 and tm["ms_total"] > 0 and tm["n_launches"] >= 8
    eng.close()
