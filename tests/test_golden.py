"""Golden vectors produced by the reference itself (tests/golden/make_golden.py: ptr2->transfer after the reference's own
transfer2_init, run in the build container) against the oracle (CPU) and against the CUDA path through the C-ABI (GPU).
The inputs are rebuilt by the deterministic workload generator; their digests are stored with the outputs, so a drifted
generator is told apart from a wrong result."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

NAMES = sorted(G.CASES)


def load(name):
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    wl, proj = G.inputs_of(name)
    d = G.input_digests(wl, proj)
    if any(str(z[k]) != d[k] for k in d):
        pytest.skip(f"{name}: the regenerated inputs differ from the ones the fixture was computed from "
                    "(another libm / numpy than the build container's): regenerate with tests/golden/make_golden.py")
    return z, wl, proj


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_reference_golden(oracle, name):
    z, wl, proj = load(name)
    ref, _ = oracle.OracleRun(wl, proj).run()
    assert int(z["count"]) == wl.n_integrals and z["transfer"].shape == ref.shape
    assert np.array_equal(ref, z["transfer"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_equals_reference_golden(oracle, name):
    from song_b200.engine import MODE_EXACT, MODE_FAST, LosEngine, Projection
    z, wl, proj = load(name)
    gold = z["transfer"]
    eng = LosEngine(wl, device=0)
    eng.set_projection(Projection(proj.rows, proj.index_xmin, proj.x_size, proj.ddrows))
    got = eng.transfer_pairs(wl, mode=MODE_EXACT)
    fast = eng.transfer_pairs(wl, mode=MODE_FAST)
    eng.close()
    assert np.array_equal(got, gold), f"EXACT differs from the reference's own output, max abs diff {np.abs(got - gold).max()}"
    scale = np.abs(gold).max(axis=1, keepdims=True)
    sel = np.abs(gold) > 1e-6 * scale
    assert float(np.max(np.abs(fast - gold)[sel] / np.abs(gold)[sel])) < 1e-9
