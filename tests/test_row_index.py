"""Row index of the tensor-core path, restated on the CPU.

The reference takes index_x = (int)(x / xx_step) (transfer2.c:2199).  The producer warps of los_tc_kernel multiply by
1/xx_step and fall back to the division when the product lies within 1e-9 of an integer (tc_index_x, los_tc.cu): the
product differs from the correctly rounded quotient by a few ulp, which can change the integer part only that close to an
integer.  The statement is IEEE double arithmetic without contraction (one multiplication, one conversion, one
subtraction), so numpy reproduces it bit for bit; this test checks the claim -- the index equals the reference's for every
x -- on the grids of the BASELINE workloads (x up to xx_max = 5250 at xx_step = 0.1, 52 502 rows), on and around every grid
point.  The GPU-side check of the same property is tests/test_parity_configs_gpu.py::test_row_index_on_grid_points.
"""
import numpy as np
import pytest


def tc_index_x(x, step):
    """los_tc.cu tc_index_x, vectorised."""
    inv = 1.0 / step                                   # TcParams.inv_xx_step, computed once on the host
    q = x * inv
    ix = q.astype(np.int64)                            # (int)q: truncation, x >= 0
    f = q - ix.astype(np.float64)
    near = (f < 1e-9) | (f > 1.0 - 1e-9)
    ix[near] = (x[near] / step).astype(np.int64)       # tc_index_x_div
    return ix, near


def reference_index_x(x, step):
    return (x / step).astype(np.int64)                 # transfer2.c:2199


@pytest.mark.parametrize("step", [0.1, 0.2, 0.05, 0.25, 0.3, 1.0 / 3.0, 0.123456789])
def test_row_index_equals_the_reference_on_and_around_grid_points(step):
    rows = np.arange(0, 60000, dtype=np.float64)
    on_grid = rows * step                              # xx[i] as lin_space / i * step produce it
    cases = [on_grid]
    for k in range(1, 40):                             # neighbours of every grid point, 1 ulp ... 2^39 ulp away
        cases.append(np.nextafter(on_grid, np.inf) + (2.0 ** k - 1) * np.spacing(on_grid))
        cases.append(np.maximum(np.nextafter(on_grid, -np.inf) - (2.0 ** k - 1) * np.spacing(on_grid), 0.0))
    x = np.concatenate(cases)
    ix, near = tc_index_x(x, step)
    assert np.array_equal(ix, reference_index_x(x, step))
    assert near.any() and not near.all()               # both branches are exercised


def test_row_index_equals_the_reference_on_products_k3_times_tau():
    """x = k3 (tau0 - tau) as the kernel forms it (transfer2.c:2195), on the ranges of configs[3] and [4]."""
    rng = np.random.default_rng(7)
    tau0_minus_tau = 14180.0 - np.concatenate([np.geomspace(230.0, 330.0, 120), np.linspace(331.0, 700.0, 120)])
    for step, k_max in ((0.1, 0.36), (0.2, 0.12)):
        k3 = np.concatenate([rng.uniform(1e-6, k_max, 200_000), np.linspace(1e-6, k_max, 50_000)])
        x = (k3[:, None] * tau0_minus_tau[None, ::8]).ravel()
        ix, _ = tc_index_x(x, step)
        assert np.array_equal(ix, reference_index_x(x, step))
        # k3 moved so that x falls on a grid point for one tau: the adversarial case of the GPU test
        d = tau0_minus_tau[3]
        k_on = (np.round(k3 * d / step) * step) / d
        x_on = k_on * d
        ix_on, near = tc_index_x(x_on, step)
        assert np.array_equal(ix_on, reference_index_x(x_on, step)) and near.mean() > 0.9
