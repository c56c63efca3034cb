"""Host-side k3 integration grids (song_b200/host/los_host.c) against the oracle's restatement of
transfer2_get_k3_size / _list (transfer2.c:1329-1737): bit-identical grids and bookkeeping, plus
the invariants the reference asserts under -DDEBUG (transfer2.c:1711-1733)."""
import numpy as np
import pytest

from song_b200 import configs, workload


def both(oracle, params):
    wl = workload.build_grids(params)
    n = wl.layout.J_size * wl.layout.L_size * wl.layout.l_size * wl.layout.m_size
    proj = oracle.OracleProjection([None] * n, np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n))
    infos, k3 = oracle.OracleRun(wl, proj).k3_grids()
    return wl, infos, k3


@pytest.mark.parametrize("params", [
    configs.tiny(),
    configs.tiny(bispectra_k3_extrapolation="no_extrapolation"),
    configs.tiny(transfer2_k3_sampling="bessel"),
    configs.tiny(extra_k3_oscillations_left=-1, extra_k3_oscillations_right=-1),
    configs.tiny(extra_k3_oscillations_left=0, extra_k3_oscillations_right=30, q_linstep_song=0.17),
    configs.get("quick_song_run"),
], ids=["tiny", "no_extrap", "bessel_sampling", "full_range", "asymmetric", "quick"])
def test_grids_bit_identical_to_oracle(oracle, params):
    wl, infos, k3 = both(oracle, params)
    assert np.array_equal(k3, wl.k3_tr)
    for i, info in enumerate(infos):
        assert info.k_size == wl.k3_tr_size[i]
        assert info.k_physical_start == wl.phys_start[i]
        assert info.k_physical_size == wl.phys_size[i]
        assert info.k_min == wl.k_min_tr[i] and info.k_max == wl.k_max_tr[i]


@pytest.mark.parametrize("name", ["sn_t_10percent_lmax2000", "sn_pol_5percent_lmax3000", "sn_pol_5percent_lmax4000"])
def test_baseline_grids_bit_identical_to_oracle(oracle, name):
    """Every (k1,k2) pair of the BASELINE workloads (configs[2] shares the grids of configs[1]): the k3 grids the GPU
    parity tests feed to the oracle come from los_host.c, so they are compared here with the oracle's own
    transfer2_get_k3_size / _list restatement."""
    wl, infos, k3 = both(oracle, configs.get(name))
    assert np.array_equal(k3, wl.k3_tr)
    assert np.array_equal(np.array([i.k_size for i in infos]), wl.k3_tr_size)
    assert np.array_equal(np.array([i.k_physical_start for i in infos]), wl.phys_start)
    assert np.array_equal(np.array([i.k_physical_size for i in infos]), wl.phys_size)
    assert np.array_equal(np.array([i.k_min for i in infos]), wl.k_min_tr)
    assert np.array_equal(np.array([i.k_max for i in infos]), wl.k_max_tr)


def test_grid_invariants_quick_song_run():
    wl = workload.build_grids(configs.get("quick_song_run"))
    assert wl.n_pairs == 253 and wl.n_units == 20441
    for i in range(wl.n_pairs):
        g = wl.k3_tr[wl.tr_off[i]:wl.tr_off[i] + wl.k3_tr_size[i]]
        assert np.all(np.diff(g) > 0)                           # strictly ascending (:1727)
        k1, k2 = wl.k[wl.pairs[i, 0]], wl.k[wl.pairs[i, 1]]
        phys = g[wl.phys_start[i]:wl.phys_start[i] + wl.phys_size[i]]
        assert np.all(phys > abs(k1 - k2)) and np.all(phys < k1 + k2)   # triangular (:1719)
        assert g[-1] <= wl.k_max_tr[i] <= wl.xx_max / wl.cosmo["tau0"]   # (:1408, :1670)
        pt = wl.k3_pt[wl.pt_off[i]:wl.pt_off[i] + wl.k3_pt_size[i]]
        assert g[wl.phys_start[i]] == pt[0]


@pytest.mark.parametrize("name,expect", [
    ("sn_t_10percent_lmax2000", (120, 7260, 1599531, 54, 54, 9, 16802)),
    ("sn_pol_10percent_lmax2000", (120, 7260, 1599531, 73, 146, 14, 16802)),
    ("sn_pol_5percent_lmax3000", (179, 16110, 7492210, 165, 330, 36, 44102)),
    ("sn_pol_5percent_lmax4000", (197, 19503, 10951307, 215, 430, 50, 52502)),
])
def test_baseline_workload_shapes(name, expect):
    """The replayed samplings give the sizes of the SURVEY.md section 8 table."""
    wl = workload.build_grids(configs.get(name))
    lay = wl.layout
    assert (wl.k.size, wl.n_pairs, wl.n_units, lay.l_size, lay.tt2_size, lay.tp2_size, wl.xx.size) == expect


def test_xx_max_too_small_is_an_error():
    from song_b200 import _lib
    p = configs.tiny()
    p["k_max_tau0_over_l_max_song"] = 2.0
    wl = workload.build_grids(p)
    kp = wl.k3_params()
    kp.xx_max = 1.0        # J(x) tabulated only up to x = 1
    import ctypes as C
    err = C.create_string_buffer(512)
    r = _lib.K3Range()
    seg = np.ascontiguousarray(wl.k3_pt[:wl.k3_pt_size[0]])
    rc = _lib.host().song_host_k3_size(C.byref(kp), _lib.dptr(seg), int(seg.size), C.byref(r), err, 512)
    assert rc == 1 and b"xx_max is not large enough" in err.value
