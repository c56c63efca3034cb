"""The five BASELINE.json workloads, as the subset of `.ini` / `.pre` keys the hot path reads.

Values are transcribed from the reference's parameter files (`/root/reference/pre/*.pre`,
`/root/reference/ini/intrinsic.ini`) -- data, not code -- because `/root/reference` does not
exist on the GPU box.  Keys keep SONG's own names (input2.c:417-422, 521, 1425-1543, 1699-1782).

Everything that would come out of CLASS (background / thermodynamics) is a stated assumption of
the synthetic workload: tau0, tau_rec, rs_rec and the source time sampling.
"""
from __future__ import annotations

import copy

#: Planck-like conformal times [Mpc] assumed by the synthetic workloads
COSMO = {"tau0": 14180.0, "tau_rec": 281.0, "rs_rec": 145.0}

#: keys shared by the four sn_* files (pre/sn_pol_10percent_lmax2000.pre and siblings)
_SN_COMMON = {
    "sources2_k_sampling": "smart",
    "sources2_k3_sampling": "smart",
    "k3_size_min": 5,
    "transfer2_k3_sampling": "smart",
    "transfer2_tau_sampling": "sources",
    "tau_linstep_song": 0.6,
    "sources_time_interpolation": "cubic",
    "bessels_interpolation": "linear",
    "sources_k3_interpolation": "cubic",
    "bispectra_k3_extrapolation": "flat_extrapolation",
    "bessel_j_cut_song": 1e-15,
    "bessel_J_cut_song": 1e-10,
    "bessel_tol_x_min": 1e-4,
    "k_step_transition_song": 0.2,
    "modes_song": [0],                 # ini/intrinsic.ini:51
    "smallest_allowed_variation": 1e-10,   # CLASS default (recalled; not in /root/reference)
}

CONFIGS = {
    # ini/intrinsic.ini + pre/quick_song_run.pre (README test run)
    "quick_song_run": {
        "l_max_scalars": 100, "l_logstep": 1.35, "l_linstep": 50,
        "k_min_tau0_song": 0.1, "k_max_tau0_over_l_max_song": 2.0,
        "k_step_sub_song": 0.1, "k_step_super_song": 0.025, "k_logstep_super_song": 1.8,
        "k_step_transition_song": 0.2,
        "sources2_k_sampling": "smart", "sources2_k3_sampling": "smart", "k3_size_min": 5,
        "q_linstep_song": 0.4, "transfer2_k3_sampling": "smart",
        "transfer2_tau_sampling": "sources", "tau_linstep_song": 0.6,
        "sources_time_interpolation": "linear", "bessels_interpolation": "linear",
        "sources_k3_interpolation": "linear",
        "l_max_los_t": 5, "l_max_los_p": 5,
        "bispectra_k3_extrapolation": "flat_extrapolation",
        "extra_k3_oscillations_right": 50, "extra_k3_oscillations_left": 50,
        "bessel_x_step_song": 0.2, "bessel_j_cut_song": 1e-15, "bessel_J_cut_song": 1e-10,
        "bessel_tol_x_min": 1e-4,
        "perturb_sampling_stepsize_song": 0.4,
        "modes_song": [0], "output": ["tBisp", "eBisp"],
        "smallest_allowed_variation": 1e-10,
        "tau_size": 60,
    },
    # ini/intrinsic.ini with eBisp removed from `output` + pre/sn_t_10percent_lmax2000.pre
    "sn_t_10percent_lmax2000": dict(_SN_COMMON, **{
        "l_max_scalars": 2000, "l_logstep": 1.35, "l_linstep": 50,
        "k_min_tau0_song": 0.02, "k_max_tau0_over_l_max_song": 1.6,
        "k_step_sub_song": 0.1, "k_step_super_song": 0.025, "k_logstep_super_song": 1.2,
        "q_linstep_song": 0.4,
        "l_max_los_t": 8, "l_max_los_p": 8,
        "extra_k3_oscillations_right": 20, "extra_k3_oscillations_left": 20,
        "bessel_x_step_song": 0.2,
        "perturb_sampling_stepsize_song": 0.4,
        "output": ["tBisp"],
        "tau_size": 60,
    }),
    # ini/intrinsic.ini + pre/sn_pol_10percent_lmax2000.pre (T+E)
    "sn_pol_10percent_lmax2000": dict(_SN_COMMON, **{
        "l_max_scalars": 2000, "l_logstep": 1.25, "l_linstep": 36,
        "k_min_tau0_song": 0.02, "k_max_tau0_over_l_max_song": 1.6,
        "k_step_sub_song": 0.1, "k_step_super_song": 0.025, "k_logstep_super_song": 1.2,
        "q_linstep_song": 0.4,
        "l_max_los_t": 6, "l_max_los_p": 6,
        "extra_k3_oscillations_right": 20, "extra_k3_oscillations_left": 20,
        "bessel_x_step_song": 0.2,
        "perturb_sampling_stepsize_song": 0.4,
        "output": ["tBisp", "eBisp"],
        "tau_size": 60,
    }),
    # ini/intrinsic.ini + pre/sn_pol_5percent_lmax3000.pre
    "sn_pol_5percent_lmax3000": dict(_SN_COMMON, **{
        "l_max_scalars": 3000, "l_logstep": 1.25, "l_linstep": 20,
        "k_min_tau0_song": 0.01, "k_max_tau0_over_l_max_song": 1.4,
        "k_step_sub_song": 0.075, "k_step_super_song": 0.020, "k_logstep_super_song": 1.16,
        "q_linstep_song": 0.33,
        "l_max_los_t": 17, "l_max_los_p": 17,
        "extra_k3_oscillations_right": 50, "extra_k3_oscillations_left": 50,
        "bessel_x_step_song": 0.1,
        "perturb_sampling_stepsize_song": 0.1,
        "output": ["tBisp", "eBisp"],
        "tau_size": 240,
    }),
    # ini/intrinsic.ini + pre/sn_pol_5percent_lmax4000.pre (8xB200 headline)
    "sn_pol_5percent_lmax4000": dict(_SN_COMMON, **{
        "l_max_scalars": 4000, "l_logstep": 1.25, "l_linstep": 20,
        "k_min_tau0_song": 0.01, "k_max_tau0_over_l_max_song": 1.25,
        "k_step_sub_song": 0.075, "k_step_super_song": 0.020, "k_logstep_super_song": 1.16,
        "q_linstep_song": 0.33,
        "l_max_los_t": 24, "l_max_los_p": 24,
        "extra_k3_oscillations_right": 50, "extra_k3_oscillations_left": 50,
        "bessel_x_step_song": 0.1,
        "perturb_sampling_stepsize_song": 0.1,
        "output": ["tBisp", "eBisp"],
        "tau_size": 240,
    }),
}

#: BASELINE.json `configs`, in order
BASELINE_ORDER = [
    "quick_song_run",
    "sn_t_10percent_lmax2000",
    "sn_pol_10percent_lmax2000",
    "sn_pol_5percent_lmax3000",
    "sn_pol_5percent_lmax4000",
]


def get(name: str, **overrides) -> dict:
    """A private copy of config `name` with `overrides` applied."""
    if name not in CONFIGS:
        raise KeyError(f"unknown workload {name!r}; known: {sorted(CONFIGS)}")
    cfg = copy.deepcopy(CONFIGS[name])
    cfg["name"] = name
    cfg.update(overrides)
    return cfg


def tiny(**overrides) -> dict:
    """A seconds-scale workload for parity tests: quick_song_run geometry, l_max = 30."""
    cfg = get("quick_song_run")
    cfg.update({"name": "tiny", "l_max_scalars": 30, "l_linstep": 10, "tau_size": 24,
                "k_logstep_super_song": 2.5, "k_step_super_song": 0.08, "k_step_sub_song": 0.3,
                "extra_k3_oscillations_right": 5, "extra_k3_oscillations_left": 5})
    cfg.update(overrides)
    return cfg
