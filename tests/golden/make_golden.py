"""Golden vectors from the REFERENCE ITSELF: ptr2->transfer as the reference's own transfer2_init leaves it
(oracle/_ref/libsong_ref.so = /root/reference/source/transfer2.c compiled where it lies, oracle/Makefile target `ref`),
for small workloads whose inputs the deterministic generator of song_b200/workload.py rebuilds anywhere.

    python tests/golden/make_golden.py          (in the build container, where /root/reference exists)

Each fixture holds the transfer functions [tt2_size][k3 values in the reference loop order], the counter of
transfer2.c:607-610, and digests of the inputs (sources, k3 grids, projection tables) so that a test can tell a changed
generator from a changed result.  The fixtures travel with the repository; the library need not.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

CASES = {
    "tiny_TE_cubic": dict(sources_k3_interpolation="cubic"),
    "tiny_TE_linear_noextrap": dict(sources_k3_interpolation="linear", bispectra_k3_extrapolation="no_extrapolation"),
    "tiny_m012_B_cubic": dict(modes_song=[0, 1, 2], sources_k3_interpolation="cubic", has_source_B=True),
}


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def inputs_of(name):
    """(workload with sources, oracle projection tables) of a case -- what the fixture's outputs were computed from."""
    import oracle_lib as O
    from song_b200 import configs, workload
    wl = workload.fill_sources(workload.build_grids(configs.tiny(**CASES[name])), kind="smooth")
    proj = O.build_projection(wl)
    return wl, proj


def input_digests(wl, proj) -> dict:
    rows = np.concatenate([r for r in proj.rows if r is not None])
    return {"sources": digest(wl.src), "grids": digest(wl.k3_pt, wl.k3_tr, wl.tau, wl.xx), "projection": digest(rows, proj.index_xmin, proj.x_size)}


if __name__ == "__main__":
    import dropin_lib
    assert os.path.exists(dropin_lib.REF_PATH), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    for name in CASES:
        wl, proj = inputs_of(name)
        with dropin_lib.Harness(wl, proj, lib_path=dropin_lib.REF_PATH) as h:
            h.transfer2_init()
            T, count, labels_ok = h.read()
            h.transfer2_free()
        assert labels_ok and count == wl.n_integrals
        d = input_digests(wl, proj)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, transfer=T, count=np.int64(count), sources=d["sources"], grids=d["grids"], projection=d["projection"])
        print(f"{name}: {T.shape} -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
