import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests against the reference's own output (golden fixtures, oracle/_ref) run after the GPU-vs-oracle parity tests
    they build on: under `-x` a failure there must not hide the results of the parity suite."""
    late = {"test_golden.py", "test_reference_pin.py"}
    items.sort(key=lambda it: 1 if (it.get_closest_marker("gpu") is not None and os.path.basename(str(it.fspath)) in late) else 0)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def tiny_workload():
    from song_b200 import configs, workload
    wl = workload.build_grids(configs.tiny())
    workload.fill_sources(wl)
    return wl
