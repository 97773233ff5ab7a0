import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sigops():
    """The product: Python mirror of the reference API over libsigops.so.  No fallback of any kind."""
    lib = os.path.join(ROOT, "wgpu-sigops_b200", "libsigops.so")
    if not os.path.exists(lib):  # fresh checkout: build the product library first (nvcc, ~3 min); never a substitute
        import __graft_entry__

        __graft_entry__.build()
    import wgpu_sigops_b200 as w

    w.load()
    return w


@pytest.fixture(scope="session")
def gpu_units(sigops):
    from simlib import UnitRunner

    lib = sigops.load()
    return UnitRunner(lib.sigops_test_unit, lib.sigops_test_unit_shape)


@pytest.fixture(scope="session")
def sim_units():
    from simlib import UnitRunner, load_hostsim

    lib = load_hostsim()
    return UnitRunner(lib.hostsim_unit, lib.hostsim_unit_shape)
