"""pytest configuration: the `gpu` marker and shared loaders.

`-m "not gpu"` covers the oracle against its golden vectors / invariants, the
host-side problem assembly and the C-ABI export list (no compute calls on a GPU);
`-m gpu` holds the parity tests proper, which call the CUDA path through the C ABI
and compare with the oracle.  Nothing here reads /root/reference at run time
except tests explicitly skipped when it is absent.
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GAIT_DATA = os.path.join(ROOT, "hkd-mpc_b200", "data")


def GAIT_PATH(gait):
    """Gait tables (inputs of the product's workloads; generated from the reference data by tools/make_fixtures.py)."""
    return os.path.join(GAIT_DATA, f"gait_{gait}.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_pkg():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("hkd-mpc_b200")


def load_workloads():
    return importlib.import_module("hkd-mpc_b200.workloads")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def workloads():
    return load_workloads()


@pytest.fixture(scope="session")
def orc():
    import oracle_py
    oracle_py.lib()
    return oracle_py


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
