import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(name):
    return os.path.join(GOLDEN, name)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture
def real_kernel_path():
    """Run the real-env step on a named kernel path for the rest of the test (``sap_real_select_kernel``):
    0 automatic, 1 generic one-CTA-per-env kernel, 2 / 3 multi-CTA path (keyed lists / exact float64 selection),
    4 first-generation shared-memory kernel, 5 automatic but with the run-time-shape instantiation of the bench kernel at
    100 x 100.  Reset to automatic afterwards."""
    from marl_sap_b200 import _lib

    lib = _lib.load()

    def select(which):
        assert lib.sap_real_select_kernel(int(which)) >= 0

    yield select
    lib.sap_real_select_kernel(0)
