import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def bundled():
    g = golden("inputs_bundled.npz")
    return {k: g[k] for k in g.files}


@pytest.fixture(scope="session")
def built():
    """Make sure the native pieces exist (compiles on first use; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build()
    return True
