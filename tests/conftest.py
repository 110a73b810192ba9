import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_utils.npz"))


@pytest.fixture
def want_last_ids():
    """Keep last_ids in the wide forward (otherwise skipped when no geometry gradient is needed)."""
    from gags_b200 import rasterization as R
    old = R.want_last_ids
    R.want_last_ids = True
    yield
    R.want_last_ids = old
