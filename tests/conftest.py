import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def smplx_data():
    from airpose_b200 import synthetic
    return synthetic.make_smplx_model(0)


@pytest.fixture(scope="session")
def smplx_oracle(smplx_data):
    import airpose_oracle as orc
    return orc.SmplxModel(smplx_data)


@pytest.fixture(scope="session")
def net_state():
    from airpose_b200 import synthetic
    return synthetic.make_network_state(123)


@pytest.fixture(scope="session")
def golden_twoview():
    return dict(np.load(os.path.join(GOLDEN, "twoview_b2.npz")))


@pytest.fixture(scope="session")
def golden_lbs():
    return dict(np.load(os.path.join(GOLDEN, "smplx_lbs.npz")))


def rel_err(a, b):
    """max |a-b| / max |b| -- the 'relative fp32' measure used throughout."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
