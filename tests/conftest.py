import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz")))


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build(ref=False)
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle
    if not oracle.have_reference():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle.Reference()


@pytest.fixture(scope="session")
def ctx():
    import hb_mcmc_b200 as hb
    c = hb.Context(0)  # raises HBError loudly when the CUDA path is unavailable
    yield c
    c.close()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
