import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `pytest -m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding

    binding.load()
    return binding


@pytest.fixture(scope="session")
def crlib():
    """The product library.  Missing .so => error, never a skip: there is no CPU fallback."""
    from crucible_b200 import abi

    if not os.path.exists(abi.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return abi.load()


@pytest.fixture(scope="session")
def gpu_device(crlib):
    n = crlib.cr_device_count()
    assert n > 0, "no sm_100 device visible: GPU tests must run on a B200 (there is no CPU fallback)"
    return 0


def random_rays(n, lo, hi, seed):
    """Batch (iii) of SURVEY 8d: origins uniform in 1.2x the scene box, directions uniform on the sphere."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    c, h = (lo + hi) / 2, (hi - lo) / 2 * 1.2
    o = c + (rng.random((n, 3)) * 2 - 1) * h
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d, np.zeros((n, 1))], axis=1)
