import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "tests"))   # helper modules (amg_ref.py)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def csc_unpack(z, prefix):
    import scipy.sparse as sp
    shape = tuple(int(s) for s in z[prefix + "_shape"])
    return sp.csc_matrix((z[prefix + "_data"], z[prefix + "_indices"], z[prefix + "_indptr"]), shape=shape)


def x0(n, m0, seed):
    """The stand-in for Julia's rand(ComplexF64, n, m0) used throughout the tests."""
    rng = np.random.default_rng(seed)
    return rng.random((n, m0)) + 1j * rng.random((n, m0))


@pytest.fixture(scope="session")
def nep_fixtures():
    return load_golden("nep_fixtures.npz")


@pytest.fixture(scope="session")
def linear_golden():
    return load_golden("linear_golden.npz")


@pytest.fixture(scope="session")
def nlfeast_golden():
    return load_golden("nlfeast_golden.npz")
