import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as o

    if not os.path.exists(o.ORACLE_SO):
        o.build()
    return o


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def ref_lib(oracle_mod):
    """The unmodified reference, when its prebuilt .so travelled with the repo (None otherwise)."""
    return oracle_mod.load_ref()
