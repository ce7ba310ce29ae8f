import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib_built():
    from metakssd_b200 import build as B
    B.build()
    import metakssd_b200 as M
    return M


_SHUF_CACHE = {}


@pytest.fixture(scope="session")
def shuf(oracle):
    """shuf(seed, k, subk, L) -> (shuf_id, perm), cached per session."""
    def get(seed, k, subk, L):
        key = (seed, k, subk, L)
        if key not in _SHUF_CACHE:
            _SHUF_CACHE[key] = oracle.make_shuf(seed, k, subk, L)
        return _SHUF_CACHE[key]
    return get
