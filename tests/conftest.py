import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def keyset(oracle):
    """The shared keyset: oracle keygen, seed 0 (uploaded unchanged to the GPU engine in gpu tests)."""
    return oracle.keygen(0)


@pytest.fixture(scope="session")
def engine(keyset):
    import redsec_b200 as rs
    eng = rs.Engine(0)
    eng.load_eval_key(keyset.bsk, keyset.ksk)
    yield eng
    eng.close()
