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


def _cuda_device_present() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the `gpu` tests are skipped (not errored): the product has no CPU fallback, so there is
    nothing they could run on.  On a GPU box a missing extension still fails loudly inside the engine fixture."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine(keyset):
    import redsec_b200 as rs
    eng = rs.Engine(0)
    eng.load_eval_key(keyset.bsk, keyset.ksk)
    yield eng
    eng.close()
