"""redsec_b200 -- B200-native TFHE bootstrap engine behind REDsec's lib/ API (hot path only).

The product is redsec_b200/libredsec_b200.so (CUDA for sm_100a + C-ABI, include/redsec_b200.h).
This package is the Python harness over it; importing it does not need a GPU, creating an Engine does.
"""
from . import _lib  # noqa: F401
from .engine import Engine, LweArray, RsError, torus  # noqa: F401
