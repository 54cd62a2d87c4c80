"""Client tools over the C-ABI client entry points (redsec_b200/csrc/client.cpp): keygen, encrypt, decrypt, files.

Replaces client/gen_secure_keyset.cpp, client/encrypt_image.cpp, client/decrypt_image.cpp of the reference.
Host-side C++ does the work; this is the ctypes veneer used by bench.py and the tests.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import BSK_WORDS, KSK_WORDS, LWE_N, LWE_WORDS, TLWE_N, RsError

SECALPHA = 2.0 ** -15        # client/encrypt_image.cpp:10
ALPHA_GATE = 2.0 ** -25      # bootsSymEncrypt noise = lwe alpha_min (client/gen_secure_keyset.cpp:77)
UNIT = 1 << 20               # 1/4096
EIGHTH = 1 << 29             # 1/8


class KeySet:
    def __init__(self, lwe_key, tlwe_key, bsk, ksk):
        self.lwe_key, self.tlwe_key, self.bsk, self.ksk = lwe_key, tlwe_key, bsk, ksk


def _chk(rc, what):
    if rc != 0:
        raise RsError(f"{what} failed with code {rc}")


def keygen(seed: int | None = None) -> KeySet:
    """redsec_params_small_v2 keyset (client/gen_secure_keyset.cpp:70-102).  seed=None (default): OS entropy + ChaCha20
    (rs_keygen_secure).  An integer seed gives the DETERMINISTIC test keyset (rs_keygen) -- tests, benchmarks and
    known-answer vectors only; the reference's own keygen seeds its generator with {0,0,0}, i.e. every user's key is the same."""
    lib = _lib.load()
    lwe_key = np.empty(LWE_N, np.int32)
    tlwe_key = np.empty(TLWE_N, np.int32)
    bsk = np.empty(BSK_WORDS, np.uint32)
    ksk = np.empty(KSK_WORDS, np.uint32)
    if seed is None:
        _chk(lib.rs_keygen_secure(lwe_key.ctypes.data, tlwe_key.ctypes.data, bsk.ctypes.data, ksk.ctypes.data), "rs_keygen_secure")
    else:
        _chk(lib.rs_keygen(seed, lwe_key.ctypes.data, tlwe_key.ctypes.data, bsk.ctypes.data, ksk.ctypes.data), "rs_keygen")
    return KeySet(lwe_key, tlwe_key, bsk, ksk)


def encrypt(mu, lwe_key, alpha: float, seed: int | None = None) -> np.ndarray:
    """LWE encryption of torus32 messages.  seed=None: fresh OS entropy per call (rs_lwe_encrypt_secure).  An integer seed is
    for tests only: the same seed gives the same masks, so never encrypt two different messages under one seed."""
    lib = _lib.load()
    mu = np.ascontiguousarray(np.asarray(mu, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32).reshape(-1)
    ct = np.empty((mu.size, LWE_WORDS), np.uint32)
    if seed is None:
        _chk(lib.rs_lwe_encrypt_secure(ct.ctypes.data, mu.ctypes.data, mu.size, alpha, lwe_key.ctypes.data), "rs_lwe_encrypt_secure")
    else:
        _chk(lib.rs_lwe_encrypt(ct.ctypes.data, mu.ctypes.data, mu.size, alpha, lwe_key.ctypes.data, seed), "rs_lwe_encrypt")
    return ct


def encrypt_bits(bits, lwe_key, seed: int | None = None, mu: int = EIGHTH, alpha: float = ALPHA_GATE) -> np.ndarray:
    bits = np.asarray(bits)
    return encrypt(np.where(bits != 0, mu, -mu), lwe_key, alpha, seed)


def encrypt_image(pixels, lwe_key, seed: int | None = None) -> np.ndarray:
    """client/encrypt_image.cpp:76-77: LWE(modSwitchToTorus32(2p-255, 4096), alpha=2^-15) per pixel (all pixels: R1 not reproduced)."""
    v = 2 * np.asarray(pixels, dtype=np.int64) - 255
    return encrypt(v * UNIT, lwe_key, SECALPHA, seed)


def phase(ct, lwe_key) -> np.ndarray:
    lib = _lib.load()
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty(ct.shape[0], np.uint32)
    _chk(lib.rs_lwe_phase(out.ctypes.data, ct.ctypes.data, ct.shape[0], lwe_key.ctypes.data), "rs_lwe_phase")
    return out


def decrypt(ct, lwe_key, msize: int = 4096) -> np.ndarray:
    """client/decrypt_image.cpp:52-58: centred message in (-msize/2, msize/2]."""
    lib = _lib.load()
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty(ct.shape[0], np.int32)
    _chk(lib.rs_lwe_decrypt(out.ctypes.data, ct.ctypes.data, ct.shape[0], lwe_key.ctypes.data, msize), "rs_lwe_decrypt")
    return out


def decrypt_bits(ct, lwe_key) -> np.ndarray:
    ph = phase(ct, lwe_key).astype(np.int64)
    return ((ph < 2 ** 31) & (ph > 0)).astype(np.int64)   # sign of the phase


def write_keys(ks: KeySet, secret_path: str, eval_path: str):
    lib = _lib.load()
    _chk(lib.rs_write_secret_key(secret_path.encode(), ks.lwe_key.ctypes.data, ks.tlwe_key.ctypes.data), "rs_write_secret_key")
    _chk(lib.rs_write_eval_key(eval_path.encode(), ks.bsk.ctypes.data, ks.ksk.ctypes.data), "rs_write_eval_key")


def read_keys(secret_path: str | None, eval_path: str | None) -> KeySet:
    lib = _lib.load()
    lwe_key = tlwe_key = bsk = ksk = None
    if secret_path:
        lwe_key, tlwe_key = np.empty(LWE_N, np.int32), np.empty(TLWE_N, np.int32)
        _chk(lib.rs_read_secret_key(secret_path.encode(), lwe_key.ctypes.data, tlwe_key.ctypes.data), "rs_read_secret_key")
    if eval_path:
        bsk, ksk = np.empty(BSK_WORDS, np.uint32), np.empty(KSK_WORDS, np.uint32)
        _chk(lib.rs_read_eval_key(eval_path.encode(), bsk.ctypes.data, ksk.ctypes.data), "rs_read_eval_key")
    return KeySet(lwe_key, tlwe_key, bsk, ksk)


def write_ctxt(path: str, ct, variance: float = 0.0, append: bool = False):
    lib = _lib.load()
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    _chk(lib.rs_write_ctxt(path.encode(), ct.ctypes.data, ct.shape[0], variance, int(append)), "rs_write_ctxt")


def read_ctxt(path: str, count: int) -> np.ndarray:
    lib = _lib.load()
    ct = np.empty((count, LWE_WORDS), np.uint32)
    _chk(lib.rs_read_ctxt(path.encode(), ct.ctypes.data, count), "rs_read_ctxt")
    return ct
