"""ctypes front-end to oracle/_build/libtfhe_oracle.so -- the CPU oracle.

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from redsec_b200/ (the product).
Parity unpinned vs upstream TFHE v1.1 (un-vendored dependency; see oracle/tfhe_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libtfhe_oracle.so")

n, N, L, KS_T, KS_BASE = 350, 1024, 10, 9, 8
LWE_WORDS = n + 1
BSK_WORDS = n * 2 * L * 2 * N
KSK_WORDS = N * KS_T * KS_BASE * LWE_WORDS

GATE = {"NAND": 0, "OR": 1, "AND": 2, "NOR": 3, "XOR": 4, "XNOR": 5}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "tfhe_oracle.c")
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE], stdout=subprocess.DEVNULL)
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        _lib = C.CDLL(SO)
        _lib.orc_modswitch_to_torus32.restype = C.c_uint32
        _lib.orc_modswitch_to_torus32.argtypes = [C.c_int32, C.c_int32]
        _lib.orc_modswitch_from_torus32.restype = C.c_int32
        _lib.orc_modswitch_from_torus32.argtypes = [C.c_uint32, C.c_int32]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def _p(a, t=C.c_uint32):
    return a.ctypes.data_as(C.POINTER(t))


def to_torus(mu: int, msize: int) -> int:
    return int(lib().orc_modswitch_to_torus32(mu, msize))


def from_torus(phase: int, msize: int) -> int:
    return int(lib().orc_modswitch_from_torus32(int(phase) & 0xFFFFFFFF, msize))


class KeySet:
    """lwe_key[n], tlwe_key[N], bsk[n][2l][2][N] (torus32), ksk[N][t][base][n+1]."""

    def __init__(self, lwe_key, tlwe_key, bsk, ksk):
        self.lwe_key = np.ascontiguousarray(lwe_key, dtype=np.int32)
        self.tlwe_key = np.ascontiguousarray(tlwe_key, dtype=np.int32)
        self.bsk = np.ascontiguousarray(bsk, dtype=np.uint32).reshape(-1)
        self.ksk = np.ascontiguousarray(ksk, dtype=np.uint32).reshape(-1)
        assert self.bsk.size == BSK_WORDS and self.ksk.size == KSK_WORDS
        self._fft = None

    @property
    def bsk_fft(self):
        if self._fft is None:
            self._fft = np.empty(BSK_WORDS, dtype=np.float64)
            lib().orc_bsk_to_fft(_p(self.bsk), _p(self._fft, C.c_double))
        return self._fft


def keygen(seed: int) -> KeySet:
    lwe_key = np.empty(n, np.int32)
    tlwe_key = np.empty(N, np.int32)
    bsk = np.empty(BSK_WORDS, np.uint32)
    ksk = np.empty(KSK_WORDS, np.uint32)
    lib().orc_keygen(C.c_uint64(seed), _p(lwe_key, C.c_int32), _p(tlwe_key, C.c_int32), _p(bsk), _p(ksk))
    return KeySet(lwe_key, tlwe_key, bsk, ksk)


def encrypt(mu, alpha: float, lwe_key, seed: int):
    mu = np.ascontiguousarray(np.asarray(mu, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32).reshape(-1)
    ct = np.empty((mu.size, LWE_WORDS), np.uint32)
    key = np.ascontiguousarray(lwe_key, dtype=np.int32)
    lib().orc_lwe_encrypt(_p(ct), _p(mu), C.c_int(mu.size), C.c_double(alpha), _p(key, C.c_int32), C.c_uint64(seed))
    return ct


def phase(ct, lwe_key):
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty(ct.shape[0], np.uint32)
    key = np.ascontiguousarray(lwe_key, dtype=np.int32)
    lib().orc_lwe_phase(_p(out), _p(ct), C.c_int(ct.shape[0]), _p(key, C.c_int32))
    return out


def decrypt(ct, lwe_key, msize: int):
    """lweSymDecrypt + modSwitchFromTorus32, centred to (-msize/2, msize/2] (client/decrypt_image.cpp:52-58)."""
    ph = phase(ct, lwe_key).astype(np.uint64)
    interv = ((1 << 63) // msize) * 2
    v = (((ph << np.uint64(32)) + np.uint64(interv // 2)) & np.uint64(0xFFFFFFFFFFFFFFFF)) // np.uint64(interv)
    v = v.astype(np.int64) % msize
    return np.where(v > msize // 2, v - msize, v)


def blind_rotate(lwe_in, mu: int, ks: KeySet, exact: bool, stats=None):
    lwe_in = np.ascontiguousarray(lwe_in, dtype=np.uint32).reshape(LWE_WORDS)
    acc = np.empty(2 * N, np.uint32)
    if exact:
        lib().orc_blind_rotate_exact(_p(acc), _p(lwe_in), C.c_uint32(mu), _p(ks.bsk))
    else:
        st = _p(stats, C.c_double) if stats is not None else None
        lib().orc_blind_rotate_fft(_p(acc), _p(lwe_in), C.c_uint32(mu), _p(ks.bsk_fft, C.c_double), st)
    return acc


def sample_extract(acc):
    acc = np.ascontiguousarray(acc, dtype=np.uint32).reshape(2 * N)
    ext = np.empty(N + 1, np.uint32)
    lib().orc_sample_extract(_p(ext), _p(acc))
    return ext


def keyswitch(ext, ks: KeySet):
    ext = np.ascontiguousarray(ext, dtype=np.uint32).reshape(-1, N + 1)
    out = np.empty((ext.shape[0], LWE_WORDS), np.uint32)
    for i in range(ext.shape[0]):
        lib().orc_keyswitch(_p(out[i]), _p(ext[i]), _p(ks.ksk))
    return out


def pbs(ct, mu: int, ks: KeySet, exact: bool = False, threads: int = 0, stats=None):
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty_like(ct)
    st = _p(stats, C.c_double) if stats is not None else None
    lib().orc_pbs_batch(_p(out), _p(ct), C.c_int(ct.shape[0]), C.c_uint32(mu), _p(ks.bsk), _p(ks.bsk_fft, C.c_double),
                        _p(ks.ksk), C.c_int(int(exact)), C.c_int(threads), st)
    return out


def pbs_lut(ct, luts, ks: KeySet, exact: bool = False, threads: int = 0):
    """Programmable bootstrap with per-ciphertext test vectors: ciphertext c uses luts[c % len(luts)] (row f4)."""
    ct = np.ascontiguousarray(ct, dtype=np.uint32).reshape(-1, LWE_WORDS)
    luts = np.ascontiguousarray(luts, dtype=np.uint32).reshape(-1, N)
    out = np.empty_like(ct)
    lib().orc_pbs_lut_batch(_p(out), _p(ct), C.c_int(ct.shape[0]), _p(luts), C.c_int(luts.shape[0]), _p(ks.bsk),
                            _p(ks.bsk_fft, C.c_double), _p(ks.ksk), C.c_int(int(exact)), C.c_int(threads))
    return out


def gate_linear(op: str, a, b):
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, LWE_WORDS)
    b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty_like(a)
    lib().orc_gate_linear(C.c_int(GATE[op]), _p(out), _p(a), _p(b), C.c_int(a.shape[0]))
    return out


def gate(op: str, a, b, mu: int, ks: KeySet, exact: bool = False, threads: int = 0):
    a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, LWE_WORDS)
    b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, LWE_WORDS)
    out = np.empty_like(a)
    lib().orc_gate_batch(C.c_int(GATE[op]), _p(out), _p(a), _p(b), C.c_int(a.shape[0]), C.c_uint32(mu), _p(ks.bsk),
                         _p(ks.bsk_fft, C.c_double), _p(ks.ksk), C.c_int(int(exact)), C.c_int(threads))
    return out


def lincomb(inp, rowptr, col, sign, bias=None):
    inp = np.ascontiguousarray(inp, dtype=np.uint32).reshape(-1, LWE_WORDS)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    sign = np.ascontiguousarray(sign, dtype=np.int8)
    out_count = rowptr.size - 1
    out = np.empty((out_count, LWE_WORDS), np.uint32)
    b = None
    if bias is not None:
        bias = np.ascontiguousarray(np.asarray(bias, dtype=np.int64) & 0xFFFFFFFF, dtype=np.uint32)
        b = _p(bias)
    lib().orc_lwe_lincomb(_p(out), C.c_int(out_count), _p(inp), _p(rowptr, C.c_int32), _p(col, C.c_int32),
                          _p(sign, C.c_int8), b)
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
