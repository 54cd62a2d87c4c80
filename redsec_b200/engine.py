"""Thin Python host over the C-ABI (include/redsec_b200.h).

Python is only the test/bench harness and the torch.distributed plumbing; all compute is in
libredsec_b200.so (CUDA, sm_100a).  Nothing here imports oracle/.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

LWE_N, LWE_WORDS, LWE_STRIDE, TLWE_N, EXT_STRIDE = 350, 351, 352, 1024, 1028
BSK_WORDS = LWE_N * 20 * 2 * TLWE_N
KSK_WORDS = TLWE_N * 9 * 8 * LWE_WORDS
GATE = {"NAND": 0, "OR": 1, "AND": 2, "NOR": 3, "XOR": 4, "XNOR": 5}
K_BLIND_ROTATE, K_KEYSWITCH, K_LINEAR, K_OTHER = 0, 1, 2, 3


def torus(mu: int, msize: int) -> int:
    """modSwitchToTorus32(mu, msize) (TFHE; lib/BinOps_enc.cpp:184)."""
    interv = ((1 << 63) // msize) * 2
    return ((mu * interv) & 0xFFFFFFFFFFFFFFFF) >> 32


class RsError(RuntimeError):
    pass


class LweArray:
    """count LWE samples resident on the device, rows of 352 words."""

    def __init__(self, eng: "Engine", count: int, ptr: int | None = None, owner=None):
        self.eng, self.count, self.owner = eng, int(count), owner
        if ptr is None:
            p = C.c_void_p()
            eng._chk(eng.lib.rs_lwe_alloc(eng.ctx, self.count, C.byref(p)))
            self.ptr, self._owned = p.value, True
        else:
            self.ptr, self._owned = int(ptr), False

    def slice(self, start: int, count: int) -> "LweArray":
        assert 0 <= start and start + count <= self.count
        return LweArray(self.eng, count, self.ptr + start * LWE_STRIDE * 4, owner=self)

    def free(self):
        if self._owned and self.ptr:
            self.eng.lib.rs_lwe_free(self.eng.ctx, self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Engine:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        ctx = C.c_void_p()
        rc = self.lib.rs_ctx_create(C.byref(ctx), device)
        if rc != 0:
            raise RsError(f"rs_ctx_create failed ({rc}): {self.lib.rs_last_error(None).decode()}")
        self.ctx = ctx
        self.device = device
        self._children = []       # weak references to nets / communicators created on this context (closed before it)

    def _adopt(self, child):
        import weakref
        self._children.append(weakref.ref(child))

    def close(self):
        if self.ctx:
            # layers, nets and communicators free their device tables through the context: close them first
            # (rs_ctx_destroy refuses while any is alive)
            for ref in self._children:
                child = ref()
                if child is not None:
                    child.close()
            self._children = []
            rc = self.lib.rs_ctx_destroy(self.ctx)
            if rc != 0:
                raise RsError(f"rs_ctx_destroy failed ({rc}): {self.lib.rs_last_error(self.ctx).decode()}")
            self.ctx = None

    def _chk(self, rc: int):
        if rc != 0:
            raise RsError(f"redsec_b200 error {rc}: {self.lib.rs_last_error(self.ctx).decode()}")

    # ---- keys / data movement
    def load_eval_key(self, bsk: np.ndarray, ksk: np.ndarray):
        bsk = np.ascontiguousarray(bsk, dtype=np.uint32).reshape(-1)
        ksk = np.ascontiguousarray(ksk, dtype=np.uint32).reshape(-1)
        assert bsk.size == BSK_WORDS and ksk.size == KSK_WORDS
        self._chk(self.lib.rs_load_eval_key(self.ctx, bsk.ctypes.data_as(_lib.u32p), ksk.ctypes.data_as(_lib.u32p)))

    def alloc(self, count: int) -> LweArray:
        return LweArray(self, count)

    def upload(self, host: np.ndarray, out: LweArray | None = None) -> LweArray:
        host = np.ascontiguousarray(host, dtype=np.uint32).reshape(-1, LWE_WORDS)
        out = out or self.alloc(host.shape[0])
        self._chk(self.lib.rs_lwe_upload(self.ctx, out.ptr, host.ctypes.data, host.shape[0]))
        self.sync()
        return out

    def download(self, arr: LweArray) -> np.ndarray:
        host = np.empty((arr.count, LWE_WORDS), np.uint32)
        self._chk(self.lib.rs_lwe_download(self.ctx, host.ctypes.data, arr.ptr, arr.count))
        return host

    def sync(self):
        self._chk(self.lib.rs_sync(self.ctx))

    def set_stream(self, cuda_stream: int | None):
        self._chk(self.lib.rs_set_stream(self.ctx, cuda_stream))

    def stream_handle(self) -> int:
        """cudaStream_t of the selected lane (rs_get_stream), e.g. for torch.cuda.ExternalStream."""
        h = C.c_void_p()
        self._chk(self.lib.rs_get_stream(self.ctx, C.byref(h)))
        return h.value or 0

    def pool_trim(self):
        self._chk(self.lib.rs_pool_trim(self.ctx))

    # ---- lanes (extra streams with their own bootstrap scratch)
    def lanes(self, n: int):
        self._chk(self.lib.rs_lanes(self.ctx, n))

    def lane_select(self, k: int):
        self._chk(self.lib.rs_lane_select(self.ctx, k))

    def lane_fork(self):
        self._chk(self.lib.rs_lane_fork(self.ctx))

    def lane_join(self):
        self._chk(self.lib.rs_lane_join(self.ctx))

    # ---- hot path
    def pbs(self, inp: LweArray, mu: int, out: LweArray | None = None) -> LweArray:
        out = out or self.alloc(inp.count)
        self._chk(self.lib.rs_pbs_batch(self.ctx, out.ptr, inp.ptr, inp.count, mu & 0xFFFFFFFF))
        return out

    def pbs_lut(self, inp: LweArray, luts: np.ndarray, out: LweArray | None = None) -> LweArray:
        """rs_pbs_lut_batch: ciphertext c is bootstrapped with test vector luts[c % len(luts)] (uint32 [m][1024])."""
        luts = np.ascontiguousarray(luts, dtype=np.uint32).reshape(-1, TLWE_N)
        out = out or self.alloc(inp.count)
        dev = C.c_void_p()
        self._chk(self.lib.rs_dev_alloc(self.ctx, luts.nbytes, C.byref(dev)))
        try:
            self._chk(self.lib.rs_dev_upload(self.ctx, dev.value, luts.ctypes.data, luts.nbytes))
            self._chk(self.lib.rs_pbs_lut_batch(self.ctx, out.ptr, inp.ptr, inp.count, dev.value, luts.shape[0]))
            self.sync()
        finally:
            self.lib.rs_dev_free(self.ctx, dev.value)
        return out

    def add_const(self, arr: LweArray, value: int) -> LweArray:
        self._chk(self.lib.rs_lwe_add_const(self.ctx, arr.ptr, arr.count, value & 0xFFFFFFFF))
        return arr

    def gate(self, op: str, a: LweArray, b: LweArray, mu: int, out: LweArray | None = None) -> LweArray:
        out = out or self.alloc(a.count)
        self._chk(self.lib.rs_gate_batch(self.ctx, GATE[op], out.ptr, a.ptr, b.ptr, a.count, mu & 0xFFFFFFFF))
        return out

    def pbs_host(self, host_in: np.ndarray, mu: int, host_out: np.ndarray | None = None) -> np.ndarray:
        count = host_in.size // LWE_WORDS
        if host_out is None:
            host_out = np.empty((count, LWE_WORDS), np.uint32)
        self._chk(self.lib.rs_pbs_batch_host(self.ctx, host_out.ctypes.data, host_in.ctypes.data, count, mu & 0xFFFFFFFF))
        return host_out

    def gate_host(self, op: str, a: np.ndarray, b: np.ndarray, mu: int, host_out: np.ndarray | None = None) -> np.ndarray:
        count = a.size // LWE_WORDS
        if host_out is None:
            host_out = np.empty((count, LWE_WORDS), np.uint32)
        self._chk(self.lib.rs_gate_batch_host(self.ctx, GATE[op], host_out.ctypes.data, a.ctypes.data, b.ctypes.data, count,
                                              mu & 0xFFFFFFFF))
        return host_out

    def blind_rotate(self, inp: LweArray, mu: int) -> np.ndarray:
        """Blind rotation + sample extract only; returns host [count][1025] (a'[0..1023], b')."""
        p = C.c_void_p()
        self._chk(self.lib.rs_ext_alloc(self.ctx, inp.count, C.byref(p)))
        try:
            self._chk(self.lib.rs_blind_rotate_batch(self.ctx, p.value, inp.ptr, inp.count, mu & 0xFFFFFFFF))
            host = np.empty((inp.count, TLWE_N + 1), np.uint32)
            self._chk(self.lib.rs_ext_download(self.ctx, host.ctypes.data, p.value, inp.count))
        finally:
            self.lib.rs_dev_free(self.ctx, p.value)
        return host

    def keyswitch(self, ext_host: np.ndarray) -> np.ndarray:
        ext_host = np.ascontiguousarray(ext_host, dtype=np.uint32).reshape(-1, TLWE_N + 1)
        count = ext_host.shape[0]
        p = C.c_void_p()
        self._chk(self.lib.rs_ext_alloc(self.ctx, count, C.byref(p)))
        out = self.alloc(count)
        try:
            self._chk(self.lib.rs_ext_upload(self.ctx, p.value, ext_host.ctypes.data, count))
            self._chk(self.lib.rs_keyswitch_batch(self.ctx, out.ptr, p.value, count))
            return self.download(out)
        finally:
            self.lib.rs_dev_free(self.ctx, p.value)

    # ---- device buffers for layer tables
    def dev_upload(self, host: np.ndarray) -> int:
        host = np.ascontiguousarray(host)
        p = C.c_void_p()
        self._chk(self.lib.rs_dev_alloc(self.ctx, max(host.nbytes, 1), C.byref(p)))
        if host.nbytes:
            self._chk(self.lib.rs_dev_upload(self.ctx, p.value, host.ctypes.data, host.nbytes))
        return p.value

    def dev_free(self, ptr: int):
        self.lib.rs_dev_free(self.ctx, ptr)

    def lincomb(self, out: LweArray, inp: LweArray, rowptr_dev: int, col_dev: int, sign_dev: int, bias_dev: int | None):
        self._chk(self.lib.rs_lwe_lincomb(self.ctx, out.ptr, out.count, inp.ptr, rowptr_dev, col_dev, sign_dev, bias_dev))

    # ---- measurement
    def profile(self, on: bool):
        self._chk(self.lib.rs_profile_enable(self.ctx, int(on)))

    def profile_reset(self):
        self._chk(self.lib.rs_profile_reset(self.ctx))

    def profile_get(self, kind: int):
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self.lib.rs_profile_get(self.ctx, kind, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self.lib.rs_launch_count(self.ctx))

    def fp64_peak_tflops(self) -> float:
        v = C.c_double()
        self._chk(self.lib.rs_fp64_peak(self.ctx, C.byref(v)))
        return v.value

    def fp64_peak_three_operand_tflops(self) -> float:
        v = C.c_double()
        self._chk(self.lib.rs_fp64_peak_three_operand(self.ctx, C.byref(v)))
        return v.value

    def set_tuning(self, br_variant: int):
        self._chk(self.lib.rs_set_tuning(self.ctx, br_variant))

    def set_ks_variant(self, ks_variant: int):
        self._chk(self.lib.rs_set_ks_variant(self.ctx, ks_variant))

    def device_info(self):
        sm, ma, mi, sh = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._chk(self.lib.rs_device_info(self.ctx, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(sh)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "smem_optin": sh.value}


class Comm:
    """NCCL communicator of one rank (rs_comm_init_rank).  `dist` is an initialised torch.distributed module: it is used once,
    to hand rank 0's NCCL id to the other ranks; every collective afterwards is issued by the library on the engine stream."""

    def __init__(self, eng: Engine, dist, rank: int, world: int):
        import torch
        self.eng, self.lib, self.rank, self.world = eng, eng.lib, rank, world
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            rc = self.lib.rs_comm_unique_id(ident)
            if rc != 0:
                raise RsError(f"rs_comm_unique_id failed ({rc}): {self.lib.rs_comm_last_error().decode()}")
        backend = dist.get_backend()
        dev = torch.device("cuda", eng.device) if backend == "nccl" else torch.device("cpu")
        t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        h = C.c_void_p()
        rc = self.lib.rs_comm_init_rank(eng.ctx, C.byref(h), ident, rank, world)
        if rc != 0:
            raise RsError(f"rs_comm_init_rank failed ({rc}): {self.lib.rs_comm_last_error().decode()}")
        self.handle = h
        eng._adopt(self)

    def allgather(self, arr: LweArray) -> LweArray:
        out = self.eng.alloc(arr.count * self.world)
        rc = self.lib.rs_allgather(self.handle, out.ptr, arr.ptr, arr.count)
        if rc != 0:
            raise RsError(f"rs_allgather failed ({rc}): {self.lib.rs_comm_last_error().decode()}")
        return out

    def close(self):
        if self.handle:
            self.lib.rs_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.eng.ctx:
                self.close()
        except Exception:
            pass
