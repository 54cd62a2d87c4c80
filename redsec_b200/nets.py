"""Python harness around the C++ Layer/Net classes (redsec_b200/csrc/layers.cpp) for tests and bench.py.

Mirrors HeBNN::init / HeBNN::run of nets/*/net.cu: build the layer list from a spec (redsec_b200/netspec.py),
prep() reads var_prep.dat, run() executes layer by layer with activations resident on the device.  With
torch.distributed initialised (world > 1) every shardable layer computes only this rank's output-channel block
and the slices are all-gathered over NCCL between layers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import LWE_STRIDE, Engine, LweArray, RsError
from .netspec import ACT, CONV, POOL


class LayerParams(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "conv_win_h", "conv_win_w", "conv_stride_h", "conv_stride_w", "conv_same_pad",
        "pool_win_h", "pool_win_w", "pool_stride_h", "pool_stride_w", "pool_same_pad",
        "e_bias", "shift_bits", "version")]


class _DryEngine:
    """Stands in for an Engine when a net is only prepared and asked for its shapes / shard plan (no device, no context)."""
    ctx = None
    device = -1

    def __init__(self):
        self.lib = _lib.load()

    def _chk(self, rc: int):
        if rc != 0:
            raise RsError(f"redsec_b200 error {rc}")

    def _adopt(self, child):
        pass


class EncryptedNet:
    def __init__(self, eng: Engine | None, spec: dict):
        """eng=None: a dry net (host logic only: layer_info, shard_plan, bootstraps); running it needs an Engine."""
        eng = eng if eng is not None else _DryEngine()
        self.eng, self.spec, self.lib = eng, spec, eng.lib
        self.net = self.lib.rs_net_create(eng.ctx)
        if not self.net:
            raise RsError("rs_net_create failed")
        for ls in spec["layers"]:
            p = LayerParams(ls["conv_win"][0], ls["conv_win"][1], ls["conv_stride"][0], ls["conv_stride"][1], int(ls["conv_same_pad"]),
                            ls["pool_win"][0], ls["pool_win"][1], ls["pool_stride"][0], ls["pool_stride"][1], int(ls["pool_same_pad"]),
                            ls["e_bias"], ls.get("shift_bits") or 1, ls["version"])
            eng._chk(self.lib.rs_net_add_layer(self.net, int(ls["kind"] == "int"), CONV[ls["conv"]], ls["depth"], POOL[ls["pool"]],
                                               ACT[ls["act"]], C.byref(p)))
        h, w, c = spec["input"]
        if "input_scale" in spec:      # the generated net's input tDimensions (nets/mnist/relu1024x1/net.cpp:96-110)
            rc = self.lib.rs_net_prep_ex(self.net, spec["weights"].encode(), h, w, c, spec.get("input_bits", 2),
                                         spec.get("input_up_bound", 2), float(spec["input_scale"]))
        else:
            rc = self.lib.rs_net_prep(self.net, spec["weights"].encode(), h, w, c)
        if rc != 0:
            raise RsError(f"rs_net_prep({spec['weights']}) failed with code {rc} (bad or mismatching weight file)")
        self.num_layers = self.lib.rs_net_num_layers(self.net)
        eng._adopt(self)

    def build_tables(self, rank: int = 0, world: int = 1):
        """Upload every device table of this rank's slices now (otherwise built lazily by the first forward)."""
        self.eng._chk(self.lib.rs_net_build_tables(self.net, rank, world))

    def close(self):
        if self.net:
            self.lib.rs_net_destroy(self.net)
            self.net = None

    def __del__(self):      # a net dropped without close() must still give its context reference back
        try:
            if self.eng.ctx or isinstance(self.eng, _DryEngine):
                self.close()
        except Exception:
            pass

    def layer_info(self, i: int) -> dict:
        oc, ch, bs, oh, ow = C.c_size_t(), C.c_int(), C.c_size_t(), C.c_int(), C.c_int()
        self.eng._chk(self.lib.rs_net_layer_info(self.net, i, C.byref(oc), C.byref(ch), C.byref(bs), C.byref(oh), C.byref(ow)))
        return dict(out_count=oc.value, channels=ch.value, bootstraps=bs.value, out_h=oh.value, out_w=ow.value)

    def shard_plan(self, i: int, world: int) -> dict:
        """rs_net_shard_plan: mode 0 replicated / 1 channel blocks / 2 pixel blocks, rows each rank contributes, channels per block."""
        mode, rows, cl = C.c_int(), C.c_size_t(), C.c_int()
        self.eng._chk(self.lib.rs_net_shard_plan(self.net, i, world, C.byref(mode), C.byref(rows), C.byref(cl)))
        return dict(mode=mode.value, rows_per_rank=rows.value, c_local=cl.value)

    def run_native(self, inp: LweArray, comm=None) -> LweArray:
        """rs_net_run: the whole network inside the library on the engine stream -- with `comm` (engine.Comm) every layer is
        neuron-sharded and all-gathered over NCCL without a host synchronisation.  Does not consume `inp`."""
        out, cnt = C.c_void_p(), C.c_size_t()
        self.eng._chk(self.lib.rs_net_run(self.net, comm.handle if comm is not None else None, inp.ptr, inp.count,
                                          C.byref(out), C.byref(cnt)))
        arr = LweArray(self.eng, cnt.value, out.value)
        arr._owned = True
        return arr

    def bootstraps(self) -> int:
        return sum(self.layer_info(i)["bootstraps"] for i in range(self.num_layers))

    def layer_forward(self, i: int, inp: LweArray, rank: int = 0, world: int = 1):
        """One layer for this rank's channel slice; returns (LweArray rows [pixel][c_local], c0, c1)."""
        out, cnt, c0, c1 = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_int()
        self.eng._chk(self.lib.rs_net_layer_forward(self.net, i, inp.ptr, inp.count, rank, world, C.byref(out), C.byref(cnt),
                                                    C.byref(c0), C.byref(c1)))
        arr = LweArray(self.eng, cnt.value, out.value)
        arr._owned = True     # allocated with rs_lwe_alloc inside the library; freed through rs_lwe_free
        return arr, c0.value, c1.value

    def layer_forward_sharded(self, i: int, inp: LweArray, comm=None) -> LweArray:
        """One layer, neuron-sharded over `comm` (engine.Comm; None = whole layer here): slice forward, NCCL all-gather and
        interleave all inside the library on the engine stream.  Every rank gets the full layer output."""
        out, cnt = C.c_void_p(), C.c_size_t()
        self.eng._chk(self.lib.rs_net_layer_forward_sharded(self.net, i, comm.handle if comm is not None else None, inp.ptr,
                                                            inp.count, C.byref(out), C.byref(cnt)))
        arr = LweArray(self.eng, cnt.value, out.value)
        arr._owned = True
        return arr

    def comm_for(self, dist):
        """engine.Comm for dist = (torch.distributed module, rank, world), created once per engine."""
        if dist is None or dist[2] == 1:
            return None
        comm = getattr(self.eng, "_comm", None)
        if comm is None or comm.handle is None:
            from .engine import Comm
            comm = self.eng._comm = Comm(self.eng, *dist)
        return comm

    def run(self, inp: LweArray, collect: list | None = None, dist=None, times: list | None = None) -> LweArray:
        """HeBNN::run.  dist = (torch.distributed module, rank, world) enables neuron sharding + all-gather (inside the library).
        Without collect / times the whole network is ONE library call (rs_net_run): no host synchronisation between layers.
        collect: receives every layer's output ciphertexts (host); times: host wall-clock seconds per layer (syncs; diagnostics)."""
        import time
        comm = self.comm_for(dist)
        if collect is None and times is None:
            return self.run_native(inp, comm)
        x = inp
        for i in range(self.num_layers):
            t0 = time.perf_counter() if times is not None else 0.0
            y = self.layer_forward_sharded(i, x, comm)
            if collect is not None:
                collect.append(self.eng.download(y))
            if times is not None:
                self.eng.sync()
                times.append(time.perf_counter() - t0)
            if x is not inp:
                x.free()
            x = y
        return x


def shard_range(channels: int, has_conv: bool, rank: int, world: int):
    """rs_shard_range: the output-channel block a rank computes (usable without a GPU)."""
    lib = _lib.load()
    c0, c1 = C.c_int(), C.c_int()
    rc = lib.rs_shard_range(channels, int(has_conv), rank, world, C.byref(c0), C.byref(c1))
    if rc != 0:
        raise RsError(f"rs_shard_range failed with code {rc}")
    return c0.value, c1.value


def interleave_index(pixels: int, c_local: int, world: int) -> np.ndarray:
    """Row permutation applied by rs_lwe_interleave: out row (pix, r*c_local+c) <- gathered row (r, pix, c)."""
    pix, r, c = np.meshgrid(np.arange(pixels), np.arange(world), np.arange(c_local), indexing="ij")
    return ((r * pixels + pix) * c_local + c).reshape(-1)


def _as_tensor(ptr: int, words: int, dev):
    """Zero-copy torch view of library-owned device memory (for torch.distributed collectives)."""
    import torch

    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(h, device=dev)
