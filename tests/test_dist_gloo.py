"""CPU, world_size 2 over gloo: the N>1 host-side logic of a sharded layer -- channel partition (rs_shard_range via the
C-ABI), all-gather of per-rank slices, and the row permutation rs_lwe_interleave applies -- with the CPU oracle standing
in for the CUDA compute.  The gathered result must equal the unsharded layer ciphertext for ciphertext."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import layers_oracle as LO, oracle as O
    from redsec_b200 import netspec, nets

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = netspec.tiny_cifar_like()
    spec["weights"] = os.path.join(tmpdir, "w.dat")
    layers = LO.prepare(spec, spec["weights"])
    rng = np.random.default_rng(5)
    ct = rng.integers(0, 2 ** 32, size=(8 * 8 * 3, 351), dtype=np.uint64).astype(np.uint32)   # same on every rank
    L = layers[1]                                            # 3x3 conv, 16 channels, same padding
    full = LO.enc_linear(L, ct)                              # [(h,w,c)][351]
    c0, c1 = nets.shard_range(16, True, rank, world)
    assert (c1 - c0) * world == 16
    # this rank's slice: rows [pixel][c_local]
    mine = full.reshape(64, 16, 351)[:, c0:c1, :].reshape(-1, 351).copy()
    gathered = [torch.empty(mine.shape, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(mine.astype(np.int64)))
    cat = torch.cat(gathered).numpy().astype(np.uint32)       # [world][pixels][c_local]
    out = cat[nets.interleave_index(64, c1 - c0, world)]
    ok = np.array_equal(out, full)
    # a conv-less input layer (3 channels) keeps its full channel range: it is sharded by output pixel instead, and the
    # all-gather of the per-rank row blocks [pixels/world][channels] is already in canonical (h,w,c) order
    ok = ok and nets.shard_range(3, False, rank, world) == (0, 3)
    L0 = layers[0]
    full0 = LO.enc_linear(L0, ct)                            # [(h,w,c)][351], 64 pixels x 3 channels
    per = 64 // world
    mine0 = full0.reshape(64, 3, 351)[rank * per:(rank + 1) * per].reshape(-1, 351).copy()
    g0 = [torch.empty(mine0.shape, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(g0, torch.from_numpy(mine0.astype(np.int64)))
    cat0 = torch.cat(g0).numpy().astype(np.uint32)
    ok = ok and np.array_equal(cat0[nets.interleave_index(64, 3, 1)], full0)
    flag = torch.tensor([int(ok)])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if not int(flag):
        raise SystemExit(1)


def test_sharded_layer_gather_world2(tmp_path):
    import torch.multiprocessing as mp
    from redsec_b200 import netspec
    spec = netspec.tiny_cifar_like()
    netspec.write_random_weights(spec, str(tmp_path / "w.dat"), seed=9, p_zero=0.2)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
