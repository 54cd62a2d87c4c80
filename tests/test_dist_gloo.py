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
    # the library's own shard plan (rs_net_shard_plan on a net without a device context) drives a whole sharded forward of the
    # tiny net: per layer, this rank's block of the oracle's layer output -> all-gather -> (interleave) must equal the full output
    dry = nets.EncryptedNet(None, spec)
    x = ct
    modes = []
    for li, Lr in enumerate(layers):
        plan = dry.shard_plan(li, world)
        info = dry.layer_info(li)
        modes.append(plan["mode"])
        lin = LO.enc_linear(Lr, x)                            # bootstrap-free stand-in for the layer (the exchange logic is what is tested)
        if Lr.has_maxpool:                                    # pooled layers exchange AFTER the pool: emulate by taking the window corner
            oh, ow = Lr.mp_geom[4], Lr.mp_geom[5]
            lin = lin.reshape(Lr.q_dims + (351,))[0:2 * oh:2, 0:2 * ow:2].reshape(-1, 351)
        assert lin.shape[0] == info["out_count"]
        ch = info["channels"]
        if plan["mode"] == 0:
            out_l = lin
        elif plan["mode"] == 2:                               # pixel blocks: gathered row blocks are already canonical
            per_r = plan["rows_per_rank"]
            mine_l = lin[rank * per_r:(rank + 1) * per_r].copy()
            g = [torch.empty(mine_l.shape, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(g, torch.from_numpy(mine_l.astype(np.int64)))
            out_l = torch.cat(g).numpy().astype(np.uint32)
        else:                                                 # channel blocks
            cl = plan["c_local"]
            assert cl * world == ch and plan["rows_per_rank"] * world == info["out_count"]
            mine_l = lin.reshape(-1, ch, 351)[:, rank * cl:(rank + 1) * cl].reshape(-1, 351).copy()
            g = [torch.empty(mine_l.shape, dtype=torch.int64) for _ in range(world)]
            dist.all_gather(g, torch.from_numpy(mine_l.astype(np.int64)))
            cat_l = torch.cat(g).numpy().astype(np.uint32)
            out_l = cat_l[nets.interleave_index(lin.shape[0] // ch, cl, world)]
        ok = ok and np.array_equal(out_l, lin)
        x = lin
    ok = ok and modes == [2, 1, 1, 1, 1]                      # input layer by pixel, conv / FC layers by channel (10 % 2 == 0)
    dry.close()
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


def test_shard_plan_of_the_shipped_nets():
    """rs_net_shard_plan (host logic, no device): how every layer of the BASELINE nets splits over 2 / 4 / 8 GPUs -- conv / FC
    layers by output-channel block, conv-less input layers by output pixel, whatever does not divide stays whole."""
    from redsec_b200 import netspec, nets
    cifar = nets.EncryptedNet(None, netspec.NETS["cifar/binarynet"]())
    assert cifar.bootstraps() == 635904
    for world in (2, 4, 8):
        plans = [cifar.shard_plan(i, world) for i in range(cifar.num_layers)]
        assert [p["mode"] for p in plans] == [2, 1, 1, 1, 1, 1, 1, 1, 1, 0 if 10 % world else 1]
        assert plans[1]["rows_per_rank"] == 32 * 32 * 128 // world and plans[1]["c_local"] == 128 // world
        assert plans[2]["rows_per_rank"] == 16 * 16 * 128 // world          # exchanged AFTER the max-pool: 4x fewer ciphertexts
        assert plans[0]["rows_per_rank"] == 3072 // world
    cifar.close()
    mnist = nets.EncryptedNet(None, netspec.NETS["mnist/sign1024x1"]())
    assert [mnist.shard_plan(i, 8)["mode"] for i in range(3)] == [0, 1, 0]     # 196 pooled pixels do not divide by 8; FC10 neither
    assert [mnist.shard_plan(i, 2)["mode"] for i in range(3)] == [2, 1, 1]
    assert [mnist.shard_plan(i, 1)["mode"] for i in range(3)] == [0, 0, 0]
    mnist.close()
