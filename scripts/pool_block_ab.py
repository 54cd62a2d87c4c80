"""Max-pool layer (sign + 3-OR tree) on one GPU at the sizes one rank sees when cifar/binarynet is sharded over 8 / 2 / 1 GPUs:
block-pipelined on two lanes (RS_POOL_BLOCKS=1) against one launch per tree level (the default).  usage: python scripts/pool_block_ab.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client, netspec, nets

ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True, e_bias=2)
rng = np.random.default_rng(0)
for (h, ch, what) in [(32, 16, "conv2 / 8 GPUs"), (16, 32, "conv4 / 8"), (8, 64, "conv6 / 8"), (32, 64, "conv2 / 2"), (16, 128, "conv4 / 2"), (32, 128, "conv2 / 1")]:
    layers = [netspec._layer("int", "none", 1, "none", "sign"), netspec._layer("bin", "conv", ch, "max", "sign", **conv),
              netspec._layer("bin", "fc_final", 10, "none", "none")]
    spec = dict(name=f"test/pool{h}x{ch}", input=(h, h, 3), layers=layers, weights=f"/tmp/pool_{h}_{ch}.dat", image=None)
    netspec.write_random_weights(spec, spec["weights"], seed=3, p_zero=0.2)
    net = nets.EncryptedNet(eng, spec)
    net.build_tables(0, 1)
    ct = client.encrypt(rng.integers(-100, 100, h * h * 3) * client.UNIT, ks.lwe_key, client.SECALPHA, 5)
    d = eng.upload(ct)
    res = {}
    for mode in ("lanes", "single"):
        if mode == "lanes":
            os.environ["RS_POOL_BLOCKS"] = "1"
        else:
            os.environ.pop("RS_POOL_BLOCKS", None)
        best = None
        for rep in range(3):
            times = []
            out = net.run(d, times=times)
            o = eng.download(out); out.free()
            best = times[1] if best is None else min(best, times[1])
        res[mode] = (best, o)
    os.environ.pop("RS_POOL_BLOCKS", None)
    same = bool(np.array_equal(res["lanes"][1], res["single"][1]))
    n = h * h * ch
    print(f"{what:16s} {n:7d} sign + {3 * n // 4:7d} OR bootstraps: blocks on 2 lanes {res['lanes'][0] * 1e3:8.2f} ms, one launch per level {res['single'][0] * 1e3:8.2f} ms, same ciphertexts {same}", flush=True)
    net.close(); d.free()
