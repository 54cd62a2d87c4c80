"""Debug: block-pipelined max-pool layer (lanes) vs the single-launch form vs the oracle."""
import os, sys
os.environ["RS_POOL_BLOCKS"] = "1"      # the block-pipelined form is opt-in since the end of round 2
import numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import netspec, nets
from oracle import oracle as O, layers_oracle as LO

ks = O.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True, e_bias=2)
spec = dict(name="test/pool_lanes", input=(16, 16, 8), weights=None, image=None,
            layers=[netspec._layer("bin", "conv", 64, "max", "sign", **conv)])
spec["weights"] = netspec.write_random_weights(spec, "/tmp/w_lanes.dat", seed=9, p_zero=0.2, bias_range=3)
rng = np.random.default_rng(10)
bits = rng.integers(0, 2, 16 * 16 * 8) * 2 - 1
ct = O.encrypt((bits * LO.UNIT) & 0xFFFFFFFF, 2.0 ** -25, ks.lwe_key, 47)
net = nets.EncryptedNet(eng, spec)
x = eng.upload(ct)
runs = []
for k in range(3):
    y, _, _ = net.layer_forward(0, x); runs.append(eng.download(y)); y.free()
os.environ["RS_NO_LANES"] = "1"
single = []
for k in range(2):
    y, _, _ = net.layer_forward(0, x); single.append(eng.download(y)); y.free()
del os.environ["RS_NO_LANES"]
print("single deterministic:", np.array_equal(single[0], single[1]))
for k, r in enumerate(runs):
    bad = np.nonzero((r != single[0]).any(axis=1))[0]
    print(f"lanes run {k}: {bad.size} of {r.shape[0]} rows differ from single; first {bad[:12].tolist()}")
    if bad.size:
        oph = bad // (8 * 64); opw = (bad // 64) % 8; c = bad % 64
        print("   oph hist", np.bincount(oph, minlength=8).tolist(), " c range", c.min(), c.max())
L = LO.prepare(spec, spec["weights"])[0]
bad = np.nonzero((runs[0] != single[0]).any(axis=1))[0]
idx = np.unique(np.concatenate([bad[:6], [0, 5, 4095]])).astype(np.int64)
want = LO.enc_layer_rows(L, ct, idx, ks)
print("single == oracle on", idx.tolist(), (single[0][idx] == want).all(axis=1).tolist())
print("lanes  == oracle on", idx.tolist(), (runs[0][idx] == want).all(axis=1).tolist())
dec_s = O.decrypt(single[0], ks.lwe_key, 4096); dec_l = O.decrypt(runs[0], ks.lwe_key, 4096)
print("decrypted bits differ:", int((dec_s != dec_l).sum()), " values lanes", np.unique(dec_l).tolist()[:8])
