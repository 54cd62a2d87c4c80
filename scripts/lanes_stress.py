"""Stress: block-pipelined max-pool layer (lanes) vs the single-launch form, repeated; reports where rows differ."""
import os, sys
os.environ["RS_POOL_BLOCKS"] = "1"      # the block-pipelined form is opt-in since the end of round 2
import numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import netspec, nets
from oracle import oracle as O, layers_oracle as LO

H, W, CIN, COUT = [int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (32, 32, 8, 32))]
REPS = int(sys.argv[5]) if len(sys.argv) > 5 else 12
ks = O.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
conv = dict(conv_win=(3, 3), conv_stride=(1, 1), conv_same_pad=True, e_bias=2)
spec = dict(name="test/pool_lanes", input=(H, W, CIN), weights=None, image=None,
            layers=[netspec._layer("bin", "conv", COUT, "max", "sign", **conv)])
spec["weights"] = netspec.write_random_weights(spec, "/tmp/w_lanes.dat", seed=9, p_zero=0.2, bias_range=3)
rng = np.random.default_rng(10)
bits = rng.integers(0, 2, H * W * CIN) * 2 - 1
ct = O.encrypt((bits * LO.UNIT) & 0xFFFFFFFF, 2.0 ** -25, ks.lwe_key, 47)
net = nets.EncryptedNet(eng, spec)
x = eng.upload(ct)
os.environ["RS_NO_LANES"] = "1"
y, _, _ = net.layer_forward(0, x); ref = eng.download(y); y.free()
y, _, _ = net.layer_forward(0, x); ref2 = eng.download(y); y.free()
del os.environ["RS_NO_LANES"]
print("single-launch form deterministic:", np.array_equal(ref, ref2), " rows", ref.shape[0])
dec_ref = O.decrypt(ref, ks.lwe_key, 4096)
bad_total = 0
for k in range(REPS):
    y, _, _ = net.layer_forward(0, x); got = eng.download(y); y.free()
    bad = np.nonzero((got != ref).any(axis=1))[0]
    bad_total += bad.size
    if bad.size:
        oph = bad // ((W // 2) * COUT)
        dec = O.decrypt(got[bad], ks.lwe_key, 4096)
        nwords = (got[bad] != ref[bad]).sum(axis=1)
        for r in bad[:4]:
            w = np.nonzero(got[r] != ref[r])[0]
            print(f"   row {r} (block-local {r % ((W // 2) * COUT)}): differing words {w.min()}..{w.max()} ({w.size}); contiguous {bool(w.size == w.max() - w.min() + 1)}")
        print(f"run {k}: {bad.size} rows differ; oph values {sorted(set(oph.tolist()))[:10]}; first rows {bad[:8].tolist()}; "
              f"words differing per row min/max {nwords.min()}/{nwords.max()}; decrypt {dec[:8].tolist()} vs ref {dec_ref[bad][:8].tolist()}")
print("lanes in use:", eng.lib.rs_lane_count(eng.ctx), " total differing rows over", REPS, "runs:", bad_total)
