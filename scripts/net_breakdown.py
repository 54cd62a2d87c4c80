"""Per-kernel-kind time of one encrypted inference (rs_profile_*): usage python scripts/net_breakdown.py [net]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client, netspec, nets
name = sys.argv[1] if len(sys.argv) > 1 else "cifar/binarynet"
ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
spec = netspec.NETS[name]()
label, px = netspec.load_image_csv(spec["image"])
ct = client.encrypt_image(px, ks.lwe_key, seed=7)
net = nets.EncryptedNet(eng, spec)
net.build_tables()
d = eng.upload(ct)
if name.startswith("mnist"):
    net.run(d).free(); eng.sync()
eng.profile(True); eng.profile_reset()
t0 = time.perf_counter(); out = net.run(d); eng.sync(); dt = time.perf_counter() - t0
kinds = ["blind_rotate", "keyswitch", "linear", "other"]
tot = 0
for k, n in enumerate(kinds):
    ms, cnt = eng.profile_get(k); tot += ms
    print(f"{n:14s} {ms:10.2f} ms in {cnt} launches")
print(f"sum of kernels {tot:.2f} ms, wall {dt*1e3:.2f} ms, bootstraps {net.bootstraps()} -> {net.bootstraps()/dt:.0f}/s")
for i in range(net.num_layers):
    print(i, net.layer_info(i))
# per layer: host wall clock (sync after the layer) against the sum of its kernels
x = d
for i in range(net.num_layers):
    eng.sync(); eng.profile_reset()
    t0 = time.perf_counter(); y, _, _ = net.layer_forward(i, x); eng.sync(); dt = time.perf_counter() - t0
    ks_ms = [eng.profile_get(k)[0] for k in range(4)]
    print(f"layer {i}: wall {dt*1e3:9.2f} ms, kernels {sum(ks_ms):9.2f} ms (blind rotate {ks_ms[0]:.2f}, keyswitch {ks_ms[1]:.2f}, linear {ks_ms[2]:.2f}), host gap {dt*1e3-sum(ks_ms):7.2f} ms")
    x = y
