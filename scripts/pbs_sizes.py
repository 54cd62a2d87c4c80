"""Time rs_pbs_batch at several batch sizes (device-resident, CUDA-synchronised wall clock): looks for wave-quantisation or
desynchronisation anomalies.  usage: python scripts/pbs_sizes.py [count ...]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from redsec_b200 import client
counts = [int(a) for a in sys.argv[1:]] or [148, 592, 1184, 1536, 1776, 3072, 4096, 16384]
ks = client.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
rng = np.random.default_rng(0)
for n in counts:
    ct = client.encrypt(rng.integers(-500, 500, n) * client.UNIT, ks.lwe_key, client.SECALPHA, 3)
    d = eng.upload(ct); out = eng.alloc(n)
    eng.pbs(d, client.UNIT, out); eng.sync()
    eng.profile(True); eng.profile_reset()
    t0 = time.perf_counter()
    for _ in range(3):
        eng.pbs(d, client.UNIT, out)
    eng.sync(); dt = (time.perf_counter() - t0) / 3
    br, _ = eng.profile_get(0); ksw, _ = eng.profile_get(1)
    eng.profile(False)
    print(f"count {n:6d}: {dt*1e3:8.2f} ms wall  blind-rotate {br/3:8.2f} ms  keyswitch {ksw/3:6.2f} ms  -> {n/dt:9.0f} PBS/s  ({n/592:.2f} waves of 592)")
    d.free(); out.free()
