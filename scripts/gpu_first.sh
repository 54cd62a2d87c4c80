set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_gpu_pbs.py -x -q 2>&1 | tail -25
timeout 300 python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
print("device", eng.device_info())
print("fp64 peak TFLOP/s", eng.fp64_peak_tflops())
eng.load_eval_key(ks.bsk, ks.ksk)
rng = np.random.default_rng(1)
for groups in (4, 6):
    eng.set_tuning(groups)
    for count in (888, 4096):
        ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
        dev = eng.upload(ct); out = eng.alloc(count)
        eng.pbs(dev, 0x20000000, out); eng.sync()
        eng.profile(True); eng.profile_reset()
        t = time.time(); eng.pbs(dev, 0x20000000, out); eng.sync(); dt = time.time() - t
        br = eng.profile_get(0); ksw = eng.profile_get(1)
        eng.profile(False)
        print(f"groups={groups} count={count}: wall {dt*1e3:.1f} ms -> {count/dt:.0f} PBS/s ; blind_rotate {br[0]:.2f} ms, keyswitch {ksw[0]:.2f} ms")
PY
