# Quick check of the PBS path on a GPU box: parity tests of the bootstrap, then blind-rotate / keyswitch timings per variant.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests/test_gpu_pbs.py -x -q 2>&1 | tail -25
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/quick_perf.log
import sys, time, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
print("device", eng.device_info())
print("fp64 peak TFLOP/s", eng.fp64_peak_tflops())
eng.load_eval_key(ks.bsk, ks.ksk)
for variant in (0, 1):
    eng.set_tuning(variant)
    for count in (592, 1024, 8192):
        ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
        dev = eng.upload(ct); out = eng.alloc(count)
        eng.pbs(dev, 0x20000000, out); eng.sync()
        eng.profile(True); eng.profile_reset()
        t = time.time(); eng.pbs(dev, 0x20000000, out); eng.sync(); dt = time.time() - t
        br = eng.profile_get(0); ksw = eng.profile_get(1)
        eng.profile(False)
        print(f"variant={variant} count={count}: wall {dt*1e3:.1f} ms -> {count/dt:.0f} PBS/s ; blind_rotate {br[0]:.2f} ms ({count/br[0]*1e3:.0f}/s), keyswitch {ksw[0]:.2f} ms")
PY
