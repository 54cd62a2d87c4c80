# usage: bash scripts/gpu_variants.sh "0 3 4"  -- parity of the blind-rotate variants, then blind-rotate / keyswitch timings per variant
VARS=${1:-"0 3 4"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pbs.py -x -q 2>&1 | tail -15
RS_VARS="$VARS" timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/variants_perf.log
import os, sys, time, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
eng.load_eval_key(ks.bsk, ks.ksk)
for variant in [int(v) for v in os.environ["RS_VARS"].split()]:
    eng.set_tuning(variant)
    for count in (148, 592, 1024, 65536):
        ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
        dev = eng.upload(ct); out = eng.alloc(count)
        eng.pbs(dev, 0x20000000, out); eng.sync()
        eng.profile(True); eng.profile_reset()
        eng.pbs(dev, 0x20000000, out); eng.sync()
        br = eng.profile_get(0); ksw = eng.profile_get(1)
        eng.profile(False)
        print(f"variant={variant} count={count}: blind_rotate {br[0]:.2f} ms ({count/br[0]*1e3:.0f}/s), keyswitch {ksw[0]:.2f} ms", flush=True)
        dev.free(); out.free()
PY
