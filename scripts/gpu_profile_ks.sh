mkdir -p gpurun_out
cat > /tmp/ks_run.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
rng = np.random.default_rng(1)
ext = rng.integers(0, 2**32, size=(65536, 1025), dtype=np.uint64).astype(np.uint32)
for _ in range(2):
    eng.keyswitch(ext[:1]) if False else None
out = eng.keyswitch(ext)
out = eng.keyswitch(ext)
PY
ncu --set full --clock-control none --import-source on -k regex:keyswitch_tiled -s 1 -c 1 -f -o gpurun_out/prof_ks_final python /tmp/ks_run.py > gpurun_out/prof_ks_final.log 2>&1
tail -3 gpurun_out/prof_ks_final.log
