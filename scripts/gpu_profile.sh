# usage: bash scripts/gpu_profile.sh <tag> <variant> <count>   -- one ncu --set full capture of the blind-rotate kernel
TAG=${1:-r1}; V=${2:-0}; COUNT=${3:-592}
mkdir -p gpurun_out
cat > /tmp/prof_run.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0); eng.set_tuning($V)
eng.load_eval_key(ks.bsk, ks.ksk)
ct = O.encrypt(np.full($COUNT, 0x20000000), 2.0**-25, ks.lwe_key, 3)
dev = eng.upload(ct); out = eng.alloc($COUNT)
for _ in range(2):
    eng.pbs(dev, 0x20000000, out); eng.sync()
PY
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -f -o gpurun_out/prof_${TAG} python /tmp/prof_run.py > gpurun_out/prof_${TAG}.log 2>&1
tail -5 gpurun_out/prof_${TAG}.log
