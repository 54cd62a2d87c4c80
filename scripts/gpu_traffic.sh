# usage: bash scripts/gpu_traffic.sh <tag>  -- DRAM bytes of ONE full-size (2^16 ciphertexts) blind-rotate launch and one keyswitch launch
TAG=${1:-r1}
mkdir -p gpurun_out
cat > /tmp/traffic_run.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
eng.load_eval_key(ks.bsk, ks.ksk)
count = 65536
ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
dev = eng.upload(ct); out = eng.alloc(count)
for _ in range(2):
    eng.pbs(dev, 0x20000000, out); eng.sync()
PY
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none -k regex:"blind_rotate|keyswitch_mma|keyswitch_tiled" -s 2 -c 2 --csv --log-file gpurun_out/traffic_${TAG}.csv python /tmp/traffic_run.py > gpurun_out/traffic_${TAG}.log 2>&1
tail -3 gpurun_out/traffic_${TAG}.log; cat gpurun_out/traffic_${TAG}.csv | tail -12
