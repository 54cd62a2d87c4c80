# DRAM / L2 traffic of one 16 384-ciphertext blind-rotate launch under several knob settings (A/B of the BSK stream's L2 behaviour)
mkdir -p gpurun_out
cat > /tmp/traffic_run.py <<PY
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0)
eng.load_eval_key(ks.bsk, ks.ksk)
import os
count = int(os.environ.get("RS_TRAFFIC_COUNT", "16384"))
ct = O.encrypt(np.full(count, 0x20000000), 2.0**-25, ks.lwe_key, 3)
dev = eng.upload(ct); out = eng.alloc(count)
for _ in range(2):
    eng.pbs(dev, 0x20000000, out); eng.sync()
PY
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__t_sector_hit_rate.pct
for setting in "$@"; do
  echo "== $setting"
  env $setting ncu --metrics $M --clock-control none -k regex:"blind_rotate" -s 1 -c 1 --csv python /tmp/traffic_run.py 2>/dev/null | grep -E "blind_rotate" | awk -F'","' '{printf "   %-45s %s %s\n", $13, $15, $14}' | tr -d '"'
done
