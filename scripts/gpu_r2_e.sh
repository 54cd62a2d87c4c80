set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-200 | tail -8 | tee gpurun_out/r2e_gpu_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 2> gpurun_out/r2e_bench_n1.err | tail -1 > gpurun_out/r2e_bench_n1.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "ks", d["roofline"]["keyswitch"]["launch_ms"], d["roofline"]["keyswitch"]["achieved_tops"], d["cpu_baseline"].get("sample_bit_exact_vs_gpu"))
for k, v in d.get("extra", {}).get("encrypted_inference", {}).items():
    print(k, v if isinstance(v, str) else (round(v["s_per_image"], 5), v["argmax"]))
PY
cat > /tmp/ks_prof.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
import redsec_b200 as rs
from oracle import oracle as O
ks = O.keygen(0)
eng = rs.Engine(0); eng.load_eval_key(ks.bsk, ks.ksk)
rng = np.random.default_rng(1)
ext = rng.integers(0, 2 ** 32, size=(16384, 1025), dtype=np.uint64).astype(np.uint32)
for _ in range(2):
    eng.keyswitch(ext)
PY
ncu --set full --clock-control none --import-source on -k regex:keyswitch_mma -s 1 -c 1 -f -o gpurun_out/prof_r2_ksmma python /tmp/ks_prof.py > gpurun_out/prof_r2_ksmma.log 2>&1
ncu -i gpurun_out/prof_r2_ksmma.ncu-rep --page raw --csv > gpurun_out/prof_r2_ksmma_raw.csv 2>/dev/null
tail -3 gpurun_out/prof_r2_ksmma.log
