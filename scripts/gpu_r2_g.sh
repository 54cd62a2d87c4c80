# Round 2, final evidence pass for the producer-warp blind rotation + tensor-core keyswitch build (1 GPU)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-200 | tail -8 | tee gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py --steps 8 --warmup 3 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 --no-extra > gpurun_out/r2_bench_under_ncu.log 2>&1
bash scripts/gpu_traffic.sh r2_final
timeout 300 python scripts/pbs_sizes.py 1 64 128 148 296 444 592 1024 1536 2048 3072 4096 16384 65536 2>&1 | tee gpurun_out/r2_pbs_sizes.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("sample_bit_exact_vs_gpu"))
for k, v in d.get("extra", {}).get("encrypted_inference", {}).items():
    print(k, v if isinstance(v, str) else (round(v["s_per_image"], 5), v["argmax"]))
PY
