# final pass of round 2: GPU tests + the N=1 bench line (+ reference arm) on the committed build
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-200 | tail -8 | tee gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py 2> gpurun_out/r2_bench_n1.err | tail -1 > gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/r2_bench_ref.err | tail -1 > gpurun_out/r2_bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print("steps", d["steps"], "value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"].get("sample_bit_exact_vs_gpu"))
for k, v in d.get("extra", {}).get("encrypted_inference", {}).items():
    print(k, v if isinstance(v, str) else (round(v["s_per_image"], 5), v["argmax"]))
print(open("gpurun_out/r2_bench_reference_arm.json").read()[:600])
PY
