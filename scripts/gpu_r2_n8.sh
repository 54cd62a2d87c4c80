set -x
mkdir -p gpurun_out
N=8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $RUN bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/r2_bench_n$N.err | tail -1 > gpurun_out/r2_bench_n$N.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n$N.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"])
for k, v in d.get("extra", {}).get("encrypted_inference", {}).items():
    print(k, v if isinstance(v, str) else (round(v["s_per_image"], 5), [round(x, 5) for x in v["s_per_image_all"]], v["argmax"]))
PY
RS_LAYER_TIMES=1 timeout 300 $RUN scripts/dist_net_check.py cifar/binarynet 2>&1 | grep -v "^W\|^\[W\|NCCL\|\*\*\*\|OMP_NUM" | tail -4 | tee gpurun_out/r2_dist_cifar_n$N.log
