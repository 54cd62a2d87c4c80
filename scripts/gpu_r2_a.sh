# Round 2, first GPU pass: parity tests (incl. the sampled full-size tests), bench line with the native net runner.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
nproc
timeout 2400 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 | tee gpurun_out/r2a_gpu_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 2> gpurun_out/r2a_bench.err | tee gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
timeout 300 python scripts/pbs_sizes.py 128 148 296 592 1024 1536 3072 4096 2>&1 | tee gpurun_out/r2a_pbs_sizes.log
