set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | cut -c1-300 | tail -60 | tee gpurun_out/r2b_gpu_tests.log
