set -x
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-250 | tail -40 | tee gpurun_out/r2c_gpu_tests.log
