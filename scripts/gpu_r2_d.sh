set -x
mkdir -p gpurun_out
bash scripts/shim_stress.sh 2>&1 | tail -8 | tee gpurun_out/r2d_dropin_stress.log
timeout 3000 python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | cut -c1-250 | tail -25 | tee gpurun_out/r2d_gpu_tests.log
