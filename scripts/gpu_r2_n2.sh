set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin.py -m gpu -q 2>&1 | tail -3 | tee gpurun_out/r2_gpu_tests_n2.log
bash scripts/gpu_r2_multi.sh 2
