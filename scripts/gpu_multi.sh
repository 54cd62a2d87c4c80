# usage: bash scripts/gpu_multi.sh N   -- N-GPU checks: sharded nets reproduce single-GPU ciphertexts; bench line at N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/dist_net_check.py mnist/sign1024x1 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/dist_mnist_n$N.log
RS_CHECK_SINGLE=1 timeout 900 $TR scripts/dist_net_check.py cifar/binarynet_small 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/dist_cifar_small_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
