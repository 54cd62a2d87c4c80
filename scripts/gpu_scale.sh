# usage: bash scripts/gpu_scale.sh N  -- bench line at N GPUs (includes CIFAR / MNIST seconds per image, sharded)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -3 gpurun_out/bench_n$N.err
