set -x
mkdir -p gpurun_out
N=8
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
RS_LAYER_TIMES=1 timeout 300 $RUN scripts/dist_net_check.py cifar/binarynet 2>&1 | grep -v "^W\|^\[W\|NCCL\|\*\*\*\|OMP_NUM" | tail -4 | tee gpurun_out/r2_dist_cifar_n$N.log
