#!/bin/bash
# usage (under gpurun --gpus N): tools/gpu_scale.sh N TAG  - multi-GPU parity check + fused / NCCL bench lines
set -u
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | grep -v 'NCCL\|Warning\|warn' | tee gpurun_out/multi_gpu_check_${TAG}_n${N}.log | tail -20
for ex in fused allreduce; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 5 --exchange $ex 2>&1 | grep "^{" > gpurun_out/bench_${TAG}_n${N}_${ex}.json
done
python tools/show_bench.py gpurun_out/bench_${TAG}_n${N}_fused.json gpurun_out/bench_${TAG}_n${N}_allreduce.json
