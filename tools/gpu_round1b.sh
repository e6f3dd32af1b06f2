#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 300 --warmup 5 2>&1 | grep "^{" > gpurun_out/bench_r01b_n2_fused.json
CLOVER_GEMV_IMPL=ring64 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 5 2>&1 | grep "^{" > gpurun_out/bench_r01b_n2_fused_ring64.json
python tools/show_bench.py gpurun_out/bench_r01b_n2_fused.json gpurun_out/bench_r01b_n2_fused_ring64.json
