#!/bin/bash
# one GPU call: new-kernel parity + micro-benchmarks (round 1, session 3)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "mvm8_pipelined or c5 or matrix_quantize_and_mvm or matrix_stochastic or experimental_pipelines" > gpurun_out/t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t1.log
tail -5 gpurun_out/t1.log
for impl in tma simple; do
  CLOVER_GEMV_IMPL=$impl timeout 120 python tools/gemv_bench.py 8 32768 100 2>&1 | tail -1
done
timeout 120 python tools/gemv_bench.py 8 16384 200 2>&1 | tail -1
timeout 200 python tools/gemm_bench.py 16384 10 2>&1 | tail -1
CLOVER_GEMM_KERNEL=p192 timeout 200 python tools/gemm_bench.py 16384 10 2>&1 | tail -1
CLOVER_GEMM_KERNEL=p192 timeout 200 python tools/gemm_bench.py 8192 10 2>&1 | tail -1
