#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_multi_gpu.py -x -q -k "mvm4 or matrix_quantize_and_mvm or matrix_stochastic or fused" > gpurun_out/t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t1.log
tail -4 gpurun_out/t1.log
for rows in 65536 32768 16384 8192; do
for impl in tma ring64; do CLOVER_GEMV_IMPL=$impl timeout 120 python tools/gemv_bench.py 4 65536 100 $rows 2>&1 | tail -1; done
done
