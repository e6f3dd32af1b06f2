#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "transpose" > gpurun_out/t1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t1.log
tail -12 gpurun_out/t1.log
for b in 4 8; do timeout 100 python tools/transpose_bench.py $b 16384 50 2>&1 | tail -1; done
timeout 100 python tools/transpose_bench.py 4 32768 20 2>&1 | tail -1
