#!/bin/bash
# Run under gpurun: ncu launch list of the bench command + one --set full capture per hot kernel.
# Outputs land in gpurun_out/ (scratch); summaries are written into profiles/ by tools/summarize_ncu.py here.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for k in ${KERNELS:-gemv4 quantize4 dot4 gemv8 gemm4 mquantize4}; do
  case $k in
    gemv4) re=k_m4_mvm_tma;; gemv8) re=k_m8_mvm;; quantize4|quantize8|quantize4_sr) re=k_vquantize;; dot4) re=k_vdot_fast;;
    mquantize4) re=k_mquantize;; gemm4) re=k_gemm4_tc;; transpose4|transpose8) re=k_mtranspose;;
    threshold4_cluster|threshold8_cluster) re=k_thr_cluster;; threshold4_large) re=k_thr_hist;; mvmf32|mvm8f32) re=k_mvm_f32_ring;; mvm4v8) re=k_m4v8_mvm_tma;; axpy4) re=k_vscale_add;;
  esac
  GEMM_N=${GEMM_N:-16384} ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 1 -f -o gpurun_out/prof_${k}_${TAG} \
      python tools/profile_driver.py $k 5 > gpurun_out/prof_${k}_${TAG}.log 2>&1
done
ls -la gpurun_out
