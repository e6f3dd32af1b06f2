#!/bin/bash
# Run under gpurun: one `ncu --set full` capture per named kernel (no launch list). usage: tools/gpu_profile_kernels.sh TAG k1 k2 ...
set -u
TAG=$1; shift
mkdir -p gpurun_out
for k in "$@"; do
  case $k in
    gemv4) re=k_m4_mvm_tma;; gemv8) re=k_m8_mvm;; quantize4|quantize8|quantize4_sr) re=k_vquantize;; dot4) re=k_vdot_fast;;
    mquantize4) re=k_mquantize;; gemm4) re=k_gemm4_tc;; transpose4|transpose8) re=k_mtranspose;;
    threshold4_cluster|threshold8_cluster) re=k_thr_cluster;; threshold4_large) re=k_thr4_apply;; mvmf32|mvm8f32) re=k_mvm_f32_ring;;
    mvm4v8) re=k_m4v8_mvm_tma;; axpy4|axpy8) re=k_vscale_add;;
  esac
  GEMM_N=${GEMM_N:-16384} ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 1 -f -o gpurun_out/prof_${k}_${TAG} \
      python tools/profile_driver.py $k 5 > gpurun_out/prof_${k}_${TAG}.log 2>&1
  tail -1 gpurun_out/prof_${k}_${TAG}.log
done
