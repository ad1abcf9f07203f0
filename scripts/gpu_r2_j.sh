#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'gemm|nvjet|cutlass|sm100|xmma|Kernel' -c 12 -f -o gpurun_out/r2b_gemm_vs_cublas python scripts/gpu_gemm_vs_cublas_ncu.py > gpurun_out/r2b_gemm_vs_cublas.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2b_gemm_vs_cublas.log
