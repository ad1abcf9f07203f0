#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
  for w in 0 1; do
    echo "=== wide_out=$w run $i ==="
    MERV_GEMM_WIDE_OUT=$w timeout 600 python scripts/gpu_gemm_lab.py > gpurun_out/gemm_lab_wide${w}_$i.log 2>&1; echo "rc=$?"
    cp gpurun_out/gemm_lab.json gpurun_out/gemm_lab_wide${w}_$i.json
  done
done
