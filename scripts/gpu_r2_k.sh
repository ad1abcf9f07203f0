#!/bin/bash
# A/B: barrier waits parked in hardware (try_wait + suspend-time hint) vs the un-hinted spin, alternating processes on the same box.
mkdir -p gpurun_out
for i in 1 2; do
  for tag in nohint hint; do
    if [ $tag = nohint ]; then export MERV_FUSION_LIB=$PWD/merv_b200/libmerv_fusion_nohint.so; else unset MERV_FUSION_LIB; fi
    echo "=== $tag run $i ==="
    timeout 600 python scripts/gpu_gemm_lab.py > gpurun_out/gemm_lab_${tag}_$i.log 2>&1; echo "rc=$?"
    cp gpurun_out/gemm_lab.json gpurun_out/gemm_lab_${tag}_$i.json
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/bench_${tag}_$i.log 2>&1
    tail -1 gpurun_out/bench_${tag}_$i.log | cut -c1-160
  done
done
unset MERV_FUSION_LIB
