#!/bin/bash
# A/B: warp-uniform producer / MMA-issuer loops (elect.sync) vs lane-0 guarded loops, alternating processes on the same box.
mkdir -p gpurun_out
echo "=== pytest GEMM subset (warp-uniform build) ==="; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mn_major or wgrad_video or tcgen05_linear or full_size or benchmark_batch or pool_assist" > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gemm.log | cut -c1-300
for i in 1 2; do
  for tag in lane0 uniform; do
    if [ $tag = lane0 ]; then export MERV_FUSION_LIB=$PWD/merv_b200/libmerv_fusion_lane0.so; else unset MERV_FUSION_LIB; fi
    echo "=== $tag run $i ==="
    timeout 600 python scripts/gpu_gemm_lab.py > gpurun_out/gemm_lab_${tag}_$i.log 2>&1; echo "rc=$?"
    cp gpurun_out/gemm_lab.json gpurun_out/gemm_lab_${tag}_$i.json
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/bench_${tag}_$i.log 2>&1
    tail -1 gpurun_out/bench_${tag}_$i.log | cut -c1-160
  done
done
unset MERV_FUSION_LIB
