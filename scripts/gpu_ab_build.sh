#!/bin/bash
# A/B of two builds of the library on ONE box, alternating processes (the GPU drifts with temperature / power by more than most effects).
#   here:   MERV_BUILD_TAG=<tag> MERV_BUILD_DEFINES="-DX=0" python merv_b200/build.py     (writes merv_b200/libmerv_fusion_<tag>.so)
#   then:   gpurun -- bash scripts/gpu_ab_build.sh <tag>
# Runs scripts/gpu_gemm_lab.py (every GEMM shape next to cuBLAS) and a short bench.py with each library twice; results in gpurun_out/.
# Used for: the warp-uniform issue loops (-DMERV_GEMM_WARP_UNIFORM=0, profiles/r2b_warp_uniform_ab.json) and the parked barrier waits
# (-DMERV_WAIT_HINT_NS=2000000, profiles/r2b_wait_hint_ab.json); MERV_GEMM_WIDE_OUT=0|1 in the environment gave profiles/r2b_gemm_lab_wide.json.
TAG=${1:?usage: gpu_ab_build.sh <tag>}
mkdir -p gpurun_out
for i in 1 2; do
  for which in $TAG production; do
    if [ $which = production ]; then unset MERV_FUSION_LIB; else export MERV_FUSION_LIB=$PWD/merv_b200/libmerv_fusion_$TAG.so; fi
    echo "=== $which run $i ==="
    timeout 600 python scripts/gpu_gemm_lab.py > gpurun_out/gemm_lab_${which}_$i.log 2>&1; echo "rc=$?"
    cp gpurun_out/gemm_lab.json gpurun_out/gemm_lab_${which}_$i.json
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/bench_${which}_$i.log 2>&1
    tail -1 gpurun_out/bench_${which}_$i.log | cut -c1-160
  done
done
unset MERV_FUSION_LIB
