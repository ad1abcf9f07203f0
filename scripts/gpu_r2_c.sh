#!/bin/bash
# Round 2, GPU call (one GPU): parity + default bench + ncu full captures of the hot kernels (pair GEMM, pool).
mkdir -p gpurun_out
echo "=== pytest ==="; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -5 gpurun_out/pytest.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== bench ===";  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-300
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained"
for K in gemm_bf16_tcgen05 pool3d_tma; do
  echo "=== ncu full: $K ==="
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/r2_prof_$K $BENCH > gpurun_out/r2_prof_$K.log 2>&1
  echo "rc=$?"; tail -1 gpurun_out/r2_prof_$K.log | cut -c1-200
done
echo "=== ncu full: plain pair GEMM K=1024 (gpu_diag shapes) ==="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 -s 40 -c 2 -f -o gpurun_out/r2_prof_gemm_plain python scripts/gpu_diag.py > gpurun_out/r2_prof_gemm_plain.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
