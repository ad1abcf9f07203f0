#!/bin/bash
# Round 2 final validation on one GPU: full suite, smoke, default bench, reference arm, ncu capture of the fused GEMM and the pool kernel, launch list, train step.
mkdir -p gpurun_out
echo "=== pytest ==="; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== smoke ==="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench ==="; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-300
echo "=== bench reference arm ==="; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-200
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained"
for K in gemm_bf16_tcgen05 pool3d_tma; do
  echo "=== ncu full: $K ==="
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -f -o gpurun_out/r2c_prof_$K $BENCH > gpurun_out/r2c_prof_$K.log 2>&1
  echo "rc=$?"
done
echo "=== launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/r2c_launches.log 2>&1
echo "rc=$?"
echo "=== train step ==="; timeout 600 python scripts/gpu_train_step.py > gpurun_out/train_step.log 2>&1; echo "rc=$?"
bash scripts/gpu_sanitize_r2b.sh
