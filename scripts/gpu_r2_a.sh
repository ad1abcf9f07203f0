#!/bin/bash
# Round 2, GPU call A (one GPU): smoke, parity tests, PDL A/B, the default bench line, the reference arm, diagnostics, launch list.
# Every step is bounded by `timeout`; a failing smoke or parity run stops the call (a broken kernel must not burn GPU minutes).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== smoke ===";  timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -4 gpurun_out/smoke.log | cut -c1-200
[ $rc -ne 0 ] && exit 1
echo "=== pytest ==="; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/pytest.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== pytest without PDL (MERV_PDL=0), fused-path tests ==="; MERV_PDL=0 timeout 300 python -m pytest tests -m gpu -x -q -k "fused or full_size or benchmark_batch or graph or embedding_buffer" > gpurun_out/pytest_nopdl.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_nopdl.log | cut -c1-300
echo "=== bench (default line) ===";  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-600; tail -5 gpurun_out/bench_n1.err
echo "=== bench, PDL off ===";  MERV_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager > gpurun_out/bench_n1_nopdl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_n1_nopdl.log | cut -c1-300
echo "=== bench, 8 videos (the per-rank shard of config 3 at 8 GPUs), PDL on / off ==="
timeout 300 python bench.py --batch 8 --steps 300 --warmup 10 --input-sets 8 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/bench_b8.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_b8.log | cut -c1-300
MERV_PDL=0 timeout 300 python bench.py --batch 8 --steps 300 --warmup 10 --input-sets 8 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/bench_b8_nopdl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_b8_nopdl.log | cut -c1-300
echo "=== reference arm ==="; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
echo "=== diag ===";   timeout 300 python scripts/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?"; grep -vE "^gemm \[" gpurun_out/diag.log | tail -20
echo "=== launch list ==="
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-torch-eager --no-sustained > gpurun_out/r2_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r2_launches.log | cut -c1-300
