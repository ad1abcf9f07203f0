#!/bin/bash
# One gpurun call: diagnostics, GPU parity tests, smoke and a short bench. Every step is bounded by `timeout`, and the
# round stops at the first step that fails or hangs (a broken kernel must not burn GPU minutes in later steps).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== smoke ===";  timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -4 gpurun_out/smoke.log | cut -c1-200
[ $rc -ne 0 ] && exit 1
echo "=== pytest ==="; timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -12 gpurun_out/pytest.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== pytest, wide output boxes (the fused all-gather variant of the GEMM) ==="; MERV_GEMM_WIDE_OUT=1 timeout 240 python -m pytest tests -m gpu -x -q -k "tcgen05_linear or modules_bf16 or cta_pair or fused or embedding_buffer or strided_output or full_size" > gpurun_out/pytest_wide.log 2>&1; rc=$?; echo "pytest-wide rc=$rc"; tail -5 gpurun_out/pytest_wide.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== diag ===";   timeout 180 python scripts/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?"; grep -vE "^gemm \[" gpurun_out/diag.log | tail -20
echo "=== bench ===";  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
