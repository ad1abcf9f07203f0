#!/bin/bash
# One gpurun call: diagnostics, GPU parity tests, smoke and a short bench. Every step is bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== diag ===";   timeout 300 python scripts/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?"; tail -40 gpurun_out/diag.log
echo "=== pytest ==="; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest.log
echo "=== smoke ===";  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "=== bench ===";  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
