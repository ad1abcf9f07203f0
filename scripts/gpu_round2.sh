#!/bin/bash
# One gpurun call for the variant kernels: smoke, the new parity tests first (all failures reported), then the whole GPU suite,
# then the default bench.  Every step is bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
echo "=== smoke ===";  timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc"; tail -3 gpurun_out/smoke.log | cut -c1-200
[ $rc -ne 0 ] && exit 1
echo "=== pytest: variants ==="; timeout 240 python -m pytest tests -m gpu -q -k "variants_match or layernorm or first_and_token or positional" > gpurun_out/pytest_variants.log 2>&1; echo "variants rc=$?"; tail -25 gpurun_out/pytest_variants.log | cut -c1-300
echo "=== pytest: all ==="; timeout 420 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_variants_match_reference > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log | cut -c1-300
echo "=== bench ===";  timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-400
