#!/bin/bash
# Round 2, validation of the tree after pool assist / 3dconv / per-video weight gradient: full suite, sanitizer, default bench, launch list.
mkdir -p gpurun_out
echo "=== pytest ==="; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 gpurun_out/pytest.log | cut -c1-300
[ $rc -ne 0 ] && exit 1
echo "=== smoke ==="; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench ==="; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_n1.log | cut -c1-400
echo "=== bench reference arm ==="; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-300
bash scripts/gpu_sanitize_r2b.sh
