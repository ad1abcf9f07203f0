mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest.log | cut -c1-300
timeout 100 python scripts/gpu_variants_diag.py > gpurun_out/variants_diag.log 2>&1; tail -6 gpurun_out/variants_diag.log | cut -c1-220
SEL='layernorm or variants_match or positional or first_and_token'
for TOOL in memcheck racecheck; do
  echo "=== $TOOL ==="
  timeout 240 compute-sanitizer --tool $TOOL --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_variants_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_variants_$TOOL.log | tail -3
done
