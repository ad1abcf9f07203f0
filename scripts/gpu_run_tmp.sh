mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log | cut -c1-300
SEL='layernorm or variants_match or positional or first_and_token or helper_kernels or (backward_matches and tiny_linear) or square_grid'
for TOOL in memcheck racecheck; do
  MERV_GEMM_CTA_GROUP=1 timeout 300 compute-sanitizer --tool $TOOL --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_r1g_$TOOL.log 2>&1
  echo "$TOOL rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_r1g_$TOOL.log | tail -3
done
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-300
