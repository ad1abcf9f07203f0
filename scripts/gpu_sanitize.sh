#!/bin/bash
# compute-sanitizer over the small parity cases (memcheck + racecheck + synccheck). Slow: keep the selection small.
mkdir -p gpurun_out
SEL='tiny_linear or ragged_windows or (test_tcgen05_linear_matches_oracle and 100-136) or (cta_pair and mid_linear) or graph_capturable or prefix_written'
for TOOL in memcheck racecheck synccheck; do
  echo "=== $TOOL ==="
  timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 99 python -m pytest tests -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$TOOL.log | tail -4
done
