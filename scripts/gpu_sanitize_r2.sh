#!/bin/bash
# compute-sanitizer over the kernels added / changed in round 2 (memcheck + racecheck + synccheck), small shapes only.
mkdir -p gpurun_out
SEL='(mn_major and 104-136) or (mn_major and 256-512) or (cross_attention_kernel and 2-32-37-70) or (cross_attention_backward and 2-32-37-70) or (attentive_pooler_training and 256-128) or tiny_linear or graph_capturable or (full_size_fused_matches)'
for TOOL in memcheck racecheck synccheck; do
  echo "=== $TOOL ==="
  timeout 1200 compute-sanitizer --tool $TOOL --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_r2_$TOOL.log 2>&1
  echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_r2_$TOOL.log | tail -4
done
