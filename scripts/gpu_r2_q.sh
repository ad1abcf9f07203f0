#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest wgrad ==="; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wgrad_video or backward_matches_reference or fused_training" > gpurun_out/pytest_wgrad.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_wgrad.log | cut -c1-300
echo "=== lab ==="; timeout 600 python scripts/gpu_mn_pair_lab.py 2>&1 | tail -32
echo "=== train step ==="; timeout 600 python scripts/gpu_train_step.py > gpurun_out/train_step.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
r=json.load(open('gpurun_out/train_step.json'))
for B in r:
    print(B, 'module', round(r[B]['module_by_module'],3), 'fused', round(r[B]['fused_training'],3), r[B]['fused_training_device_ms_by_entry_point'].get('merv_wgrad_video'))
PY
