#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest: GEMM / backward subset ==="; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mn_major or wgrad_video or tcgen05_linear or backward_matches_reference or fused_training or variants_match" > gpurun_out/pytest_gemm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gemm.log | cut -c1-300
echo "=== train step ==="; timeout 600 python scripts/gpu_train_step.py > gpurun_out/train_step.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
r=json.load(open('gpurun_out/train_step.json'))
for B in r:
    print(B, 'module', round(r[B]['module_by_module'],3), 'fused', round(r[B]['fused_training'],3))
    print('  module:', r[B]['module_by_module_device_ms_by_entry_point'])
    print('  fused :', r[B]['fused_training_device_ms_by_entry_point'])
PY
