#!/bin/bash
# Round 2: per-video weight gradient (merv_wgrad_video) — parity subset + training-step timing.
mkdir -p gpurun_out
echo "=== pytest subset ==="; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wgrad_video" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_bwd.log | cut -c1-300
echo "=== train step ==="; timeout 600 python scripts/gpu_train_step.py > gpurun_out/train_step.log 2>&1; echo "rc=$?"; tail -45 gpurun_out/train_step.log | cut -c1-200
