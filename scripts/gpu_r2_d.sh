#!/bin/bash
# Round 2, pool-assist bring-up: lab (bit-identity + head sweep + per-warp cycle counters).
mkdir -p gpurun_out
echo "=== assist lab ==="; timeout 480 python scripts/gpu_assist_lab.py --batches 64 --heads 24,32,40 --dbg --out gpurun_out/assist_lab.json > gpurun_out/assist_lab.log 2>&1; echo "lab rc=$?"; tail -25 gpurun_out/assist_lab.log | cut -c1-420
