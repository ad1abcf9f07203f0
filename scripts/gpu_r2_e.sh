#!/bin/bash
# Round 2: full GPU suite after the pool-assist / 3dconv additions.
mkdir -p gpurun_out
echo "=== pytest ==="; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -15 gpurun_out/pytest.log | cut -c1-400
