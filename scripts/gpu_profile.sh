#!/bin/bash
# ncu evidence for the three hot kernels (run under gpurun, ONE GPU). Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${BENCH_ARGS}"
TAG=${TAG:-r1}
echo "=== launch list ==="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_launches.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/${TAG}_launches.log | cut -c1-300
for K in gemm_bf16_tcgen05 pool3d_tma ${EXTRA_KERNELS}; do
  echo "=== full capture: $K ==="
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_$K $BENCH > gpurun_out/${TAG}_prof_$K.log 2>&1
  echo "rc=$?"; tail -2 gpurun_out/${TAG}_prof_$K.log | cut -c1-300
done
ls -la gpurun_out/
