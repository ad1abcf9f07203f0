#!/bin/bash
# Round 2 (2 GPUs): multi-GPU parity (fused gather, FSDP) and the N = 2 bench line on the final tree.
mkdir -p gpurun_out
echo "=== pytest multi-GPU ==="; timeout 1500 python -m pytest tests/test_multi_gpu.py tests/test_fsdp.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_2gpu.log | cut -c1-300
echo "=== bench N=2 ==="; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-300
