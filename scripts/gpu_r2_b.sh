#!/bin/bash
# Round 2, GPU call B (TWO GPUs): multi-GPU parity (FSDP drop-in under real FSDP, fused GEMM + all-gather both transports), then the
# default bench at N = 2 (weak line + strong / gather / e2e sub-records).
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== multi-GPU pytest ==="; timeout 900 python -m pytest tests/test_fsdp.py tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_2gpu.log 2>&1; rc=$?; echo "rc=$rc"; tail -30 gpurun_out/pytest_2gpu.log | cut -c1-1500
echo "=== bench N=2 ==="; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-400; tail -5 gpurun_out/bench_n2.err | cut -c1-400
