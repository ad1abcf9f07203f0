#!/bin/bash
# Round 2 (8 GPUs): the N = 8 bench line on the final tree (weak + strong + gather + e2e).
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.log 2>gpurun_out/bench_n8.err; echo "rc=$?"; tail -1 gpurun_out/bench_n8.log | cut -c1-300
