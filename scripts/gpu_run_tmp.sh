mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "backward or fused_training or helper_kernels" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_bwd.log | cut -c1-300
timeout 300 python scripts/gpu_train_step.py 16 64 > gpurun_out/train_step.log 2>&1; echo "train rc=$?"; tail -4 gpurun_out/train_step.log | cut -c1-1500
