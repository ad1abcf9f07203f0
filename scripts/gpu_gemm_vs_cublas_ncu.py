"""Two launches for an ncu side-by-side: this repo's plain pair GEMM and cuBLAS (F.linear) on the SigLIP-single shape
(M = 262144, K = 768, N = 4096) and on M = 65536, K = 1024.  Run under `ncu --set full -k regex:'gemm|nvjet|cutlass|sm100|xmma'`."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

from merv_b200 import ops

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(1)
N = 4096
bias = torch.randn(N, device=dev).to(torch.bfloat16)
for M, K in ((262144, 768), (65536, 1024)):
    A = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device=dev) / 28).to(torch.bfloat16)
    for _ in range(3):
        ops.linear_bias_act(A, W, bias, 0)
    for _ in range(3):
        torch.nn.functional.linear(A, W, bias)
    torch.cuda.synchronize()
    del A, W
