"""GEMM lab (one GPU): per-iteration durations from a cold start (power / clock behaviour), our variants against cuBLAS on the SAME shapes.

    python scripts/gpu_gemm_lab.py            # writes gpurun_out/gemm_lab.json
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch

from merv_b200 import ops

dev = "cuda:0"
torch.cuda.set_device(0)
rep = {}


def series(fn, n=24, idle=1.5):
    """durations (ms) of n back-to-back calls after `idle` seconds of nothing: shows the clock / power trajectory"""
    fn(); fn()
    torch.cuda.synchronize()
    time.sleep(idle)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]


def summarize(name, d, flops):
    tf = [flops / x / 1e9 for x in d]
    rep[name] = {"ms_first4": [round(x, 4) for x in d[:4]], "ms_last4": [round(x, 4) for x in d[-4:]], "ms_mean": sum(d) / len(d),
                 "TF_first": tf[0], "TF_mean": flops / (sum(d) / len(d)) / 1e9, "TF_last": tf[-1]}
    print(f"{name:44s} first {d[0]:.4f} ms ({tf[0]:.0f} TF)  mean {sum(d)/len(d):.4f} ms ({rep[name]['TF_mean']:.0f} TF)  last {d[-1]:.4f} ms ({tf[-1]:.0f} TF)", flush=True)


g = torch.Generator(device=dev).manual_seed(1)
M, N = 65536, 4096
Ks = [1024, 1024, 768, 768]
As = [torch.randn(M, k, generator=g, device=dev).to(torch.bfloat16) for k in Ks]
Ws = [(torch.randn(N, k, generator=g, device=dev) / 32).to(torch.bfloat16) for k in Ks]
scale = torch.rand(M // 1024, 4, generator=g, device=dev)
bias_mix = torch.randn(M // 1024, N, generator=g, device=dev)
out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
flops_fused = 2 * M * N * sum(Ks)
Acat = torch.cat(As, 1).contiguous()
Wcat = torch.cat(Ws, 1).contiguous()
bias = torch.randn(N, device=dev).to(torch.bfloat16)

for grp in ("1", "2"):
    os.environ["MERV_GEMM_CTA_GROUP"] = grp
    summarize(f"fused 4-seg K=3584 cta_group {grp}", series(lambda: ops.fused_linear_mix(As, Ws, scale, bias_mix, 1024, out=out)), flops_fused)
    summarize(f"plain K=3584 (one segment) cta_group {grp}", series(lambda: ops.linear_bias_act(Acat, Wcat, bias, 0)), flops_fused)
    summarize(f"plain K=1024 cta_group {grp}", series(lambda: ops.linear_bias_act(As[0], Ws[0], bias, 0)), 2 * M * N * 1024)
    summarize(f"gelu K=1024 cta_group {grp}", series(lambda: ops.linear_bias_act(As[0], Ws[0], bias, 1)), 2 * M * N * 1024)
os.environ.pop("MERV_GEMM_CTA_GROUP", None)
summarize("cuBLAS F.linear K=3584", series(lambda: torch.nn.functional.linear(Acat, Wcat, bias)), flops_fused)
summarize("cuBLAS F.linear K=1024", series(lambda: torch.nn.functional.linear(As[0], Ws[0], bias)), 2 * M * N * 1024)
a8 = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
b8 = torch.randn(8192, 8192, generator=g, device=dev).to(torch.bfloat16)
summarize("cuBLAS matmul 8192^3 (the MEASURED_PEAKS shape)", series(lambda: torch.matmul(a8, b8)), 2 * 8192**3)
# SigLIP single (config 4): M = 262144, K = 768
A4 = torch.randn(262144, 768, generator=g, device=dev).to(torch.bfloat16)
W4 = (torch.randn(N, 768, generator=g, device=dev) / 28).to(torch.bfloat16)
for grp in ("1", "2"):
    os.environ["MERV_GEMM_CTA_GROUP"] = grp
    summarize(f"plain M=262144 K=768 cta_group {grp}", series(lambda: ops.linear_bias_act(A4, W4, bias, 0), n=12), 2 * 262144 * N * 768)
os.environ.pop("MERV_GEMM_CTA_GROUP", None)
summarize("cuBLAS F.linear M=262144 K=768", series(lambda: torch.nn.functional.linear(A4, W4, bias), n=12), 2 * 262144 * N * 768)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(REPO, "gpurun_out", "gemm_lab.json"), "w"), indent=1)
