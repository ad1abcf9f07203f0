"""Backward-GEMM lab: K-major (pre-transposed copies) vs MN-major (in place) operands at the shapes of the fused training step."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from merv_b200 import ops
dev = "cuda:0"
g0 = torch.Generator(device=dev).manual_seed(1)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rep = {}
for B in (16, 64):
    M = B * 1024
    g = torch.randn(M, 4096, generator=g0, device=dev).to(torch.bfloat16)
    for C in (1024, 768):
        W = (torch.randn(4096, C, generator=g0, device=dev) / 64).to(torch.bfloat16)
        P = torch.randn(M, C, generator=g0, device=dev).to(torch.bfloat16)
        Wt, gT, Pt = W.t().contiguous(), g.t().contiguous(), P.t().contiguous()
        fl = 2 * M * 4096 * C
        for grp in ("1", "2"):
            os.environ["MERV_GEMM_CTA_GROUP"] = grp
            r = {
                "Z_kmajor": timeit(lambda: ops.linear_bias_act(g, Wt, None, 0)),
                "Z_w_t": timeit(lambda: ops.gemm_ex(g, W, w_t=True)),
                "dW_kmajor": timeit(lambda: ops.linear_bias_act(gT, Pt, None, 0)),
                "dW_a_t_w_t": timeit(lambda: ops.gemm_ex(g, P, a_t=True, w_t=True)),
                "dW_a_t_only": timeit(lambda: ops.gemm_ex(g, Pt, a_t=True)),
                "dW_w_t_only": timeit(lambda: ops.gemm_ex(gT, P, w_t=True)),
            }
            r["cublas_Z"] = timeit(lambda: torch.matmul(g, W))
            r["cublas_dW"] = timeit(lambda: torch.matmul(gT, P))
            rep[f"B{B}_C{C}_grp{grp}"] = {k: (round(v, 4), round(fl / v / 1e9)) for k, v in r.items()}
            print(f"B={B} C={C} cta_group={grp}: " + "  ".join(f"{k} {v:.3f}ms/{fl / v / 1e9:.0f}TF" for k, v in r.items()), flush=True)
os.environ.pop("MERV_GEMM_CTA_GROUP", None)
json.dump(rep, open(os.path.join(REPO, "gpurun_out", "bwd_lab.json"), "w"), indent=1)
