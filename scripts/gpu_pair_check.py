"""cta_group::2 GEMM vs the single-CTA kernel: bit-level agreement on several shapes, then timing."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from merv_b200 import ops

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)


def run(ctas, fn):
    os.environ["MERV_GEMM_CTA_GROUP"] = str(ctas)
    out = fn()
    torch.cuda.synchronize()
    return out


ok = True
for (M, N, K) in [(256, 256, 64), (256, 512, 256), (128, 256, 128), (1000, 264, 200), (4096, 4096, 1024)]:
    a = torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device=dev) / K**0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device=dev).to(torch.bfloat16)
    y1, _ = run(1, lambda: ops.linear_bias_act(a, w, b, 0))
    y2, _ = run(2, lambda: ops.linear_bias_act(a, w, b, 0))
    ref = a.float() @ w.float().T + b.float()
    e1 = ((y1.float() - ref).abs().max() / ref.abs().max()).item()
    e2 = ((y2.float() - ref).abs().max() / ref.abs().max()).item()
    same = torch.equal(y1, y2)
    ok &= e2 < 6e-3
    print(f"{M}x{N}x{K}: single err {e1:.2e}  pair err {e2:.2e}  bit-identical {same}", flush=True)
if not ok:
    print("PAIR KERNEL WRONG", flush=True)
    sys.exit(1)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


B = 64
a = torch.randn(B * 1024, 1024, generator=g, device=dev).to(torch.bfloat16)
w = (torch.randn(4096, 1024, generator=g, device=dev) / 32).to(torch.bfloat16)
bias = torch.randn(4096, device=dev).to(torch.bfloat16)
As = [torch.randn(B * 1024, c, generator=g, device=dev).to(torch.bfloat16) for c in (1024, 1024, 768, 768)]
Ws = [(torch.randn(4096, c, generator=g, device=dev) / 32).to(torch.bfloat16) for c in (1024, 1024, 768, 768)]
scale = torch.softmax(torch.randn(B, 4, device=dev), -1).contiguous()
bm = torch.randn(B, 4096, device=dev)
for ctas in (1, 2):
    os.environ["MERV_GEMM_CTA_GROUP"] = str(ctas)
    t = timeit(lambda: ops.linear_bias_act(a, w, bias, 0))
    t1 = timeit(lambda: ops.linear_bias_act(a, w, bias, 1))
    tf = timeit(lambda: ops.fused_linear_mix(As, Ws, scale, bm, 1024))
    print(f"cta_group {ctas}: plain {t:.3f} ms = {2*B*1024*4096*1024/t/1e9:.0f} TF/s | gelu {t1:.3f} ms | fused 4-seg {tf:.3f} ms = {2*B*1024*4096*3584/tf/1e9:.0f} TF/s", flush=True)
o1 = run(1, lambda: ops.fused_linear_mix(As, Ws, scale, bm, 1024))
o2 = run(2, lambda: ops.fused_linear_mix(As, Ws, scale, bm, 1024))
print("fused outputs bit-identical:", torch.equal(o1, o2), flush=True)
