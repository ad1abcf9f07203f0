"""GPU diagnostics for a fresh box: localises tcgen05/TMA descriptor mistakes and prints quick kernel timings.

    python scripts/gpu_diag.py            # writes gpurun_out/diag.json and prints a summary
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np
import torch

from merv_b200 import ops

dev = "cuda:0"
report = {"device": torch.cuda.get_device_name(0), "cap": torch.cuda.get_device_capability(0)}


def gemm_check(M, N, K, pattern):
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K)
    if pattern == "rand":
        a = torch.randn(M, K, generator=g, device=dev)
        w = torch.randn(N, K, generator=g, device=dev) / K**0.5
    elif pattern == "eye":  # y = a[:, :N] when K >= N: exposes row/column permutations and swizzle errors
        a = torch.randn(M, K, generator=g, device=dev)
        w = torch.eye(N, K, device=dev)
    elif pattern == "kblock":  # only one 16-wide K slice is non-zero: exposes wrong K advance inside the swizzle atom
        a = torch.zeros(M, K, device=dev)
        a[:, 16:32] = torch.randn(M, 16, generator=g, device=dev)
        w = torch.randn(N, K, generator=g, device=dev)
    a, w = a.to(torch.bfloat16), w.to(torch.bfloat16)
    y, _ = ops.linear_bias_act(a, w, None, 0)
    torch.cuda.synchronize()
    want = a.float() @ w.float().T
    err = (y.float() - want).abs()
    scale = want.abs().max().item() + 1e-9
    res = {"shape": [M, N, K], "pattern": pattern, "rel_err": err.max().item() / scale}
    if res["rel_err"] > 1e-2:
        bad = (err > 1e-2 * scale)
        res["bad_frac"] = bad.float().mean().item()
        res["bad_rows_by_32"] = [bad[r:r + 32].float().mean().item() for r in range(0, min(M, 128), 32)]
        res["bad_cols_by_32"] = [bad[:, c:c + 32].float().mean().item() for c in range(0, min(N, 256), 32)]
        res["y00"] = y[0, :8].float().tolist()
        res["want00"] = want[0, :8].tolist()
    return res


report["gemm"] = []
try:
    for shape in [(128, 256, 64), (128, 256, 128), (128, 256, 256), (256, 512, 512), (100, 136, 72)]:
        for pattern in ("rand", "eye", "kblock"):
            r = gemm_check(*shape, pattern)
            report["gemm"].append(r)
            print("gemm", r["shape"], r["pattern"], f"rel_err={r['rel_err']:.3e}", "BAD" if r["rel_err"] > 1e-2 else "ok", flush=True)
except Exception as e:  # keep going: the pool / mix kernels are independent
    report["gemm_exception"] = repr(e)
    print("gemm exception:", e, flush=True)


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


try:
    B = 64
    g = torch.Generator(device=dev).manual_seed(1)
    shapes = [(B, 16, 256, 1024), (B, 16, 256, 1024), (B, 16, 196, 768), (B, 16, 196, 768)]
    xs = [torch.randn(s, generator=g, device=dev).to(torch.bfloat16) for s in shapes]
    in_bytes = sum(x.numel() * 2 for x in xs)
    out_bytes = sum(B * 1024 * s[3] * 2 for s in shapes)
    svs = [torch.randn(s[3], device=dev) for s in shapes]
    t = timeit(lambda: ops.pool3d(xs, [16] * 4, 8, score_vecs=svs))
    report["pool_ms"] = t
    report["pool_GBps"] = (in_bytes + out_bytes) / t / 1e6
    t2 = timeit(lambda: ops.pool3d(xs, [16] * 4, 8))
    report["pool_noscore_ms"] = t2
    print(f"pool3d B=64: {t:.3f} ms = {report['pool_GBps']:.0f} GB/s (without score partials {t2:.3f} ms)", flush=True)
    for i in (0, 2):
        ti = timeit(lambda: ops.pool3d([xs[i]], [16], 8))
        b = xs[i].numel() * 2 + B * 1024 * shapes[i][3] * 2
        print(f"  pool3d encoder {i} alone: {ti:.3f} ms = {b / ti / 1e6:.0f} GB/s", flush=True)
        report[f"pool_enc{i}_GBps"] = b / ti / 1e6
    os.environ["MERV_POOL_IMPL"] = "direct"
    td = timeit(lambda: ops.pool3d(xs, [16] * 4, 8, score_vecs=svs))
    del os.environ["MERV_POOL_IMPL"]
    report["pool_direct_ms"] = td
    print(f"  direct (non-TMA) kernel: {td:.3f} ms = {(in_bytes + out_bytes) / td / 1e6:.0f} GB/s", flush=True)
    del xs
    Ys = [torch.randn(B, 1024, 4096, generator=g, device=dev).to(torch.bfloat16) for _ in range(4)]
    sc = torch.randn(B, 4, device=dev)
    t = timeit(lambda: ops.softmax_mix(Ys, 1024, scores=sc))
    report["mix_ms"] = t
    report["mix_GBps"] = 5 * B * 1024 * 4096 * 2 / t / 1e6
    print(f"softmax_mix B=64: {t:.3f} ms = {report['mix_GBps']:.0f} GB/s", flush=True)
    u = torch.randn(4096, device=dev)
    t = timeit(lambda: ops.scores_from_tokens(Ys, u, 1024))
    report["token_scores_ms"] = t
    print(f"scores_from_tokens B=64: {t:.3f} ms = {4 * B * 1024 * 4096 * 2 / t / 1e6:.0f} GB/s", flush=True)
    del Ys
    a = torch.randn(B * 1024, 1024, generator=g, device=dev).to(torch.bfloat16)
    w = (torch.randn(4096, 1024, generator=g, device=dev) / 32).to(torch.bfloat16)
    bias = torch.randn(4096, device=dev).to(torch.bfloat16)
    for act in (0, 1):
        t = timeit(lambda: ops.linear_bias_act(a, w, bias, act), n=5)
        tf = 2 * B * 1024 * 4096 * 1024 / t / 1e9
        report[f"gemm_65536x4096x1024_act{act}_ms"] = t
        report[f"gemm_65536x4096x1024_act{act}_TFLOPs"] = tf
        print(f"tcgen05 GEMM 65536x4096x1024 act={act}: {t:.3f} ms = {tf:.0f} TFLOP/s", flush=True)
    rv = torch.randn(4096, device=dev)
    t = timeit(lambda: ops.linear_bias_act(a, w, bias, 1, rowdot_vec=rv), n=5)
    report["gemm_65536x4096x1024_gelu_rowdot_ms"] = t
    print(f"tcgen05 GEMM 65536x4096x1024 act=1 + row-dot: {t:.3f} ms = {2 * B * 1024 * 4096 * 1024 / t / 1e9:.0f} TFLOP/s", flush=True)
    for grp in ("1", "2"):  # which variant is faster for the activated GEMM (the default policy picks 1)
        os.environ["MERV_GEMM_CTA_GROUP"] = grp
        t = timeit(lambda: ops.linear_bias_act(a, w, bias, 1), n=5)
        report[f"gemm_65536x4096x1024_gelu_cta_group{grp}_ms"] = t
        print(f"  act=1 with cta_group {grp}: {t:.3f} ms = {2 * B * 1024 * 4096 * 1024 / t / 1e9:.0f} TFLOP/s", flush=True)
    os.environ.pop("MERV_GEMM_CTA_GROUP", None)
    ref_t = timeit(lambda: torch.nn.functional.linear(a, w, bias), n=5)
    report["cublas_65536x4096x1024_TFLOPs"] = 2 * B * 1024 * 4096 * 1024 / ref_t / 1e9
    print(f"cuBLAS (torch F.linear) same shape: {ref_t:.3f} ms = {report['cublas_65536x4096x1024_TFLOPs']:.0f} TFLOP/s", flush=True)
except Exception as e:
    report["timing_exception"] = repr(e)
    print("timing exception:", e, flush=True)

try:
    import merv_b200 as M

    mod = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024).to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
    g = torch.Generator(device=dev).manual_seed(3)
    for Bs in (1, 2, 8):
        feats = [torch.randn((Bs, 16, n, c), generator=g, device=dev).to(torch.bfloat16) for n, c in zip([256, 256, 196, 196], [1024, 1024, 768, 768])]
        with torch.inference_mode():
            dev_ms = timeit(lambda: mod(feats), n=50)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(50):
                mod(feats)
            host_issue_ms = (time.perf_counter() - t0) / 50 * 1e3
            torch.cuda.synchronize()
        report[f"latency_B{Bs}_ms"] = dev_ms
        report[f"host_issue_B{Bs}_ms"] = host_issue_ms
        print(f"fused path B={Bs}: {dev_ms * 1e3:.1f} us per call on the device timeline, host issue time {host_issue_ms * 1e3:.1f} us", flush=True)
except Exception as e:
    report["latency_exception"] = repr(e)
    print("latency exception:", e, flush=True)

try:  # training-shaped call: forward + backward at the reference's per-device batch (conf/models.py:125: 16)
    import merv_b200 as M
    from oracle import torch_port

    Bt = 16
    mod = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024).to(device=dev, dtype=torch.bfloat16).train()
    g = torch.Generator(device=dev).manual_seed(5)
    feats = [torch.randn((Bt, 16, n, c), generator=g, device=dev).to(torch.bfloat16) for n, c in zip([256, 256, 196, 196], [1024, 1024, 768, 768])]
    G = torch.randn((Bt, 1024, 4096), generator=g, device=dev).to(torch.bfloat16)

    def ours_step():
        out, w = mod(feats)
        out.backward(G)
        mod.zero_grad(set_to_none=True)

    pp = [{k: v.detach().clone().requires_grad_(True) for k, v in p.projector.state_dict().items()} for p in mod.projectors]
    fp = {k: v.detach().clone().requires_grad_(True) for k, v in mod.feature_fusion.state_dict().items()}

    def eager_step():
        out, w = torch_port.fusion_forward_autograd(feats, pp, fp, [16] * 4, 8, "linear", 1024)
        out.backward(G)
        for d in pp + [fp]:
            for t in d.values():
                t.grad = None

    t_ours = timeit(ours_step, n=10)
    t_eager = timeit(eager_step, n=10)
    report["train_step_B16_ms"] = {"ours": t_ours, "torch_eager_reference_ops": t_eager}
    print(f"training-shaped fwd+bwd B=16 (bf16, linear): ours {t_ours:.3f} ms vs reference op sequence in torch eager {t_eager:.3f} ms", flush=True)
except Exception as e:
    report["train_exception"] = repr(e)
    print("train-step exception:", e, flush=True)

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(REPO, "gpurun_out", "diag.json"), "w"), indent=1)
