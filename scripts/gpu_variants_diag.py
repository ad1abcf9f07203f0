"""Throughput of the variant kernels at production shapes (bf16, B200): LayerNorm over row segments (concat_channel_ln),
pre-projection LayerNorm, per-token-u scores (averagetoken=False).  Algorithmic bytes = distinct inputs + outputs.

    python scripts/gpu_variants_diag.py        # prints a summary, writes gpurun_out/variants_diag.json
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch

from merv_b200 import ops

dev = "cuda:0"
PEAK = 6553.6
try:
    PEAK = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", PEAK))
except Exception:
    pass


def timed(fns, iters=24, warm=4):
    """Steady-state ms per call over ROTATING argument sets (each larger than L2 together with its output), one CUDA-event pair
    around the whole loop: no dirty-line flush that the timed kernel would have to write back."""
    for i in range(warm):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    # the loop is replayed from a CUDA graph: a 60 us kernel would otherwise be timed at the rate Python can issue it
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            fns[i % len(fns)]()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


rep = {"device": torch.cuda.get_device_name(0), "hbm_peak_GBps": PEAK}
g = torch.Generator(device=dev).manual_seed(1)

# concat_channel_ln at merv-full width: 16 videos x 1024 tokens, 4 encoders x 4096 channels
M, K, E = 16 * 1024, 4096, 4
SETS = 3
segsets = [[torch.randn(M, K, generator=g, device=dev).to(torch.bfloat16) for _ in range(E)] for _ in range(SETS)]
segs = segsets[0]
gamma = torch.ones(E * K, device=dev, dtype=torch.bfloat16)
beta = torch.zeros(E * K, device=dev, dtype=torch.bfloat16)
ms = timed([lambda ss=ss: ops.layernorm(ss, gamma, beta) for ss in segsets])
b = 2 * M * E * K * 2
rep["layernorm_segments_4x4096"] = {"rows": M, "ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / PEAK}

# pre-projection LayerNorm on pooled LanguageBind tokens: 64 videos x 1024 tokens x 1024 channels (warp per row)
M2, K2 = 64 * 1024, 1024
xs = [torch.randn(M2, K2, generator=g, device=dev).to(torch.bfloat16) for _ in range(SETS)]
x = xs[0]
g2, b2 = torch.ones(K2, device=dev, dtype=torch.bfloat16), torch.zeros(K2, device=dev, dtype=torch.bfloat16)
ms = timed([lambda xx=xx: ops.layernorm([xx], g2, b2) for xx in xs])
b = 2 * M2 * K2 * 2
rep["layernorm_1024"] = {"rows": M2, "ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / PEAK}

# LayerNorm backward (dX + dY*xhat) on the same tensor: reads x, dy; writes dx, gx
dy = torch.randn(M2, K2, generator=g, device=dev).to(torch.bfloat16)
ms = timed([lambda xx=xx: ops.layernorm_backward([xx], dy, g2) for xx in xs])
b = 4 * M2 * K2 * 2
rep["layernorm_backward_1024_incl_colsums"] = {"rows": M2, "ms": ms, "GBps_of_main_kernel_bytes": b / ms / 1e6}

# scores with one u row per token (averagetoken=False), 16 videos, 4 encoders, T = 1024, K = 4096
B, T = 16, 1024
V = [s.view(B, T, K) for s in segs]
u = torch.randn(T * K, generator=g, device=dev)
Vs = [[t.view(B, T, K) for t in ss] for ss in segsets]
ms = timed([lambda vv=vv: ops.scores_from_tokens(vv, u, T, per_token_u=True, mean=False) for vv in Vs])
b = E * B * T * K * 2 + T * K * 4
rep["scores_per_token_u"] = {"ms": ms, "GBps": b / ms / 1e6, "frac": b / ms / 1e6 / PEAK}

os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(REPO, "gpurun_out", "variants_diag.json"), "w"), indent=1)
for k, v in rep.items():
    print(k, v)

# attentive pooler at the LanguageBind shape (16 frames x 256 patches x 1024 channels -> 64 queries per frame, 8 heads), 16 videos
try:
    import merv_b200 as M

    Ba = 16
    pool = M.AttentivePooler(1024, 4096, num_query_tokens=64, num_heads=8, output_frames=16, mlp_type="linear").to(device=dev, dtype=torch.bfloat16)
    pool = pool.eval().requires_grad_(False)
    xa = [torch.randn(Ba, 16, 256, 1024, generator=g, device=dev).to(torch.bfloat16) for _ in range(2)]
    with torch.inference_mode():
        ms_all = timed([lambda xx=xx: pool(xx) for xx in xa], iters=8, warm=2)
        with ops.KernelTimer(timing=True) as kt:
            pool(xa[0])
        per = {k: round(sum(v), 4) for k, v in kt.durations_ms().items()}
    att_flop = 2 * 2 * 64 * 256 * 1024 * 16 * Ba  # QK^T and PV
    kv_bytes = Ba * 16 * 256 * 2048 * 2  # the attention's compulsory read: K and V of every frame once
    rep["attentive_pooler_languagebind_16_videos"] = {
        "ms": ms_all, "device_ms_by_entry_point": dict(sorted(per.items(), key=lambda kv: -kv[1])),
        "attention_TFLOPs_tcgen05": att_flop / per.get("merv_cross_attention", float("nan")) / 1e9,
        "attention_kv_GBps": kv_bytes / per.get("merv_cross_attention", float("nan")) / 1e6,
    }
    os.environ["MERV_ATTN_IMPL"] = "simt"
    with torch.inference_mode():
        with ops.KernelTimer(timing=True) as kt:
            pool(xa[0])
    os.environ.pop("MERV_ATTN_IMPL")
    rep["attentive_pooler_languagebind_16_videos"]["attention_ms_simt_fp32_kernel"] = round(sum(kt.durations_ms()["merv_cross_attention"]), 4)
    # training step of the resampler (forward + backward, bf16)
    poolt = M.AttentivePooler(1024, 4096, num_query_tokens=64, num_heads=8, output_frames=16, mlp_type="linear").to(device=dev, dtype=torch.bfloat16).train()
    Gt = torch.randn(Ba, 16 * 64, 4096, generator=g, device=dev).to(torch.bfloat16)

    def train_step(xx):
        out = poolt(xx)
        out.backward(Gt)
        poolt.zero_grad(set_to_none=True)

    for _ in range(2):  # plain event timing: autograd's worker thread does not take part in a CUDA-graph capture
        train_step(xa[0])
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(6):
        train_step(xa[i % 2])
    t1.record()
    torch.cuda.synchronize()
    ms_train = t0.elapsed_time(t1) / 6
    with ops.KernelTimer(timing=True) as kt:
        train_step(xa[0])
    per_t = {k: round(sum(v), 4) for k, v in kt.durations_ms().items()}
    rep["attentive_pooler_languagebind_16_videos"]["train_step_ms"] = ms_train
    rep["attentive_pooler_languagebind_16_videos"]["train_device_ms_by_entry_point"] = dict(sorted(per_t.items(), key=lambda kv: -kv[1]))
    print("attentive_pooler", rep["attentive_pooler_languagebind_16_videos"])
    json.dump(rep, open(os.path.join(REPO, "gpurun_out", "variants_diag.json"), "w"), indent=1)
except Exception as e:  # noqa: BLE001
    print("attentive pooler diag failed:", repr(e))
