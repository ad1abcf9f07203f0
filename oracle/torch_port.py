"""Torch-CPU port of the reference's fusion path, used ONLY as the timed CPU baseline (bench.py's
``cpu_baseline`` leg and ``--impl reference``) and cross-checked against the goldens in tests/test_oracle.py.

TEST / MEASUREMENT INFRASTRUCTURE — never imported by the product.  The reference is Python and cannot travel
to the GPU box (it is not installable offline: draccus/timm/decord/... are absent), so the CPU baseline is
this port, which issues the SAME ATen operators in the SAME order as the reference modules so that its
host-core timing is representative of the reference's own CPU path:

    permute to B C F H W -> F.adaptive_avg_pool3d -> permute back -> F.linear [-> F.gelu -> F.linear]
        (AveragePooling3DProjector.forward, merv/util/nn_utils.py:320-330; projectors :22-59,86-108)
    torch.stack -> mean over tokens -> F.multi_head_attention_forward(need_weights=True) -> torch.bmm
        (CrossAttentionAdapterLearnableQuery.forward, merv/util/nn_utils.py:487-521)
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F


def _project(x: torch.Tensor, p: Dict[str, torch.Tensor], mlp_type: str) -> torch.Tensor:
    if mlp_type == "linear":
        return F.linear(x, p["projector.weight"], p["projector.bias"])
    if mlp_type == "gelu-mlp":
        return F.linear(F.gelu(F.linear(x, p["projector.0.weight"], p["projector.0.bias"])), p["projector.2.weight"], p["projector.2.bias"])
    if mlp_type == "fused-gelu-mlp":
        h = F.gelu(F.linear(x, p["projector.0.weight"], p["projector.0.bias"]))
        h = F.gelu(F.linear(h, p["projector.2.weight"], p["projector.2.bias"]))
        return F.linear(h, p["projector.4.weight"], p["projector.4.bias"])
    raise ValueError(f"Projector with `{mlp_type = }` is not supported!")


def fusion_forward_autograd(
    features: Sequence[torch.Tensor], projector_params: Sequence[Dict[str, torch.Tensor]], fusion_params: Dict[str, torch.Tensor],
    out_frames: Sequence[int], out_size: int, mlp_type: str, token_length: int,
) -> Tuple[torch.Tensor, torch.Tensor]:
    ys: List[torch.Tensor] = []
    for x, p, t in zip(features, projector_params, out_frames):
        B, Fr, N, Cc = x.shape
        H = int(math.sqrt(N))
        v = x.view(B, Fr, H, N // H, Cc).permute(0, 4, 1, 2, 3)  # "B F (H W) C -> B C F H W" (a view, as einops makes)
        pooled = F.adaptive_avg_pool3d(v, (t, out_size, out_size))
        pooled = pooled.permute(0, 2, 3, 4, 1).reshape(B, t * out_size * out_size, Cc)  # "B C F H W -> B (F H W) C"
        ys.append(_project(pooled, p, mlp_type))
    if "scalar" in fusion_params:  # ScalarAdapter, nn_utils.py:529-537
        stacked = torch.stack(ys, 0)
        w = fusion_params["scalar"].softmax(0)
        return (stacked * w.view(-1, 1, 1, 1)).sum(0), w.unsqueeze(0)
    for emb in ys:
        assert emb.shape[1] == token_length or emb.shape[1] == 1
    B, E = ys[0].shape[0], len(ys)
    embed = fusion_params["Q"].shape[1]
    Q = fusion_params["Q"].repeat(B, 1).unsqueeze(1)
    V = torch.stack([(e.repeat(1, token_length, 1) if e.shape[1] == 1 else e) for e in ys], 1)
    V_ = V.mean(2)
    # the same functional the reference's nn.MultiheadAttention(batch_first=True, kdim != embed_dim) dispatches to
    q, k = Q.transpose(0, 1), V_.transpose(0, 1)
    _, weights = F.multi_head_attention_forward(
        q, k, k, embed, 1, None, fusion_params["attention.in_proj_bias"], None, None, False, 0.0,
        fusion_params["attention.out_proj.weight"], fusion_params["attention.out_proj.bias"], training=False, need_weights=True,
        use_separate_proj_weight=True, q_proj_weight=fusion_params["attention.q_proj_weight"],
        k_proj_weight=fusion_params["attention.k_proj_weight"], v_proj_weight=fusion_params["attention.v_proj_weight"],
    )
    K = V.shape[-1]
    out = torch.bmm(weights, V.reshape(B, E, K * token_length)).reshape(B, token_length, K)
    return out, weights[:, 0]


@torch.no_grad()
def fusion_forward(*args, **kwargs):
    """The timed CPU baseline / forward checker (no autograd graph)."""
    return fusion_forward_autograd(*args, **kwargs)
