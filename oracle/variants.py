"""Seeded workloads for the configuration variants of the path that sit outside the ``Case`` grid of
``oracle/cases.py``.  TEST INFRASTRUCTURE ONLY.

Each variant exercises ONE reference module directly on seeded numpy inputs (PCG64, platform-stable):

    pre_ln_*            LinearProjector / MLPProjector / MLPDeepProjector / FusedMLPProjector with pre_proj_layernorm=True
                        (merv/util/nn_utils.py:22-108; constructed at merv/models/vidlms/merv.py:165-171)
    concat_channel_ln   Sequential(LayerNorm(E*K), LinearProjector(E*K, K))       (merv.py:219-223,603-606)
    xattn_flat          CrossAttentionAdapterLearnableQuery(averagetoken=False)    (nn_utils.py:514-518), one T==1 encoder
    xattn_pe            CrossAttentionAdapterLearnableQuery(averagetoken=True, positional_embedding=True)  (nn_utils.py:510-511)
    attntv_*            AttentivePooler (cross-attention resampler with learned queries)                    (nn_utils.py:177-246,380-452)
    conv3d_*            Convolutional3DProjector (Conv3d 3x3x3 -> AdaptiveAvgPool3d -> projector)            (nn_utils.py:341-377)
    *_wide              the same at tcgen05 tile sizes / CTA-per-row LayerNorm widths (E*K = 4 * 1024 channels)

``oracle/make_golden.py`` runs the unmodified reference modules on them and stores the outputs in
``tests/golden/variants.npz``; the tests regenerate inputs and weights from the seeds below.
"""

from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np

from .cases import _linear_params, _uniform

# name -> (kind, settings)
VARIANTS: Dict[str, dict] = {
    "pre_ln_linear": dict(kind="projector", cls="LinearProjector", mlp_type="linear", vision_dim=64, llm_dim=128, shape=(2, 50, 64)),
    "pre_ln_gelu": dict(kind="projector", cls="MLPProjector", mlp_type="gelu-mlp", vision_dim=64, llm_dim=128, shape=(2, 50, 64)),
    "pre_ln_deep": dict(kind="projector", cls="MLPDeepProjector", mlp_type="deep-gelu-mlp", vision_dim=48, llm_dim=64, shape=(3, 7, 48)),
    "pre_ln_fused_gelu": dict(kind="projector", cls="FusedMLPProjector", mlp_type="fused-gelu-mlp", vision_dim=40, llm_dim=64, shape=(1, 33, 40)),
    # tcgen05-sized: M = 256 rows, K = 1024 channels (the LanguageBind / DINOv2 width), N = 512
    "pre_ln_linear_wide": dict(kind="projector", cls="LinearProjector", mlp_type="linear", vision_dim=1024, llm_dim=512, shape=(2, 128, 1024)),
    "concat_channel_ln": dict(kind="concat_channel_ln", E=4, T=64, K=128, B=2),
    "concat_channel_ln_wide": dict(kind="concat_channel_ln", E=4, T=128, K=1024, B=1),
    "xattn_flat": dict(kind="xattn", E=4, T=16, K=128, B=3, embed=96, averagetoken=False, pe=False, single=(2,)),
    "xattn_pe": dict(kind="xattn", E=4, T=64, K=128, B=3, embed=96, averagetoken=True, pe=True, single=()),
    "xattn_pe_single": dict(kind="xattn", E=3, T=32, K=256, B=2, embed=64, averagetoken=True, pe=True, single=(0,)),
    # AttentivePooler ("attntv" resampler, merv/util/nn_utils.py:177-246; constructed at merv.py:124-130 with num_heads=8)
    "attntv_tiny": dict(kind="attntv", C=32, llm_dim=48, queries=4, heads=4, F=3, N=9, B=2, mlp_type="gelu-mlp"),
    "attntv_linear": dict(kind="attntv", C=64, llm_dim=128, queries=16, heads=8, F=2, N=49, B=2, mlp_type="linear"),
    # tcgen05-sized GEMMs, head_dim 32, 196 keys (a 14 x 14 frame)
    "attntv_wide": dict(kind="attntv", C=256, llm_dim=512, queries=64, heads=8, F=2, N=196, B=1, mlp_type="linear"),
    # the SigLIP / ViViT width: head_dim 96, 64 queries, one 14 x 14 frame
    "attntv_hd96": dict(kind="attntv", C=768, llm_dim=256, queries=64, heads=8, F=1, N=196, B=1, mlp_type="linear"),
    # Convolutional3DProjector ("3dconv" resampler, merv/util/nn_utils.py:341-377; constructed at merv.py:142-150)
    "conv3d_tiny": dict(kind="conv3d", C=16, llm_dim=24, F=3, H=4, T=3, S=2, B=2, mlp_type="gelu-mlp"),
    # ragged windows in every axis (5 frames -> 2, 7 x 7 patches -> 3 x 3: overlapping 3-wide windows) and one frame / row / column of padding each way
    "conv3d_ragged": dict(kind="conv3d", C=24, llm_dim=32, F=5, H=7, T=2, S=3, B=2, mlp_type="linear"),
    # the shipped grid: 16 x 16 patches -> 8 x 8 (the square-grid pool consumer), tcgen05-sized GEMM (M = 256, N = 256, K = 27 * 128)
    "conv3d_wide": dict(kind="conv3d", C=128, llm_dim=256, F=2, H=16, T=2, S=8, B=2, mlp_type="linear"),
    "conv3d_14": dict(kind="conv3d", C=64, llm_dim=128, F=2, H=14, T=2, S=8, B=1, mlp_type="linear"),
}


def make_variant(name: str) -> Tuple[List[np.ndarray], Dict[str, np.ndarray]]:
    """(inputs, params): inputs is a list of fp32 arrays (one tensor for projector variants, E token tensors otherwise);
    params uses the reference module's own state-dict keys."""
    v = VARIANTS[name]
    rng = np.random.default_rng(abs(hash_name(name)))
    if v["kind"] == "projector":
        C, K = v["vision_dim"], v["llm_dim"]
        x = rng.standard_normal(v["shape"], dtype=np.float32) * np.float32(1.7) + np.float32(0.4)
        p = {"layernorm.weight": (1.0 + 0.3 * rng.standard_normal(C)).astype(np.float32),
             "layernorm.bias": (0.2 * rng.standard_normal(C)).astype(np.float32)}
        if v["mlp_type"] == "linear":
            p.update(_linear_params(rng, "projector", C, K))
        elif v["mlp_type"] == "gelu-mlp":
            p.update({**_linear_params(rng, "projector.0", C, K), **_linear_params(rng, "projector.2", K, K)})
        elif v["mlp_type"] == "deep-gelu-mlp":
            p.update({**_linear_params(rng, "projector.0", C, K), **_linear_params(rng, "projector.2", K, K),
                      **_linear_params(rng, "projector.4", K, K)})
        else:
            p.update({**_linear_params(rng, "projector.0", C, 4 * C), **_linear_params(rng, "projector.2", 4 * C, K),
                      **_linear_params(rng, "projector.4", K, K)})
        return [x], p
    if v["kind"] == "attntv":
        C, K, n = v["C"], v["llm_dim"], v["queries"]
        x = rng.standard_normal((v["B"], v["F"], v["N"], C), dtype=np.float32) * np.float32(1.3) + np.float32(0.2)
        p = {"query_tokens": rng.standard_normal((1, n, C), dtype=np.float32) * np.float32(0.7),
             "cross_attn.norm1.weight": (1.0 + 0.3 * rng.standard_normal(C)).astype(np.float32),
             "cross_attn.norm1.bias": (0.2 * rng.standard_normal(C)).astype(np.float32),
             "cross_attn.norm2.weight": (1.0 + 0.3 * rng.standard_normal(C)).astype(np.float32),
             "cross_attn.norm2.bias": (0.2 * rng.standard_normal(C)).astype(np.float32)}
        # 4x the nn.Linear default bound on q / kv so that the attention is far from uniform (a broken score path must show)
        p.update({k: a * np.float32(4.0) for k, a in _linear_params(rng, "cross_attn.xattn.q", C, C).items()})
        p.update({k: a * np.float32(4.0) for k, a in _linear_params(rng, "cross_attn.xattn.kv", C, 2 * C).items()})
        p.update(_linear_params(rng, "cross_attn.xattn.proj", C, C))
        p.update(_linear_params(rng, "cross_attn.mlp.fc1", C, 4 * C))
        p.update(_linear_params(rng, "cross_attn.mlp.fc2", 4 * C, C))
        if v["mlp_type"] == "linear":
            p.update(_linear_params(rng, "projector.projector", C, K))
        else:
            p.update({**_linear_params(rng, "projector.projector.0", C, K), **_linear_params(rng, "projector.projector.2", K, K)})
        return [x], p
    if v["kind"] == "conv3d":
        C, K = v["C"], v["llm_dim"]
        x = rng.standard_normal((v["B"], v["F"], v["H"] * v["H"], C), dtype=np.float32) * np.float32(1.2) + np.float32(0.3)
        bound = 1.0 / math.sqrt(27 * C)  # nn.Conv3d default init: U(+-1 / sqrt(fan_in))
        p = {"convolution_pooling.0.weight": _uniform(rng, (K, C, 3, 3, 3), bound),
             "convolution_pooling.0.bias": _uniform(rng, (K,), bound)}
        if v["mlp_type"] == "linear":
            p.update(_linear_params(rng, "projector.projector", K, K))
        else:
            p.update({**_linear_params(rng, "projector.projector.0", K, K), **_linear_params(rng, "projector.projector.2", K, K)})
        return [x], p
    E, T, K, B = v["E"], v["T"], v["K"], v["B"]
    offsets = (0.0, 0.5, -0.5, 0.25)
    if v["kind"] == "concat_channel_ln":
        V = [rng.standard_normal((B, T, K), dtype=np.float32) * np.float32(1.0 + 0.5 * e) + np.float32(offsets[e % 4]) for e in range(E)]
        p = {"0.weight": (1.0 + 0.3 * rng.standard_normal(E * K)).astype(np.float32),
             "0.bias": (0.2 * rng.standard_normal(E * K)).astype(np.float32)}
        p.update({"1." + k: w for k, w in _linear_params(rng, "projector", E * K, K).items()})
        return V, p
    # xattn
    embed, avg = v["embed"], v["averagetoken"]
    V = [rng.standard_normal((B, 1 if e in v["single"] else T, K), dtype=np.float32) + np.float32(offsets[e % 4]) for e in range(E)]
    kdim = K if avg else T * K
    xav = lambda fo, fi: math.sqrt(6.0 / (fi + fo))  # noqa: E731
    p = {
        "Q": _uniform(rng, (1, embed), xav(1, embed)) * np.float32(8.0),
        "attention.q_proj_weight": _uniform(rng, (embed, embed), xav(embed, embed)),
        # averagetoken=False sums T*K products: scale the key projection so the softmax stays non-degenerate but not flat
        "attention.k_proj_weight": _uniform(rng, (embed, kdim), xav(embed, K) / (1.0 if avg else math.sqrt(T))),
        "attention.v_proj_weight": _uniform(rng, (embed, kdim), xav(embed, K)),
        "attention.in_proj_bias": _uniform(rng, (3 * embed,), 0.05),
        "attention.out_proj.weight": _uniform(rng, (embed, embed), 1.0 / math.sqrt(embed)),
        "attention.out_proj.bias": _uniform(rng, (embed,), 0.01),
    }
    if v["pe"]:
        p["pe"] = _uniform(rng, (E, K), xav(E, K)) * np.float32(4.0)
    return V, p


def hash_name(name: str) -> int:
    """Stable (not PYTHONHASHSEED-dependent) seed from the variant name."""
    h = 2166136261
    for ch in name.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h
