"""Seeded synthetic workloads shared by the golden generator and the tests.  TEST INFRASTRUCTURE ONLY.

All randomness is numpy PCG64 (platform-stable), never torch's generator, so the GPU box regenerates
bit-identical inputs and weights from the seeds without access to /root/reference.

Degenerate-softmax guard (SURVEY.md §7 "hard parts"): with default inits the mixing weights are
~[0.25]*4 and a broken score path would still pass, so every case adds per-encoder DC offsets,
scales Q and randomises ``in_proj_bias`` (zeros by default in nn.MultiheadAttention).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


@dataclass(frozen=True)
class Case:
    name: str
    batch: int
    frames: Tuple[int, ...]  # temporal tokens per encoder at the projector input (after the backbone)
    patches: Tuple[int, ...]  # N_i = H_i * H_i patch tokens per frame
    dims: Tuple[int, ...]  # C_i
    llm_dim: int
    embed_dim: int
    out_frames: Tuple[int, ...]
    out_size: int
    mlp_type: str = "linear"
    q_scale: float = 8.0
    offsets: Tuple[float, ...] = (0.0, 0.5, -0.5, 0.25)
    seed: int = 7
    resampler: str = "3davg"  # "3davg" (AveragePooling3DProjector) | "avg" (AveragePoolingProjector, 2-D per frame)
    fusion: str = "cross_attention_avg_lq"  # | "scalar" (ScalarAdapter) | "concat_channel" (LinearProjector(E*K, K))

    @property
    def num_encoders(self) -> int:
        return len(self.dims)

    @property
    def token_length(self) -> int:
        # merv.py:175-193: every resampling projector must emit visual_feature_length tokens (or 1)
        return max(t * self.out_size * self.out_size for t in self.out_frames)


def _mk(name, **kw) -> Case:
    return Case(name=name, **kw)


# merv-full shapes: frames [16,16,32,16] -> temporal tokens 16 each (ViViT tubelet 2)
MERV_FULL = dict(
    frames=(16, 16, 16, 16),
    patches=(256, 256, 196, 196),
    dims=(1024, 1024, 768, 768),
    llm_dim=4096,
    embed_dim=3072,
    out_frames=(16, 16, 16, 16),
    out_size=8,
)

CASES: Dict[str, Case] = {
    c.name: c
    for c in [
        # small cases stored in full (inputs, weights and reference outputs) in tests/golden/<name>.npz
        _mk("tiny_linear", batch=2, frames=(4, 4, 4, 4), patches=(16, 16, 49, 49), dims=(64, 64, 48, 48),
            llm_dim=128, embed_dim=96, out_frames=(4, 4, 4, 4), out_size=4, mlp_type="linear"),
        _mk("tiny_gelu", batch=2, frames=(4, 4, 4, 4), patches=(16, 16, 49, 49), dims=(64, 64, 48, 48),
            llm_dim=128, embed_dim=96, out_frames=(4, 4, 4, 4), out_size=4, mlp_type="gelu-mlp"),
        _mk("tiny_fused_gelu", batch=1, frames=(4, 4), patches=(36, 25), dims=(32, 24),
            llm_dim=64, embed_dim=48, out_frames=(4, 4), out_size=2, mlp_type="fused-gelu-mlp",
            offsets=(0.3, -0.2)),
        # temporal down-factor ("3davg+frame2": output_frames = temporal_resolution // 2, merv.py:136-163)
        _mk("frame_factor2", batch=3, frames=(8, 8, 8, 8), patches=(64, 64, 36, 36), dims=(96, 96, 80, 80),
            llm_dim=256, embed_dim=128, out_frames=(4, 4, 4, 4), out_size=4, mlp_type="linear"),
        # non-dividing temporal pooling (5 -> 3 frames: overlapping temporal windows) and odd channel counts
        _mk("ragged_windows", batch=2, frames=(5, 5, 7), patches=(25, 81, 49), dims=(40, 72, 56),
            llm_dim=192, embed_dim=64, out_frames=(3, 3, 3), out_size=3, mlp_type="linear",
            offsets=(0.1, -0.4, 0.6)),
        # one encoder: softmax over E=1 -> weights == 1 and out == projected tokens (nn_utils.py:487-521)
        _mk("single_encoder", batch=2, frames=(4,), patches=(196,), dims=(96,),
            llm_dim=256, embed_dim=128, out_frames=(4,), out_size=8, mlp_type="linear", offsets=(0.2,)),
        # the tcgen05 tile sizes: M = B*T multiple of 128, K multiples of 64, N multiple of 256
        _mk("mid_linear", batch=2, frames=(8, 8, 8, 8), patches=(64, 64, 49, 49), dims=(256, 256, 192, 192),
            llm_dim=512, embed_dim=384, out_frames=(8, 8, 8, 8), out_size=4, mlp_type="linear"),
        _mk("mid_gelu", batch=2, frames=(8, 8, 8, 8), patches=(64, 64, 49, 49), dims=(256, 256, 192, 192),
            llm_dim=512, embed_dim=384, out_frames=(8, 8, 8, 8), out_size=4, mlp_type="gelu-mlp"),
        # the 2-D per-frame resampler ("avg", nn_utils.py:136-174) and the scalar mixer (nn_utils.py:524-537)
        _mk("avg2d_linear", batch=2, frames=(4, 4, 4, 4), patches=(64, 64, 49, 49), dims=(64, 64, 48, 48),
            llm_dim=128, embed_dim=96, out_frames=(4, 4, 4, 4), out_size=4, mlp_type="linear", resampler="avg"),
        _mk("scalar_mixer", batch=2, frames=(4, 4, 4, 4), patches=(16, 16, 49, 49), dims=(64, 64, 48, 48),
            llm_dim=128, embed_dim=96, out_frames=(4, 4, 4, 4), out_size=4, mlp_type="linear", fusion="scalar"),
        # feature_fusion == "concat_channel" (merv.py:217-218,603-606): channel concat + LinearProjector(E*llm_dim, llm_dim)
        _mk("concat_channel", batch=2, frames=(8, 8, 8, 8), patches=(64, 64, 49, 49), dims=(256, 256, 192, 192),
            llm_dim=512, embed_dim=384, out_frames=(8, 8, 8, 8), out_size=4, mlp_type="linear", fusion="concat_channel"),
        # full size, one video: stored as a digest (weights, sums, sampled elements), not in full
        _mk("merv_full_b1", batch=1, mlp_type="linear", q_scale=64.0, **MERV_FULL),
        _mk("merv_full_b1_gelu", batch=1, mlp_type="gelu-mlp", q_scale=64.0, **MERV_FULL),
    ]
}

# cases whose reference outputs (prefix, per-encoder tokens, pooled tokens) are stored in full; the rest are digests
FULL_STORE_CASES = ("tiny_linear", "tiny_gelu", "tiny_fused_gelu", "ragged_windows")


def _uniform(rng: np.random.Generator, shape, bound: float) -> np.ndarray:
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def make_features(case: Case, seed: int | None = None, batch: int | None = None) -> List[np.ndarray]:
    """Per-encoder [B, T_i, N_i, C_i] fp32 features ~ N(mu_e, 1), drawn in encoder order."""
    rng = np.random.default_rng(case.seed if seed is None else seed)
    B = case.batch if batch is None else batch
    feats = []
    for f, n, c, mu in zip(case.frames, case.patches, case.dims, case.offsets):
        feats.append(rng.standard_normal((B, f, n, c), dtype=np.float32) + np.float32(mu))
    return feats


def _linear_params(rng, prefix: str, n_in: int, n_out: int) -> Dict[str, np.ndarray]:
    bound = 1.0 / math.sqrt(n_in)
    return {f"{prefix}.weight": _uniform(rng, (n_out, n_in), bound), f"{prefix}.bias": _uniform(rng, (n_out,), bound)}


def make_projector_params(case: Case, seed: int = 1024) -> List[Dict[str, np.ndarray]]:
    """State-dict entries relative to ``AveragePooling3DProjector.projector`` (keys as in SURVEY.md §8b)."""
    rng = np.random.default_rng(seed)
    out = []
    for c in case.dims:
        if case.mlp_type == "linear":
            p = _linear_params(rng, "projector", c, case.llm_dim)
        elif case.mlp_type == "gelu-mlp":
            p = {**_linear_params(rng, "projector.0", c, case.llm_dim),
                 **_linear_params(rng, "projector.2", case.llm_dim, case.llm_dim)}
        elif case.mlp_type == "fused-gelu-mlp":
            p = {**_linear_params(rng, "projector.0", c, 4 * c),
                 **_linear_params(rng, "projector.2", 4 * c, case.llm_dim),
                 **_linear_params(rng, "projector.4", case.llm_dim, case.llm_dim)}
        else:
            raise ValueError(case.mlp_type)
        out.append(p)
    return out


def make_fusion_params(case: Case, seed: int = 2048) -> Dict[str, np.ndarray]:
    """State dict of ``CrossAttentionAdapterLearnableQuery`` (nn_utils.py:456-485), keys as in SURVEY.md §8b
    (or of ``ScalarAdapter``, nn_utils.py:524-527)."""
    rng = np.random.default_rng(seed)
    if case.fusion == "scalar":
        return {"scalar": rng.standard_normal(4).astype(np.float32)}
    if case.fusion == "concat_channel":
        return _linear_params(rng, "projector", case.num_encoders * case.llm_dim, case.llm_dim)
    E, K = case.embed_dim, case.llm_dim
    xav = lambda fo, fi: math.sqrt(6.0 / (fi + fo))  # noqa: E731
    return {
        "Q": _uniform(rng, (1, E), xav(1, E)) * np.float32(case.q_scale),
        "attention.q_proj_weight": _uniform(rng, (E, E), xav(E, E)),
        "attention.k_proj_weight": _uniform(rng, (E, K), xav(E, K)),
        "attention.v_proj_weight": _uniform(rng, (E, K), xav(E, K)),
        "attention.in_proj_bias": _uniform(rng, (3 * E,), 0.05),
        "attention.out_proj.weight": _uniform(rng, (E, E), 1.0 / math.sqrt(E)),
        "attention.out_proj.bias": _uniform(rng, (E,), 0.01),
    }


def sample_indices(case: Case, n: int = 4096, seed: int = 99) -> np.ndarray:
    """Flat indices into the [B, T, K] prefix used by the digest fixtures."""
    total = case.batch * case.token_length * case.llm_dim
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, total, size=n)
    idx[:4] = [0, 1, total - 2, total - 1]
    return idx
