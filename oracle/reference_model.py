"""The reference's CPU implementation of the fusion path, assembled for timing and checking.  TEST / MEASUREMENT
INFRASTRUCTURE ONLY: nothing under ``merv_b200/`` imports this file, and this file imports nothing from ``merv_b200``
(``bench.py --impl reference`` must not load the product's native library).

Two builders, the same seeded construction order as ``MERV.__init__`` (merv/models/vidlms/merv.py:87,152-163,214-216:
``torch.manual_seed(video_backbones[0].embed_dim)``, the projectors in encoder order, then the adapter):

* ``build_reference``: the UNMODIFIED reference modules (``oracle/ref_loader``: /root/reference or the copy staged in
  oracle/_ref) with the glue of merv.py:587-589,607-609 — ``kind: "reference"``;
* ``build_port``: plain ``torch.nn`` parameter containers + ``oracle/torch_port.py`` (the same ATen operators in the same
  order) — ``kind: "port"``, used only where the reference file is not available.

Both give bit-identical initial weights (the reference's modules are nn.Linear / nn.MultiheadAttention / xavier_uniform_
containers too), checked by tests/test_oracle.py.
"""

from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import ref_loader, torch_port

EMBED_DIM = 3072  # merv.py:215 hard-codes embed_dim=3072 for cross_attention_avg_lq


def _seed(vision_dims: Sequence[int]) -> None:
    torch.manual_seed(vision_dims[0])  # merv.py:87


def build_reference(vision_dims: Sequence[int], llm_dim: int, out_frames: Sequence[int], out_size: int, mlp_type: str,
                    token_length: int, q_scale: float = 1.0, dtype: torch.dtype = torch.float32):
    """(forward(feats) -> (prefix, weights), modules) over the real reference classes."""
    ref = ref_loader.load_reference_nn_utils()
    _seed(vision_dims)
    projs = nn.ModuleList([ref.AveragePooling3DProjector(c, llm_dim, output_frames=t, output_size=out_size, mlp_type=mlp_type)
                           for c, t in zip(vision_dims, out_frames)])
    ff = ref.CrossAttentionAdapterLearnableQuery(embed_dim=EMBED_DIM, llm_dim=llm_dim, token_length=token_length, averagetoken=True)
    with torch.no_grad():
        ff.Q.mul_(q_scale)
    projs, ff = projs.to(dtype).eval().requires_grad_(False), ff.to(dtype).eval().requires_grad_(False)

    @torch.no_grad()
    def forward(feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        projected = [p(x) for p, x in zip(projs, feats)]  # merv.py:587-589
        out, w = ff(projected)                             # merv.py:607-608
        return out, w

    return forward, (projs, ff)


def port_state(vision_dims: Sequence[int], llm_dim: int, mlp_type: str, q_scale: float = 1.0, dtype: torch.dtype = torch.float32):
    """Projector / adapter state dicts from plain torch.nn containers, consuming the RNG exactly as the reference does."""
    _seed(vision_dims)
    pps: List[dict] = []
    for c in vision_dims:
        if mlp_type == "linear":
            inner = nn.Linear(c, llm_dim, bias=True)
            sd = {"projector.weight": inner.weight, "projector.bias": inner.bias}
        elif mlp_type == "gelu-mlp":
            seq = nn.Sequential(nn.Linear(c, llm_dim, bias=True), nn.GELU(), nn.Linear(llm_dim, llm_dim, bias=True))
            sd = {f"projector.{k}": v for k, v in seq.state_dict().items()}
        else:
            raise ValueError(f"Projector with `{mlp_type = }` is not supported!")
        pps.append({k: v.detach().to(dtype) for k, v in sd.items()})
    att = nn.MultiheadAttention(embed_dim=EMBED_DIM, num_heads=1, dropout=0.0, batch_first=True, kdim=llm_dim, vdim=llm_dim)  # nn_utils.py:464-471
    Q = torch.empty((1, EMBED_DIM))
    nn.init.xavier_uniform_(Q)  # nn_utils.py:472,483-485
    fp = {f"attention.{k}": v.detach().to(dtype) for k, v in att.state_dict().items()}
    fp["Q"] = (Q * q_scale).to(dtype)
    return pps, fp


def build_port(vision_dims: Sequence[int], llm_dim: int, out_frames: Sequence[int], out_size: int, mlp_type: str,
               token_length: int, q_scale: float = 1.0, dtype: torch.dtype = torch.float32):
    pps, fp = port_state(vision_dims, llm_dim, mlp_type, q_scale, dtype)

    def forward(feats: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        return torch_port.fusion_forward(list(feats), pps, fp, list(out_frames), out_size, mlp_type, token_length)

    return forward, (pps, fp)


def build_cpu_arm(vision_dims, llm_dim, out_frames, out_size, mlp_type, token_length, q_scale=1.0,
                  dtype: torch.dtype = torch.float32) -> Tuple[Callable, str, str]:
    """(forward, kind, what): the real reference when its file is reachable, the port otherwise."""
    if ref_loader.reference_available():
        fwd, _ = build_reference(vision_dims, llm_dim, out_frames, out_size, mlp_type, token_length, q_scale, dtype)
        return fwd, "reference", ("unmodified merv/util/nn_utils.py modules (AveragePooling3DProjector x E + CrossAttentionAdapterLearnableQuery) "
                                  f"from {ref_loader.reference_origin()}, glue of merv.py:587-589,607-609")
    fwd, _ = build_port(vision_dims, llm_dim, out_frames, out_size, mlp_type, token_length, q_scale, dtype)
    return fwd, "port", "oracle/torch_port.py: the reference's ATen operator sequence over torch.nn-initialised weights (reference file not on this box)"
