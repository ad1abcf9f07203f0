"""merv_b200 — B200-native (sm_100a) implementation of MERV's multi-encoder feature-fusion hot path.

Public surface mirrors the reference's ``merv/util/nn_utils.py`` for this path (same class names and
signatures) plus the fused pipeline entry points:

    from merv_b200 import AveragePooling3DProjector, CrossAttentionAdapterLearnableQuery, MervFusion, patch_merv

All compute runs in ``libmerv_fusion.so`` (hand-written CUDA behind the C ABI of ``include/merv_fusion.h``);
importing this package without the built library raises.
"""

from . import _lib

_lib.load()  # fail loudly at import time if the CUDA library has not been built

from .nn_utils import (  # noqa: E402
    AttentivePooler,
    Convolutional3DProjector,
    AveragePooling3DProjector,
    AveragePoolingProjector,
    ConcatChannelFusion,
    ConcatChannelLNFusion,
    CrossAttentionAdapterLearnableQuery,
    DeferredProjection,
    FusedMLPProjector,
    LinearProjector,
    MervFusion,
    MLPDeepProjector,
    MLPProjector,
    ScalarAdapter,
    TokenResampler,
    fsdp_wrap_policy,
    get_mlp_projector,
    link_fused,
    patch_merv,
    unlink,
)

__all__ = [
    "AttentivePooler", "Convolutional3DProjector", "AveragePooling3DProjector", "AveragePoolingProjector", "ScalarAdapter", "ConcatChannelFusion", "ConcatChannelLNFusion", "MLPDeepProjector", "CrossAttentionAdapterLearnableQuery", "DeferredProjection", "FusedMLPProjector",
    "LinearProjector", "MervFusion", "MLPProjector", "TokenResampler", "fsdp_wrap_policy", "get_mlp_projector", "link_fused", "patch_merv", "unlink",
]
