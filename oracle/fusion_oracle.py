"""CPU oracle for MERV's multi-encoder feature-fusion hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain-numpy restatement of the reference algorithm.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker / the timed CPU baseline.  The product path
(``merv_b200``) never imports anything under ``oracle/`` and has no CPU fallback.

Parity status: the reference ships no tests, golden vectors or KATs for this path (SURVEY.md §4,
§8c), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: ``oracle/make_golden.py``
exec's the unmodified ``/root/reference/merv/util/nn_utils.py`` (via ``oracle/ref_loader.py``) and
stores its outputs in ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks every function below
against those fixtures.

Every function cites the reference lines it restates (paths relative to /root/reference).
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np

try:  # scipy is in the image; keep a pure-python fallback so the oracle never depends on it
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])


# --------------------------------------------------------------------------------------------
# adaptive 3-D average pooling, channel-last in / channel-last out
# --------------------------------------------------------------------------------------------
def adaptive_windows(n_in: int, n_out: int) -> List[Tuple[int, int]]:
    """Window i of torch's adaptive pooling is [floor(i*n_in/n_out), ceil((i+1)*n_in/n_out)).

    Semantic anchor: ``F.adaptive_avg_pool3d`` as used by ``AveragePooling3DProjector``
    (merv/util/nn_utils.py:317,328).  14 -> 8 gives the overlapping table of SURVEY.md §7.
    """
    return [((i * n_in) // n_out, -((-(i + 1) * n_in) // n_out)) for i in range(n_out)]


def avg_pool3d_tokens(x: np.ndarray, out_frames: int, out_size: int) -> np.ndarray:
    """[B, F, N, C] -> [B, T*S*S, C] with token index t*S*S + i*S + j (j fastest).

    Restates merv/util/nn_utils.py:320-329: H = int(sqrt(N)); "B F (H W) C -> B C F H W";
    AdaptiveAvgPool3d((T, S, S)); "B C F H W -> B (F H W) C".
    """
    B, F, N, C = x.shape
    H = int(math.sqrt(N))
    assert H * H == N, "reference computes H=int(sqrt(N)) and einops then requires H*W == N"
    W = N // H
    xv = x.reshape(B, F, H, W, C)
    wt, wh, ww = adaptive_windows(F, out_frames), adaptive_windows(H, out_size), adaptive_windows(W, out_size)
    acc_dtype = np.float64 if x.dtype == np.float64 else np.float32
    out = np.empty((B, out_frames, out_size, out_size, C), dtype=acc_dtype)
    for t, (f0, f1) in enumerate(wt):
        for i, (h0, h1) in enumerate(wh):
            for j, (w0, w1) in enumerate(ww):
                win = xv[:, f0:f1, h0:h1, w0:w1, :].astype(acc_dtype)
                out[:, t, i, j, :] = win.sum(axis=(1, 2, 3)) / float((f1 - f0) * (h1 - h0) * (w1 - w0))
    return out.reshape(B, out_frames * out_size * out_size, C).astype(x.dtype)


# --------------------------------------------------------------------------------------------
# projectors
# --------------------------------------------------------------------------------------------
def gelu_erf(x: np.ndarray) -> np.ndarray:
    """``nn.GELU()`` with approximate='none' (merv/util/nn_utils.py:48): x * 0.5 * (1 + erf(x / sqrt(2)))."""
    return (x * 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))).astype(x.dtype)


def linear(x: np.ndarray, weight: np.ndarray, bias: np.ndarray) -> np.ndarray:
    """``nn.Linear``: y = x @ W^T + b, W is [out, in] (merv/util/nn_utils.py:25,32)."""
    return x @ weight.T + bias


def layer_norm(x: np.ndarray, weight: np.ndarray, bias: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """``nn.LayerNorm(C)`` over the last dimension (merv/util/nn_utils.py:27,41,68,92; merv/models/vidlms/merv.py:221):
    (x - mean) / sqrt(biased_var + eps) * weight + bias, statistics in at least fp32 (ATen accumulates in fp32)."""
    acc = np.float64 if x.dtype == np.float64 else np.float32
    xf = x.astype(acc)
    mean = xf.mean(axis=-1, keepdims=True)
    var = ((xf - mean) ** 2).mean(axis=-1, keepdims=True)
    return ((xf - mean) / np.sqrt(var + acc(eps)) * weight + bias).astype(x.dtype)


def projector_forward(x: np.ndarray, params: Dict[str, np.ndarray], mlp_type: str) -> np.ndarray:
    """``get_mlp_projector`` variants (merv/util/nn_utils.py:22-59,86-121) and ``MLPDeepProjector`` (:62-83, "deep-gelu-mlp").

    ``params`` uses the reference state-dict keys relative to ``AveragePooling3DProjector.projector``:
    linear -> projector.{weight,bias}; gelu-mlp -> projector.{0,2}.*; fused-gelu-mlp / deep-gelu-mlp -> projector.{0,2,4}.*.
    ``layernorm.{weight,bias}`` present <=> the module was built with ``pre_proj_layernorm=True`` (nn_utils.py:26-32): the
    input is layer-normalised first.
    """
    if "layernorm.weight" in params:
        x = layer_norm(x, params["layernorm.weight"], params["layernorm.bias"])
    if mlp_type == "deep-gelu-mlp":
        mlp_type = "fused-gelu-mlp"  # same Sequential(Linear, GELU, Linear, GELU, Linear) structure, different widths
    if mlp_type == "linear":
        return linear(x, params["projector.weight"], params["projector.bias"])
    if mlp_type == "gelu-mlp":
        h = gelu_erf(linear(x, params["projector.0.weight"], params["projector.0.bias"]))
        return linear(h, params["projector.2.weight"], params["projector.2.bias"])
    if mlp_type == "fused-gelu-mlp":
        h = gelu_erf(linear(x, params["projector.0.weight"], params["projector.0.bias"]))
        h = gelu_erf(linear(h, params["projector.2.weight"], params["projector.2.bias"]))
        return linear(h, params["projector.4.weight"], params["projector.4.bias"])
    if mlp_type == "none":
        return x
    raise ValueError(f"Projector with `{mlp_type = }` is not supported!")


def avgpool3d_projector_forward(
    x: np.ndarray, params: Dict[str, np.ndarray], out_frames: int, out_size: int, mlp_type: str
) -> np.ndarray:
    """``AveragePooling3DProjector.forward`` (merv/util/nn_utils.py:320-330): POOL first, then project."""
    return projector_forward(avg_pool3d_tokens(x, out_frames, out_size), params, mlp_type)


# --------------------------------------------------------------------------------------------
# attentive pooler ("attntv" resampler)
# --------------------------------------------------------------------------------------------
def cross_attention_block_forward(q: np.ndarray, x: np.ndarray, params: Dict[str, np.ndarray], num_heads: int, prefix: str = "cross_attn.") -> np.ndarray:
    """``CrossAttentionBlock.forward`` (merv/util/nn_utils.py:447-451) with ``CrossAttention.forward`` (:393-412) and ``MLP.forward``
    (:425-431, dropout 0):  y = xattn(q, norm1(x));  q = q + y;  q = q + mlp(norm2(q)).

    q [B', n, C] queries, x [B', N, C] keys/values.  Heads split the channel dimension contiguously (``reshape(B, n, heads, C // heads)``,
    :395; ``kv(x).reshape(B, N, 2, heads, C // heads)``, :398: the first C output channels of ``kv`` are K, the last C are V);
    scores are scaled by ``head_dim ** -0.5`` (SDPA default == ``self.scale``, :388,403-407)."""
    g = lambda k: params[prefix + k]  # noqa: E731
    Bq, n, C = q.shape
    N = x.shape[1]
    hd = C // num_heads
    xn = layer_norm(x, g("norm1.weight"), g("norm1.bias"))
    qp = linear(q, g("xattn.q.weight"), g("xattn.q.bias")).reshape(Bq, n, num_heads, hd).transpose(0, 2, 1, 3)
    kv = linear(xn, g("xattn.kv.weight"), g("xattn.kv.bias")).reshape(Bq, N, 2, num_heads, hd).transpose(2, 0, 3, 1, 4)
    k, v = kv[0], kv[1]  # [B', heads, N, hd]
    att = _softmax_last((qp @ k.transpose(0, 1, 3, 2)) * (hd ** -0.5))
    y = (att @ v).transpose(0, 2, 1, 3).reshape(Bq, n, C)
    y = linear(y, g("xattn.proj.weight"), g("xattn.proj.bias"))
    q = q + y
    h = layer_norm(q, g("norm2.weight"), g("norm2.bias"))
    h = gelu_erf(linear(h, g("mlp.fc1.weight"), g("mlp.fc1.bias")))
    return q + linear(h, g("mlp.fc2.weight"), g("mlp.fc2.bias"))


def attentive_pooler_forward(x: np.ndarray, params: Dict[str, np.ndarray], num_heads: int, mlp_type: str) -> np.ndarray:
    """``AttentivePooler.forward`` (merv/util/nn_utils.py:229-238): every frame's N patch tokens are resampled to
    ``num_query_tokens`` tokens by one cross-attention block with learned queries, then projected.
    x [B, F, N, C] -> [B, F * num_query_tokens, llm_dim]; token index = f * num_query_tokens + query (:237)."""
    B, F, N, C = x.shape
    xf = x.reshape(B * F, N, C)
    q = np.broadcast_to(params["query_tokens"], (B * F,) + params["query_tokens"].shape[1:])
    q = cross_attention_block_forward(q, xf, params, num_heads)
    proj = {k[len("projector."):]: a for k, a in params.items() if k.startswith("projector.")}  # keys of the get_mlp_projector module (:206)
    y = projector_forward(q, proj, mlp_type)
    return y.reshape(B, F * y.shape[1], y.shape[2])


# --------------------------------------------------------------------------------------------
# learnable-query cross-attention mixer
# --------------------------------------------------------------------------------------------
def conv3d_projector_forward(x: np.ndarray, params: Dict[str, np.ndarray], out_frames: int, out_size: int, mlp_type: str) -> np.ndarray:
    """Convolutional3DProjector.forward (merv/util/nn_utils.py:359-369), restated in the reference's own order — convolve the UN-pooled
    [F, H, W] grid (nn.Conv3d(C, K, kernel_size=3, stride=1, padding=1), nn_utils.py:350: zero padding, cross-correlation), THEN
    AdaptiveAvgPool3d, then the projector — so that it checks the product's pool-first rearrangement instead of repeating it.
    x [B, F, H*W, C] -> [B, T*S*S, K]."""
    x = np.asarray(x, dtype=np.float64)
    B, F, N, C = x.shape
    H = int(math.isqrt(N))
    assert H * H == N
    w = np.asarray(params["convolution_pooling.0.weight"], dtype=np.float64)  # [K, C, 3, 3, 3]
    b = np.asarray(params["convolution_pooling.0.bias"], dtype=np.float64)
    K = w.shape[0]
    xp = np.zeros((B, F + 2, H + 2, H + 2, C))
    xp[:, 1:-1, 1:-1, 1:-1] = x.reshape(B, F, H, H, C)
    y = np.zeros((B, F, H, H, K)) + b
    for kf in range(3):
        for kh in range(3):
            for kw in range(3):
                y += xp[:, kf:kf + F, kh:kh + H, kw:kw + H] @ w[:, :, kf, kh, kw].T
    pooled = avg_pool3d_tokens(y.reshape(B, F, N, K), out_frames, out_size)  # nn_utils.py:351,368: token index (f, h, w), w fastest
    return projector_forward(pooled, {k[len("projector."):]: v for k, v in params.items() if k.startswith("projector.")}, mlp_type)


def _softmax_last(x: np.ndarray) -> np.ndarray:
    z = x - x.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


def fusion_weights_mha(vbar: np.ndarray, params: Dict[str, np.ndarray]) -> np.ndarray:
    """Attention weights of the single-head ``nn.MultiheadAttention`` exactly as the reference calls it.

    merv/util/nn_utils.py:464-471,499-501,512 with torch's need_weights=True path
    (torch/nn/functional.py multi_head_attention_forward, separate q/k/v projection weights):
    q = Q Wq^T + b_q ; k = Vbar Wk^T + b_k ; w = softmax(q k^T / sqrt(embed)).  ``vbar`` is [B, E, K].
    """
    Q = params["Q"]  # [1, embed]
    embed = Q.shape[1]
    b = params["attention.in_proj_bias"]
    q = Q @ params["attention.q_proj_weight"].T + b[:embed]  # [1, embed]
    k = vbar @ params["attention.k_proj_weight"].T + b[embed : 2 * embed]  # [B, E, embed]
    scores = np.einsum("qd,bed->be", q * (1.0 / math.sqrt(embed)), k)
    return _softmax_last(scores)


def fusion_query_vector(params: Dict[str, np.ndarray]) -> np.ndarray:
    """The input-independent vector u with softmax_e(u . Vbar_e) == ``fusion_weights_mha``.

    u = Wk^T (Wq Q^T + b_q) / sqrt(embed); b_k only adds a per-row constant that cancels in the softmax.
    """
    Q = params["Q"]
    embed = Q.shape[1]
    q = Q @ params["attention.q_proj_weight"].T + params["attention.in_proj_bias"][:embed]
    return (q @ params["attention.k_proj_weight"])[0] / math.sqrt(embed)


def cross_attention_fusion_forward(
    V: Sequence[np.ndarray], params: Dict[str, np.ndarray], token_length: int, averagetoken: bool = True
) -> Tuple[np.ndarray, np.ndarray]:
    """``CrossAttentionAdapterLearnableQuery.forward`` (merv/util/nn_utils.py:487-521).

    assert T in {token_length, 1}; broadcast T==1 encoders (:502); stack to [B, E, T, K] (:503).
    averagetoken=True  (:507-512): key = mean over tokens (+ ``pe`` when the module has positional_embedding=True, :510-511)
    averagetoken=False (:514-518): key = the flattened [T*K] token block (``k_proj_weight`` is [embed, T*K]; ``pe`` unused)
    weights from the MHA; out = sum_e w[b,e] * V[b,e] (:521).  Returns (out [B, T, K], weights [B, E]).
    """
    for emb in V:
        assert emb.shape[1] == token_length or emb.shape[1] == 1, (token_length, [e.shape for e in V])
    Vs = [np.repeat(e, token_length, axis=1) if e.shape[1] == 1 else e for e in V]
    stacked = np.stack(Vs, axis=1)  # [B, E, T, K]
    B, E, T, K = stacked.shape
    if averagetoken:
        key = stacked.mean(axis=2)
        if "pe" in params:
            key = key + params["pe"][None, :, :]
    else:
        key = stacked.reshape(B, E, T * K)
    w = fusion_weights_mha(key, params)
    out = np.einsum("be,betk->btk", w, stacked)
    return out.astype(stacked.dtype), w.astype(stacked.dtype)


def scalar_adapter_forward(V: Sequence[np.ndarray], params: Dict[str, np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """``ScalarAdapter.forward`` (merv/util/nn_utils.py:529-537): softmax over a learnable 4-vector, the same weights for
    every video; returns (sum_e w_e V_e, weights [1, E])."""
    w = _softmax_last(params["scalar"][None, :])
    out = np.einsum("e,ebtk->btk", w[0], np.stack(V, axis=0))
    return out.astype(V[0].dtype), w.astype(V[0].dtype)


def concat_channel_forward(V: Sequence[np.ndarray], params: Dict[str, np.ndarray]) -> np.ndarray:
    """feature_fusion == "concat_channel": ``torch.concat(projected, -1)`` then ``LinearProjector(E*llm_dim, llm_dim)``
    (merv/models/vidlms/merv.py:217-218,603-606; LinearProjector is nn_utils.py:22-32, keys ``projector.{weight,bias}``)."""
    return linear(np.concatenate(list(V), axis=-1), params["projector.weight"], params["projector.bias"])


def concat_channel_ln_forward(V: Sequence[np.ndarray], params: Dict[str, np.ndarray]) -> np.ndarray:
    """feature_fusion == "concat_channel_ln": ``torch.concat(projected, -1)`` (merv.py:603-605), then
    ``Sequential(LayerNorm(E*llm_dim), LinearProjector(E*llm_dim, llm_dim))`` (merv.py:219-223; keys ``0.{weight,bias}``,
    ``1.projector.{weight,bias}``)."""
    x = layer_norm(np.concatenate(list(V), axis=-1), params["0.weight"], params["0.bias"])
    return linear(x, params["1.projector.weight"], params["1.projector.bias"])


def token_concat_forward(V: Sequence[np.ndarray]) -> np.ndarray:
    """feature_fusion == "concat" (merv.py:600-601): ``torch.concat(projected, 1)``; "first" (:598-599) is ``V[0]``."""
    return np.concatenate(list(V), axis=1)


IGNORE_INDEX = -100  # merv/models/vidlms/merv.py:53


def assemble_multimodal(prefix, input_embeddings, attention_mask, labels, multimodal_indices, bos_token_length=1):
    """Restates merv/models/vidlms/merv.py:622-720: ``[BOS | prefix | text]`` rows for the multimodal examples (mask True
    and labels IGNORE_INDEX over the prefix), then the text-only examples padded at the end with zero embeddings (mask
    False, labels IGNORE_INDEX), stacked multimodal-first.  The MERV class itself cannot be imported here,
    so this function is a restatement WITHOUT a reference-run pin; it is plain indexing / concatenation.
    ``prefix`` [len(multimodal_indices), T, K]; returns (fused_embeddings, fused_attention_mask, fused_labels)."""
    mm = np.asarray(multimodal_indices, dtype=np.int64)
    Bt, L, K = input_embeddings.shape
    T = prefix.shape[1]
    bos = bos_token_length
    emb = np.concatenate([input_embeddings[mm, :bos], prefix, input_embeddings[mm, bos:]], axis=1)  # merv.py:633-640
    mask = np.concatenate([attention_mask[mm, :bos], np.ones((len(mm), T), attention_mask.dtype), attention_mask[mm, bos:]], axis=1)
    lab = np.concatenate([labels[mm, :bos], np.full((len(mm), T), IGNORE_INDEX, labels.dtype), labels[mm, bos:]], axis=1)
    uni = np.array([i for i in range(Bt) if i not in set(mm.tolist())], dtype=np.int64)  # merv.py:668-672
    if len(uni) == 0:
        return emb, mask, lab
    padcount = (emb.shape[1] - L) // T  # merv.py:699-701 (== 1)
    u_emb = np.concatenate([input_embeddings[uni]] + [np.zeros((len(uni), T, K), input_embeddings.dtype)] * padcount, axis=1)
    u_mask = np.concatenate([attention_mask[uni]] + [np.zeros((len(uni), T), attention_mask.dtype)] * padcount, axis=1)
    u_lab = np.concatenate([labels[uni]] + [np.full((len(uni), T), IGNORE_INDEX, labels.dtype)] * padcount, axis=1)
    return np.vstack([emb, u_emb]), np.vstack([mask, u_mask]), np.vstack([lab, u_lab])


# --------------------------------------------------------------------------------------------
# whole path, as MERV.forward glues it
# --------------------------------------------------------------------------------------------
def merv_fusion_forward(
    features: Sequence[np.ndarray],
    projector_params: Sequence[Dict[str, np.ndarray]],
    fusion_params: Dict[str, np.ndarray],
    out_frames: Sequence[int],
    out_size: int,
    mlp_type: str,
    token_length: int,
) -> Tuple[np.ndarray, np.ndarray, List[np.ndarray]]:
    """merv/models/vidlms/merv.py:587-589 (per-encoder projector) then :607-609 (feature_fusion).

    ``features[i]`` is [B, T_i, N_i, C_i] as produced by the reshape at merv.py:576-585.
    Returns (prefix [B, T, K], weights [B, E], per-encoder projected tokens).
    """
    ys = [
        avgpool3d_projector_forward(x, p, t, out_size, mlp_type)
        for x, p, t in zip(features, projector_params, out_frames)
    ]
    if "projector.weight" in fusion_params:  # feature_fusion == "concat_channel" (merv.py:217-218): no mixing weights
        return concat_channel_forward(ys, fusion_params), np.zeros((features[0].shape[0], 0), features[0].dtype), ys
    if "scalar" in fusion_params:  # feature_fusion == "scalar" (merv.py:224-225)
        out, w = scalar_adapter_forward(ys, fusion_params)
    else:
        out, w = cross_attention_fusion_forward(ys, fusion_params, token_length)
    return out, w, ys


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """Parity metric of BASELINE.md §5: max|a-b| / max|b| (element-wise relative error is ill-conditioned)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
