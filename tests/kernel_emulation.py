"""TEST INFRASTRUCTURE ONLY: torch-CPU stand-ins for the entry points of ``merv_b200.ops``.

The product has no CPU path (``merv_b200.ops`` rejects CPU tensors and the modules call ``_require_device``).  The build
container has no GPU, though, and the host-side logic of ``merv_b200/nn_utils.py`` — which module calls which entry point
with which tensors, the autograd wiring of the hand-written backward, the caches, the state-dict plumbing — deserves CPU
coverage (``pytest -m "not gpu"``).  ``emulate(monkeypatch)`` swaps every ``ops`` function the modules use for a plain
torch implementation OF THE ENTRY POINT'S CONTRACT (as documented in include/merv_fusion.h) and disables the device check,
so the module code runs unchanged on CPU tensors.  Results are then compared with the goldens recorded from the reference.
Nothing under ``merv_b200/`` imports this file; the GPU parity tests (``-m gpu``) never use it.
"""

from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _f(t):
    return None if t is None else t.float()


def pool3d(xs, out_frames, out_size, score_vecs=None, max_ctas=0, batch_index=None, shifts=None, outs=None):
    ys, partials = [], ([] if score_vecs is not None else None)
    for i, (x, T) in enumerate(zip(xs, out_frames)):
        if batch_index is not None:
            x = x[batch_index.long()]
        B, Fr, N, C = x.shape
        H = int(math.sqrt(N))
        v = x.float().reshape(B, Fr, H, H, C).permute(0, 4, 1, 2, 3)
        if shifts is not None:  # windows read x displaced by (sf, sh, sw); zero outside the grid, divisor = the full window
            sf, sh, sw = shifts[i]
            v = F.pad(v, (1, 1, 1, 1, 1, 1))[:, :, 1 + sf:1 + sf + Fr, 1 + sh:1 + sh + H, 1 + sw:1 + sw + H]
        y = F.adaptive_avg_pool3d(v, (T, out_size, out_size)).permute(0, 2, 3, 4, 1).reshape(B, T * out_size**2, C).to(x.dtype)
        if outs is not None:
            outs[i].copy_(y)
            y = outs[i]
        ys.append(y)
        if score_vecs is not None:  # one partial per video (the real kernel emits several; only their sum is contractual)
            partials.append((y.float() @ score_vecs[i].float()).sum(1, keepdim=True))
    return ys, partials


def linear_bias_act(a, w, bias, act=0, rowdot_vec=None):
    y = F.linear(a.float(), w.float(), _f(bias))
    if act == 1:
        y = F.gelu(y)
    y = y.to(a.dtype)
    rd = None
    if rowdot_vec is not None:
        rd = (y.float().reshape(-1, y.shape[-1]) * rowdot_vec.float()).sum(-1, keepdim=True)
    return y, rd


def fusion_query_vec(Q, Wq, Wk, in_proj_bias):
    embed = Wk.shape[0]
    q = Q.float().reshape(1, embed) @ Wq.float().T
    if in_proj_bias is not None:
        q = q + in_proj_bias.float()[:embed]
    return (q @ Wk.float())[0] / math.sqrt(embed)


def affine_score_vec(W, bias, u):
    v = W.float().T @ u
    c = (u * bias.float()).sum().reshape(1) if bias is not None else torch.zeros(1)
    return v, c


def scores_from_tokens(Vs, u, token_length, per_token_u=False, consts=None, mean=True):
    cols = []
    for v in Vs:
        vf = v.float()
        if per_token_u:
            if vf.shape[1] == 1:
                vf = vf.expand(-1, token_length, -1)
            s = (vf * u.reshape(token_length, -1)).sum((1, 2))
            if mean:
                s = s / token_length
        else:
            s = (vf @ u).sum(1)
            if mean:
                s = s / vf.shape[1]
        cols.append(s)
    scores = torch.stack(cols, 1)
    return scores + consts if consts is not None else scores


def score_consts(pe, u, c_in=None):
    c = pe.float() @ u
    if c_in is not None:
        c = c + torch.stack([ci.reshape(()) if ci is not None else torch.zeros(()) for ci in c_in])
    return c


def scores_from_partials(partials, consts, B, T):
    cols = []
    for e, p in enumerate(partials):
        s = p.reshape(B, -1).sum(1) / T
        if consts is not None and consts[e] is not None:
            s = s + consts[e].reshape(())
        cols.append(s)
    return torch.stack(cols, 1)


def softmax_weights(scores, biases=None, N=0, out=None):
    w = torch.softmax(scores.float(), -1)
    if out is not None:
        out.copy_(w)
        w = out
    bias_mix = None
    if biases is not None:
        bias_mix = torch.zeros((scores.shape[0], N))
        for e, b in enumerate(biases):
            if b is not None:
                bias_mix += w[:, e:e + 1] * b.float()[None, :]
    return w, bias_mix


def softmax_mix(Vs, token_length, scores=None, weights=None):
    w = torch.softmax(scores.float(), -1) if weights is None else weights
    out = sum(w[:, e, None, None] * (v.float() if v.shape[1] == token_length else v.float().expand(-1, token_length, -1))
              for e, v in enumerate(Vs))
    return out.to(Vs[0].dtype), w


def fused_linear_mix(As, Ws, scale, bias_mix, rows_per_video, out=None, max_ctas=0, peer_out_ptrs=None):
    M, N = As[0].reshape(-1, As[0].shape[-1]).shape[0], Ws[0].shape[0]
    B = M // rows_per_video
    acc = torch.zeros((B, rows_per_video, N))
    for s, (a, w) in enumerate(zip(As, Ws)):
        acc += scale[:, s, None, None] * (a.float().reshape(B, rows_per_video, -1) @ w.float().T)
    if bias_mix is not None:
        acc += bias_mix[:, None, :]
    res = acc.to(torch.bfloat16)
    if out is None:
        return res.reshape(M, N)
    out.copy_(res if out.dim() == 3 else res.reshape(M, N))
    return out


def concat_linear(Ys, weight, bias, out=None):
    y = F.linear(torch.cat([t.float() for t in Ys], -1), weight.float(), _f(bias)).to(torch.bfloat16)
    if out is not None:
        out.copy_(y.reshape(out.shape))
        return out
    return y


def layernorm(xs, weight, bias, eps=1e-5):
    x = torch.cat([t.float() for t in xs], -1)
    return F.layer_norm(x, (x.shape[-1],), _f(weight), _f(bias), eps).to(xs[0].dtype)


def layernorm_backward(xs, dy, weight, eps=1e-5, need_dx=True, need_params=True):
    x = torch.cat([t.float().reshape(-1, t.shape[-1]) for t in xs], -1).requires_grad_(True)
    w = weight.float().clone().requires_grad_(True)
    b = torch.zeros_like(w).requires_grad_(True)
    with torch.enable_grad():
        y = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    dx, dw, db = torch.autograd.grad(y, (x, w, b), dy.float().reshape(y.shape))
    dt = xs[0].dtype
    return (dx.to(dt) if need_dx else None), (dw.to(dt) if need_params else None), (db.to(dt) if need_params else None)


def transpose(x, pad=False, row_scale=None, rows_per_scale=1):
    xf = x.float()
    if row_scale is not None:
        xf = xf * row_scale.float().repeat_interleave(rows_per_scale)[:x.shape[0], None]
    y = xf.T.contiguous().to(x.dtype)
    vec = 8 if x.dtype == torch.bfloat16 else 4
    extra = (-y.shape[1]) % vec if pad else 0
    return F.pad(y, (0, extra)) if extra else y


def cross_attention(q, kv, batches, heads, scale=None):
    Cc = q.shape[-1]
    hd = Cc // heads
    n_kv = kv.shape[0] // batches
    scale = hd ** -0.5 if scale is None else scale
    qf = q.float().expand(batches, -1, -1) if q.dim() == 2 else q.float()
    k = kv.float()[:, :Cc].reshape(batches, n_kv, heads, hd).permute(0, 2, 1, 3)
    v = kv.float()[:, Cc:].reshape(batches, n_kv, heads, hd).permute(0, 2, 1, 3)
    qh = qf.reshape(batches, -1, heads, hd).permute(0, 2, 1, 3)
    att = torch.softmax((qh @ k.transpose(-1, -2)) * scale, -1)
    return (att @ v).permute(0, 2, 1, 3).reshape(batches, -1, Cc).to(q.dtype)


def cross_attention_backward(q, kv, dout, batches, heads, scale=None):
    """autograd through the emulated forward (shared queries are expanded so every batch entry gets its own dQ)"""
    C_ = q.shape[-1]
    n_q = q.shape[-2]
    with torch.enable_grad():
        qq = (q.float().unsqueeze(0).expand(batches, n_q, C_) if q.dim() == 2 else q.float()).clone().requires_grad_(True)
        kk = kv.float().clone().requires_grad_(True)
        n_kv, hd = kv.shape[0] // batches, C_ // heads
        k = kk.view(batches, n_kv, 2, heads, hd)
        qh = qq.view(batches, n_q, heads, hd).permute(0, 2, 1, 3)
        att = torch.softmax(qh @ k[:, :, 0].permute(0, 2, 3, 1) * (hd ** -0.5 if scale is None else scale), -1)
        out = (att @ k[:, :, 1].permute(0, 2, 1, 3)).permute(0, 2, 1, 3).reshape(batches, n_q, C_)
        dq, dkv = torch.autograd.grad(out, [qq, kk], dout.float())
    return dq.to(q.dtype), dkv.to(kv.dtype)


def add_rows(a, b):
    return (a.float().reshape(-1, b.shape[0], a.shape[1]) + b.float()).reshape(a.shape).to(a.dtype)


def video_colsum(x, scale=1.0):
    return x.float().sum(1) * scale


def pair_dot(x, y, scale=None):
    B = x.shape[0]
    full = (x.float().reshape(B, -1) * y.float().reshape(B, -1)).sum(1, keepdim=True)
    part = torch.cat([full, torch.zeros(B, 15)], 1)  # the real kernel emits several partials; only their sum is contractual
    if scale is None:
        return part
    return part, (y.float() * scale.float().reshape(B, *([1] * (y.dim() - 1)))).to(y.dtype)


def gemm_ex(a, w, bias=None, act=0, a_t=False, w_t=False):
    A = a.float().T if a_t else a.float()
    W = w.float().T if w_t else w.float()
    y = A @ W.T
    if bias is not None:
        y = y + bias.float()
    if act == 1:
        y = F.gelu(y)
    return y.to(a.dtype)


def wgrad_video(dy, x, scale, w, videos):
    T = dy.shape[0] // videos
    G = torch.einsum("btn,btc->bnc", dy.float().reshape(videos, T, -1), x.float().reshape(videos, T, -1))
    dW = (scale.float().reshape(-1, 1, 1) * G).sum(0).to(dy.dtype)
    partial = torch.zeros((videos, 32))
    partial[:, 0] = (G * w.float()).sum((1, 2))
    return dW, partial


def fused_backward(weights, dweights_out, u, gsum, dw_partials, pbars, Ws, biases, Q, Wq, Wk, in_proj_bias, dWs):
    """The contract of merv_fused_backward (include/merv_fusion.h), restated with torch ops."""
    B, E = weights.shape
    dt = Ws[0].dtype
    embed = Wk.shape[0]
    dw = torch.stack([p.sum(1) for p in dw_partials], 1)
    for e, b in enumerate(biases):
        if b is not None:
            dw[:, e] += gsum @ b.float()
    if dweights_out is not None:
        dw = dw + dweights_out
    ds = weights * (dw - (weights * dw).sum(1, keepdim=True))
    du = torch.zeros_like(u)
    dbs = []
    for e in range(E):
        g = ds[:, e] @ pbars[e]  # [C_e]
        sig = ds[:, e].sum()
        dWs[e].copy_((dWs[e].float() + torch.outer(u, g)).to(dt))
        dbs.append(None if biases[e] is None else (weights[:, e] @ gsum + u * sig).to(dt))
        du = du + Ws[e].float() @ g + (sig * biases[e].float() if biases[e] is not None else 0.0)
    rs = 1.0 / math.sqrt(embed)
    q = Wq.float() @ Q.float().reshape(-1) + (in_proj_bias.float()[:embed] if in_proj_bias is not None else 0.0)
    dq = (Wk.float() @ du) * rs
    dWk = torch.outer(q, du) * rs
    dWq = torch.outer(dq, Q.float().reshape(-1))
    dQ = (Wq.float().T @ dq).reshape(1, embed)
    dbias = torch.cat([dq, torch.zeros(2 * embed)])
    return ds, dbs, dQ.to(dt), dWq.to(dt), dWk.to(dt), dbias.to(dt)


def gelu(z, dy=None):
    if dy is None:
        return F.gelu(z.float()).to(z.dtype)
    zf = z.float()
    cdf = 0.5 * (1.0 + torch.erf(zf / math.sqrt(2.0)))
    return (dy.float() * (cdf + zf * torch.exp(-0.5 * zf * zf) / math.sqrt(2.0 * math.pi))).to(z.dtype)


def colsum(x):
    return x.float().sum(0).to(x.dtype)


def mix_backward(Vs, dout, weights, dweights, u, Q, Wq, Wk, in_proj_bias):
    embed = Wk.shape[0]
    dt = Vs[0].dtype
    leaves = [v.float().clone().requires_grad_(True) for v in Vs]
    Qf, Wqf, Wkf, bf = (t.float().clone().requires_grad_(True) for t in (Q, Wq, Wk, in_proj_bias))
    with torch.enable_grad():
        q = Qf.reshape(1, embed) @ Wqf.T + bf[:embed]
        uu = (q @ Wkf)[0] / math.sqrt(embed)
        scores = torch.stack([(v @ uu).mean(1) for v in leaves], 1)
        w = torch.softmax(scores, -1)
        out = sum(w[:, e, None, None] * v for e, v in enumerate(leaves))
        loss = (out * dout.float()).sum() + (0.0 if dweights is None else (w * dweights).sum())
    g = torch.autograd.grad(loss, leaves + [Qf, Wqf, Wkf, bf])
    E = len(Vs)
    return [t.to(dt) for t in g[:E]], g[E].reshape(1, embed).to(dt), g[E + 1].to(dt), g[E + 2].to(dt), g[E + 3].to(dt)


class FusedLinearPlan:
    def __init__(self, xs, out_frames, out_size, vs, cs, Ws, biases, B, src_B):
        self.a = (list(out_frames), out_size, list(vs), list(cs), list(Ws), list(biases))
        self.B = B
        self.x_meta = [(tuple(x.shape), tuple(x.stride())) for x in xs]

    def rebind(self, vs, cs, Ws, biases):
        self.a = (*self.a[:2], list(vs), list(cs), list(Ws), list(biases))

    def run(self, xs, out, batch_index, xs_checked=False):
        assert [(tuple(x.shape), tuple(x.stride())) for x in xs] == self.x_meta, "plan re-run with different shapes / strides"
        out_frames, out_size, vs, cs, Ws, biases = self.a
        pooled, partials = pool3d(xs, out_frames, out_size, score_vecs=vs, batch_index=batch_index)
        T = out_frames[0] * out_size**2
        scores = scores_from_partials(partials, cs, self.B, T)
        w, bias_mix = softmax_weights(scores, biases, Ws[0].shape[0])
        res = fused_linear_mix(pooled, Ws, w, bias_mix, T, out=out)
        if out is None:
            res = res.view(self.B, T, -1)
        return res, w.to(torch.bfloat16)


_NAMES = ["pool3d", "linear_bias_act", "fusion_query_vec", "affine_score_vec", "scores_from_tokens", "score_consts", "scores_from_partials",
          "softmax_weights", "softmax_mix", "fused_linear_mix", "concat_linear", "layernorm", "layernorm_backward", "transpose", "gelu",
          "colsum", "mix_backward", "FusedLinearPlan", "video_colsum", "pair_dot", "gemm_ex", "fused_backward", "cross_attention", "cross_attention_backward", "add_rows", "wgrad_video"]


def emulate(monkeypatch) -> None:
    """Route the modules of merv_b200.nn_utils to the stand-ins above for the duration of one test."""
    from merv_b200 import nn_utils, ops

    for name in _NAMES:
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(nn_utils, "_require_device", lambda t: None)
    monkeypatch.setattr(nn_utils, "_stream_key", lambda dev: (-1, 0))
