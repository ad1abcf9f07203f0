"""CPU tests: the numpy oracle (oracle/fusion_oracle.py) against fixtures produced by the UNMODIFIED
reference modules (oracle/make_golden.py -> tests/golden/*.npz).  The reference has no tests of its own
(SURVEY.md §4), so these fixtures are what pins the oracle."""
import numpy as np
import pytest

from oracle import cases as C
from oracle import fusion_oracle as O
from tests.golden_util import regenerate

FP32_TOL = 1e-5  # BASELINE.json north_star: 1e-5 relative in fp32 (metric: max|a-b| / max|b|)


def test_window_table_14_to_8():
    # 14 -> 8 has overlapping windows of size 2/3/3/2/2/3/3/2
    assert O.adaptive_windows(14, 8) == [(0, 2), (1, 4), (3, 6), (5, 7), (7, 9), (8, 11), (10, 13), (12, 14)]
    assert O.adaptive_windows(16, 8) == [(2 * i, 2 * i + 2) for i in range(8)]
    assert O.adaptive_windows(16, 16) == [(i, i + 1) for i in range(16)]
    assert O.adaptive_windows(5, 3) == [(0, 2), (1, 4), (3, 5)]


def test_token_order_is_f_h_w():
    # token index = f*S*S + h*S + w with w fastest
    F, H, S, Cc = 2, 4, 2, 3
    x = np.zeros((1, F, H * H, Cc), np.float32)
    for f in range(F):
        for h in range(H):
            for w in range(H):
                x[0, f, h * H + w, :] = 100 * f + 10 * (h // 2) + (w // 2)
    y = O.avg_pool3d_tokens(x, F, S)
    expect = [100 * f + 10 * i + j for f in range(F) for i in range(S) for j in range(S)]
    assert np.allclose(y[0, :, 0], expect)


@pytest.mark.parametrize("name", list(C.CASES))
def test_oracle_matches_reference_golden(name):
    case = C.CASES[name]
    g, feats, pp, fp = regenerate(case)
    out, w, ys = O.merv_fusion_forward(feats, pp, fp, case.out_frames, case.out_size, case.mlp_type, case.token_length)
    assert out.shape == (case.batch, case.token_length, case.llm_dim)
    scale = float(g["out_abs_max"])
    if case.fusion == "concat_channel":  # no mixing weights on this path (merv.py:603-606)
        assert w.shape == (case.batch, 0) and g["weights"].shape == (case.batch, 0)
    else:
        assert w.shape == ((1 if case.fusion == "scalar" else case.batch), case.num_encoders)  # ScalarAdapter returns [1, E]
        assert np.allclose(w.sum(-1), 1.0, atol=1e-5)
        assert np.abs(w - g["weights"]).max() < 2e-5
    idx = g["sample_idx"]
    assert np.abs(out.reshape(-1)[idx] - g["out_samples"]).max() / scale < FP32_TOL
    assert abs(float(out.astype(np.float64).sum()) - float(g["out_sum"])) < 1e-4 * max(1.0, abs(float(g["out_abs_mean"])) * out.size) 
    for e, y in enumerate(ys):
        ref = g["y_samples"][e]
        assert np.abs(y.reshape(-1)[idx % y.size] - ref).max() / max(np.abs(ref).max(), 1e-6) < FP32_TOL
    if name in C.FULL_STORE_CASES:
        assert O.rel_err(out, g["out"]) < FP32_TOL
        for e, x in enumerate(feats):
            pooled = O.avg_pool3d_tokens(x, case.out_frames[e], case.out_size)
            assert O.rel_err(pooled, g[f"pooled{e}"]) < FP32_TOL
            assert O.rel_err(ys[e], g[f"y{e}"]) < FP32_TOL


@pytest.mark.parametrize("name", ["tiny_linear", "ragged_windows", "mid_linear"])
def test_query_vector_identity(name):
    # weights == softmax_e(u . mean_t V_e): the algebra the CUDA score kernels rely on
    case = C.CASES[name]
    g, feats, pp, fp = regenerate(case)
    _, w, ys = O.merv_fusion_forward(feats, pp, fp, case.out_frames, case.out_size, case.mlp_type, case.token_length)
    u = O.fusion_query_vector(fp)
    s = np.stack([y.astype(np.float64).mean(1) @ u.astype(np.float64) for y in ys], axis=1)
    e = np.exp(s - s.max(-1, keepdims=True))
    assert np.abs(e / e.sum(-1, keepdims=True) - w).max() < 1e-5


def test_single_encoder_is_identity():
    case = C.CASES["single_encoder"]
    g, feats, pp, fp = regenerate(case)
    out, w, ys = O.merv_fusion_forward(feats, pp, fp, case.out_frames, case.out_size, case.mlp_type, case.token_length)
    assert np.array_equal(w, np.ones_like(w))
    assert np.array_equal(out, ys[0])


def test_token_length_one_broadcast():
    # nn_utils.py:502: an encoder emitting a single token is repeated to token_length
    case = C.CASES["tiny_linear"]
    fp = C.make_fusion_params(case)
    rng = np.random.default_rng(0)
    V = [rng.standard_normal((2, case.token_length, case.llm_dim)).astype(np.float32),
         rng.standard_normal((2, 1, case.llm_dim)).astype(np.float32)]
    out, w = O.cross_attention_fusion_forward(V, fp, case.token_length)
    Vb = [V[0], np.repeat(V[1], case.token_length, axis=1)]
    out2, w2 = O.cross_attention_fusion_forward(Vb, fp, case.token_length)
    assert np.array_equal(out, out2) and np.array_equal(w, w2)
    with pytest.raises(AssertionError):
        O.cross_attention_fusion_forward([V[0][:, :3]], fp, case.token_length)


def test_unsupported_projector_type():
    with pytest.raises(ValueError):
        O.projector_forward(np.zeros((1, 4), np.float32), {}, "relu-mlp")


@pytest.mark.parametrize("name", ["tiny_linear", "tiny_gelu", "tiny_fused_gelu", "ragged_windows", "mid_linear"])
def test_torch_port_cpu_baseline_matches_golden(name):
    # the timed CPU baseline (bench.py cpu_baseline / --impl reference) must compute the same thing
    import torch

    from oracle import torch_port

    case = C.CASES[name]
    g, feats, pp, fp = regenerate(case)
    tt = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}  # noqa: E731
    out, w = torch_port.fusion_forward([torch.from_numpy(f) for f in feats], [tt(p) for p in pp], tt(fp), case.out_frames,
                                       case.out_size, case.mlp_type, case.token_length)
    idx = g["sample_idx"]
    assert np.abs(out.numpy().reshape(-1)[idx] - g["out_samples"]).max() / float(g["out_abs_max"]) < FP32_TOL
    assert np.abs(w.numpy() - g["weights"]).max() < 2e-5


def test_assemble_multimodal_layout():
    # merv.py:622-720: [BOS | prefix | text] rows first, then text-only rows padded at the END; prefix positions IGNORE_INDEX / True
    rng = np.random.default_rng(0)
    Bt, L, T, K = 4, 5, 3, 2
    emb = rng.standard_normal((Bt, L, K)).astype(np.float32)
    mask = rng.random((Bt, L)) > 0.3
    labels = rng.integers(0, 100, (Bt, L))
    prefix = rng.standard_normal((2, T, K)).astype(np.float32)
    fe, fm, fl = O.assemble_multimodal(prefix, emb, mask, labels, [2, 0], 1)
    assert fe.shape == (Bt, L + T, K) and fm.shape == fl.shape == (Bt, L + T)
    assert np.array_equal(fe[0, :1], emb[2, :1]) and np.array_equal(fe[0, 1:1 + T], prefix[0]) and np.array_equal(fe[0, 1 + T:], emb[2, 1:])
    assert fm[:2, 1:1 + T].all() and (fl[:2, 1:1 + T] == O.IGNORE_INDEX).all()
    assert np.array_equal(fe[2, :L], emb[1]) and not fe[2:, L:].any()          # unimodal rows 1 and 3, zero padded at the end
    assert not fm[2:, L:].any() and (fl[2:, L:] == O.IGNORE_INDEX).all() and np.array_equal(fl[3, :L], labels[3])
    fe2, fm2, fl2 = O.assemble_multimodal(np.concatenate([prefix, prefix]), emb, mask, labels, [0, 1, 2, 3], 1)
    assert fe2.shape == (Bt, L + T, K)  # no unimodal rows: fused == multimodal


# ---- configuration variants: pre_proj_layernorm, concat_channel_ln, averagetoken=False, positional embedding ----
def oracle_variant(v, inputs, params):
    if v["kind"] == "projector":
        return O.projector_forward(inputs[0], params, v["mlp_type"]), None
    if v["kind"] == "attntv":
        return O.attentive_pooler_forward(inputs[0], params, v["heads"], v["mlp_type"]), None
    if v["kind"] == "conv3d":
        return O.conv3d_projector_forward(inputs[0], params, v["T"], v["S"], v["mlp_type"]), None
    if v["kind"] == "concat_channel_ln":
        return O.concat_channel_ln_forward(inputs, params), None
    return O.cross_attention_fusion_forward(inputs, params, v["T"], averagetoken=v["averagetoken"])


def _variant_names():
    from oracle import variants as V

    return list(V.VARIANTS)


@pytest.mark.parametrize("name", _variant_names())
def test_oracle_variants_match_reference_golden(name):
    from tests.golden_util import load_variant

    v, inputs, params, gold = load_variant(name)
    out, w = oracle_variant(v, inputs, params)
    assert out.shape == gold["out"].shape
    assert O.rel_err(out, gold["out"]) < FP32_TOL
    if w is not None:
        assert np.abs(w - gold["weights"]).max() < 2e-5 and np.allclose(w.sum(-1), 1.0, atol=1e-5)
        assert np.abs(w - 1.0 / w.shape[1]).max() > 0.03, "degenerate (flat) softmax: the fixture would not catch a broken score path"


def test_variant_fixtures_are_sensitive_to_the_variant():
    # the positional embedding / the flattened key / the LayerNorm must move the reference's output well beyond the parity tolerance
    from tests.golden_util import load_variant

    v, inputs, params, gold = load_variant("xattn_pe")
    no_pe = {k: a for k, a in params.items() if k != "pe"}
    _, w = O.cross_attention_fusion_forward(inputs, no_pe, v["T"])
    assert np.abs(w - gold["weights"]).max() > 0.02
    v, inputs, params, gold = load_variant("pre_ln_linear")
    no_ln = {k: a for k, a in params.items() if not k.startswith("layernorm")}
    assert O.rel_err(O.projector_forward(inputs[0], no_ln, "linear"), gold["out"]) > 0.1


def test_token_concat_and_first():
    rng = np.random.default_rng(0)
    V = [rng.standard_normal((2, 3, 4)).astype(np.float32) for _ in range(3)]
    out = O.token_concat_forward(V)
    assert out.shape == (2, 9, 4) and np.array_equal(out[:, 3:6], V[1])


def test_reference_arm_port_weights_equal_the_reference_modules():
    """bench.py --impl reference builds its weights without merv_b200: the torch.nn containers of oracle/reference_model.port_state
    must consume the RNG exactly like the reference's modules (merv.py:87,152-163,214-216), and both CPU arms must agree."""
    import torch

    from oracle import reference_model as RM
    from oracle.ref_loader import reference_available

    if not reference_available():
        import pytest

        pytest.skip("needs the reference's nn_utils.py")
    dims, llm, frames, S = [64, 48], 96, [4, 4], 2
    for mlp_type in ("linear", "gelu-mlp"):
        fwd_ref, (projs, ff) = RM.build_reference(dims, llm, frames, S, mlp_type, 16, q_scale=8.0)
        pps, fp = RM.port_state(dims, llm, mlp_type, q_scale=8.0)
        for p, sd in zip(projs, pps):
            ref_sd = p.projector.state_dict()
            assert sorted(ref_sd) == sorted(sd) and all(torch.equal(ref_sd[k], sd[k]) for k in sd)
        ref_sd = ff.state_dict()
        assert sorted(ref_sd) == sorted(fp) and all(torch.equal(ref_sd[k], fp[k]) for k in fp)
        fwd_port, _ = RM.build_port(dims, llm, frames, S, mlp_type, 16, q_scale=8.0)
        g = torch.Generator().manual_seed(3)
        feats = [torch.randn((2, 4, n, c), generator=g) + mu for n, c, mu in zip((16, 9), dims, (0.3, -0.4))]
        (o1, w1), (o2, w2) = fwd_ref(feats), fwd_port(feats)
        assert float((o1 - o2).abs().max() / o1.abs().max()) < 1e-5 and float((w1 - w2).abs().max()) < 1e-6
