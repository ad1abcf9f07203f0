"""Builders shared by the CPU host-logic tests and the GPU parity tests for the configuration variants of
``oracle/variants.py`` (pre_proj_layernorm, concat_channel_ln, averagetoken=False, positional embedding)."""
import torch

from tests.golden_util import load_variant


def build_variant_module(name, dtype=torch.float32, device="cpu"):
    """(module of merv_b200 for one variant with the variant's seeded weights, settings, numpy inputs, golden arrays)."""
    import merv_b200 as M

    v, inputs, params, gold = load_variant(name)
    if v["kind"] == "projector":  # merv.py:165-171 constructs these with pre_proj_layernorm=True
        mod = getattr(M, v["cls"])(v["vision_dim"], v["llm_dim"], pre_proj_layernorm=True)
    elif v["kind"] == "attntv":  # merv.py:124-130
        mod = M.AttentivePooler(v["C"], v["llm_dim"], num_query_tokens=v["queries"], num_heads=v["heads"], output_frames=v["F"],
                                mlp_type=v["mlp_type"])
    elif v["kind"] == "conv3d":  # merv.py:142-150
        mod = M.Convolutional3DProjector(v["C"], v["llm_dim"], output_frames=v["T"], output_size=v["S"], mlp_type=v["mlp_type"])
    elif v["kind"] == "concat_channel_ln":  # merv.py:219-223
        mod = M.ConcatChannelLNFusion(v["E"], v["K"])
    else:
        mod = M.CrossAttentionAdapterLearnableQuery(embed_dim=v["embed"], llm_dim=v["K"], token_length=v["T"],
                                                    averagetoken=v["averagetoken"], num_encoder=v["E"], positional_embedding=v["pe"])
    missing, unexpected = mod.load_state_dict({k: torch.from_numpy(a) for k, a in params.items()}, strict=True)
    assert not missing and not unexpected
    return mod.to(device=device, dtype=dtype).eval().requires_grad_(False), v, inputs, gold


def run_variant(mod, v, inputs, dtype=torch.float32, device="cpu", as_list=True):
    """Call the module the way MERV.forward does; returns (out, weights | None)."""
    xs = [torch.from_numpy(a).to(device).to(dtype) for a in inputs]
    if v["kind"] in ("projector", "attntv", "conv3d"):
        return mod(xs[0]), None
    if v["kind"] == "concat_channel_ln":
        return mod(xs if as_list else torch.concat(xs, -1)), None  # merv.py:603-606 passes the concatenation
    return mod(xs)
