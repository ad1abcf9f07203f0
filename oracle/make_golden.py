"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.  Run in the build container:

    python -m oracle.make_golden            # needs /root/reference (read-only), writes tests/golden/

TEST INFRASTRUCTURE ONLY.  For every case in oracle/cases.py it instantiates the reference's
``AveragePooling3DProjector`` / ``CrossAttentionAdapterLearnableQuery`` exactly as
merv/models/vidlms/merv.py:152-163,214-216 does, loads the seeded numpy weights through
``load_state_dict`` (reference state-dict keys), runs merv.py:587-589 + :607-609 on the seeded numpy
features and stores the reference's outputs.  Small cases are stored in full; the full-size cases as
a digest (mixing weights, sums, 4096 sampled elements).  Also records, for the reference's own
seed-1024 initialisation (merv.py:87), per-parameter digests so the product modules can prove they
initialise identically.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import cases as C  # noqa: E402
from oracle.ref_loader import load_reference_nn_utils  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def build_reference(ref, case: C.Case, dtype=torch.float32):
    Proj = ref.AveragePooling3DProjector if case.resampler == "3davg" else ref.AveragePoolingProjector  # merv.py:116-141
    projs = torch.nn.ModuleList(
        [
            Proj(c, case.llm_dim, output_frames=t, output_size=case.out_size, mlp_type=case.mlp_type)
            for c, t in zip(case.dims, case.out_frames)
        ]
    )
    if case.fusion == "scalar":
        fusion = ref.ScalarAdapter(case.num_encoders)  # merv.py:224-225
    elif case.fusion == "concat_channel":
        fusion = ref.LinearProjector(case.num_encoders * case.llm_dim, case.llm_dim)  # merv.py:217-218
    else:
        fusion = ref.CrossAttentionAdapterLearnableQuery(
            embed_dim=case.embed_dim, llm_dim=case.llm_dim, token_length=case.token_length, averagetoken=True,
            num_encoder=case.num_encoders,
        )
    pp, fp = C.make_projector_params(case), C.make_fusion_params(case)
    for proj, p in zip(projs, pp):
        proj.projector.load_state_dict({k: torch.from_numpy(v) for k, v in p.items()})
    fusion.load_state_dict({k: torch.from_numpy(v) for k, v in fp.items()})
    return projs.to(dtype).eval(), fusion.to(dtype).eval(), pp, fp


@torch.no_grad()
def run_reference(projs, fusion, feats, dtype=torch.float32):
    xs = [torch.from_numpy(f).to(dtype) for f in feats]
    ys = [p(x) for p, x in zip(projs, xs)]  # merv.py:587-589
    if isinstance(fusion, type(projs[0].projector)) and not hasattr(fusion, "attention"):  # concat_channel: merv.py:603-606
        out, w = fusion(torch.concat(ys, -1)), torch.zeros((xs[0].shape[0], 0), dtype=dtype)
    else:
        out, w = fusion(ys)  # merv.py:607-609
    pooled = []
    for p, x in zip(projs, xs):  # the intermediate of nn_utils.py:328-329, for the pool-kernel parity test
        import einops
        H = int(x.shape[2] ** 0.5)
        if hasattr(p, "avg_pooling"):
            v = einops.rearrange(x, "B F (H W) C -> B C F H W", H=H)
            pooled.append(einops.rearrange(p.avg_pooling(v), "B C F H W -> B (F H W) C").contiguous())
        else:  # 2-D per-frame resampler, nn_utils.py:157-164
            v = einops.rearrange(x, "B F (H W) C -> (B F) C H W", H=H)
            pooled.append(einops.rearrange(p.avg_pool(v), "(B F) C H W -> B (F H W) C", F=x.shape[1]).contiguous())
    return out, w, ys, pooled


def build_variant(ref, name: str, dtype=torch.float32):
    """The reference module of one oracle/variants.py entry, weights loaded through its own state-dict keys."""
    from oracle import variants as V

    v = V.VARIANTS[name]
    inputs, params = V.make_variant(name)
    if v["kind"] == "projector":
        cls = getattr(ref, v["cls"])
        mod = cls(v["vision_dim"], v["llm_dim"], pre_proj_layernorm=True)  # as merv.py:165-171 does
    elif v["kind"] == "attntv":  # merv.py:124-130
        mod = ref.AttentivePooler(v["C"], v["llm_dim"], num_query_tokens=v["queries"], num_heads=v["heads"], output_frames=v["F"],
                                  mlp_type=v["mlp_type"])
    elif v["kind"] == "conv3d":  # merv.py:142-150
        mod = ref.Convolutional3DProjector(v["C"], v["llm_dim"], output_frames=v["T"], output_size=v["S"], mlp_type=v["mlp_type"])
    elif v["kind"] == "concat_channel_ln":  # merv.py:219-223
        mod = torch.nn.Sequential(torch.nn.LayerNorm(v["E"] * v["K"]), ref.LinearProjector(v["E"] * v["K"], v["K"]))
    else:
        mod = ref.CrossAttentionAdapterLearnableQuery(embed_dim=v["embed"], llm_dim=v["K"], token_length=v["T"],
                                                      averagetoken=v["averagetoken"], num_encoder=v["E"], positional_embedding=v["pe"])
    mod.load_state_dict({k: torch.from_numpy(a) for k, a in params.items()})
    return mod.to(dtype).eval(), inputs, v


@torch.no_grad()
def run_variant(mod, inputs, v, dtype=torch.float32):
    xs = [torch.from_numpy(a).to(dtype) for a in inputs]
    if v["kind"] in ("projector", "attntv", "conv3d"):
        return mod(xs[0]), None
    if v["kind"] == "concat_channel_ln":
        return mod(torch.concat(xs, -1)), None  # merv.py:603-606
    return mod(xs)


def make_variants(ref) -> None:
    from oracle import variants as V

    store = {}
    for name in V.VARIANTS:
        mod, inputs, v = build_variant(ref, name)
        out, w = run_variant(mod, inputs, v)
        mod16, _, _ = build_variant(ref, name, torch.bfloat16)
        out16, w16 = run_variant(mod16, inputs, v, torch.bfloat16)
        store[f"{name}.out"] = out.numpy()
        store[f"{name}.out_bf16"] = out16.float().numpy()
        if w is not None:
            store[f"{name}.weights"] = w.numpy()
            store[f"{name}.weights_bf16"] = w16.float().numpy()
        _, params = V.make_variant(name)
        store[f"{name}.checksum"] = np.float64(sum(float(np.abs(a.astype(np.float64)).sum()) for a in inputs)
                                               + sum(float(np.abs(a.astype(np.float64)).sum()) for a in params.values()))
        dev = float(np.abs(store[f"{name}.out_bf16"] - store[f"{name}.out"]).max() / np.abs(store[f"{name}.out"]).max())
        print(f"{name:24s} out {tuple(out.shape)} bf16-vs-fp32={dev:.2e}" + ("" if w is None else f" w[0]={np.round(w.numpy()[0], 4)}"))
    path = os.path.join(GOLDEN_DIR, "variants.npz")
    np.savez_compressed(path, **store)
    print(f"wrote variants.npz ({os.path.getsize(path) / 1e6:.2f} MB)")


def main() -> None:
    ref = load_reference_nn_utils()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = set(sys.argv[1:])  # `python -m oracle.make_golden concat_channel` regenerates just that fixture
    if only == {"variants"}:  # `python -m oracle.make_golden variants`: only tests/golden/variants.npz
        make_variants(ref)
        return
    for name, case in C.CASES.items():
        if only and name not in only:
            continue
        feats = C.make_features(case)
        projs, fusion, pp, fp = build_reference(ref, case)
        out, w, ys, pooled = run_reference(projs, fusion, feats)
        projs16, fusion16, _, _ = build_reference(ref, case, torch.bfloat16)
        out16, w16, _, _ = run_reference(projs16, fusion16, feats, torch.bfloat16)
        out_np, out16_np = out.numpy(), out16.float().numpy()
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        meta = dict(
            weights=w.numpy(), weights_bf16=w16.float().numpy(),
            out_sum=np.float64(out.double().sum().item()), out_abs_mean=np.float64(out.double().abs().mean().item()),
            out_abs_max=np.float64(out.abs().max().item()),
            ref_bf16_vs_fp32=np.float64(np.abs(out16_np - out_np).max() / np.abs(out_np).max()),
            y_sums=np.array([y.double().sum().item() for y in ys]),
            pooled_sums=np.array([p.double().sum().item() for p in pooled]),
        )
        # inputs and weights are NOT stored: they regenerate bit-identically from the numpy seeds in
        # oracle/cases.py; a checksum guards against generator drift.
        meta["input_checksum"] = np.float64(sum(float(np.abs(f.astype(np.float64)).sum()) for f in feats))
        meta["param_checksum"] = np.float64(
            sum(float(np.abs(v.astype(np.float64)).sum()) for p in pp for v in p.values())
            + sum(float(np.abs(v.astype(np.float64)).sum()) for v in fp.values())
        )
        idx = C.sample_indices(case)
        meta.update(sample_idx=idx, out_samples=out_np.reshape(-1)[idx], out_samples_bf16=out16_np.reshape(-1)[idx],
                    y_samples=np.stack([y.numpy().reshape(-1)[idx % y.numel()] for y in ys]),
                    pooled_head=[p.numpy()[0, :4, :16] for p in pooled][0])
        if name in C.FULL_STORE_CASES:
            full = {f"y{i}": y.numpy() for i, y in enumerate(ys)}
            full.update({f"pooled{i}": p.numpy() for i, p in enumerate(pooled)})
            np.savez_compressed(path, out=out_np, out_bf16=out16_np, **full, **meta)
        else:
            np.savez_compressed(path, **meta)
        print(f"{name:20s} w[0]={np.round(w.numpy()[0], 4)} sum={meta['out_sum']:.4f} "
              f"bf16-vs-fp32={meta['ref_bf16_vs_fp32']:.2e} -> {os.path.getsize(path) / 1e6:.2f} MB")

    if only:
        return
    make_variants(ref)
    # the reference's own seed-1024 construction order (merv.py:87,152-163,214-216) -> init digests
    torch.manual_seed(1024)
    full = C.CASES["merv_full_b1"]
    projs = [ref.AveragePooling3DProjector(c, 4096, output_frames=16, output_size=8, mlp_type="linear") for c in full.dims]
    fusion = ref.CrossAttentionAdapterLearnableQuery(3072, 4096, 1024, averagetoken=True)
    digest = {}
    for i, p in enumerate(projs):
        for k, v in p.state_dict().items():
            digest[f"projectors.{i}.{k}"] = [list(v.shape), v.double().sum().item(), v.double().abs().sum().item()]
    for k, v in fusion.state_dict().items():
        digest[f"feature_fusion.{k}"] = [list(v.shape), v.double().sum().item(), v.double().abs().sum().item()]
    with open(os.path.join(GOLDEN_DIR, "ref_init_seed1024.json"), "w") as f:
        json.dump(digest, f, indent=1)
    print("wrote ref_init_seed1024.json")


if __name__ == "__main__":
    main()
