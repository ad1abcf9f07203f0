"""One rank of the FSDP drop-in test (launched by tests/test_fsdp.py through torch.distributed.run).

Builds a stand-in for ``MERV`` holding the REAL reference modules (``oracle/ref_loader``: the unmodified merv/util/nn_utils.py)
under the sub-module names MERV uses, with MERV's own ``get_fsdp_wrapping_policy`` (merv/models/vidlms/merv.py:465-497), wraps it
in torch FSDP exactly as the reference's training strategy does (merv/training/strategies/fsdp.py:208-241: auto-wrap policy from
the model, bf16 ``MixedPrecision``, ``use_orig_params=True``, ``limit_all_gathers=True``, full sharding), and trains it for a few
SGD steps.  A deep copy goes through ``merv_b200.patch_merv`` BEFORE the wrap and is trained on the same data; a third copy of the
reference runs the same steps under FSDP in fp32 and serves as the yardstick.  Every rank trains on its own micro-batch, so the
reduced gradients partly cancel between ranks and a plain relative tolerance would measure that cancellation, not the kernels.
The criterion is therefore "as accurate as the reference's own bf16 arithmetic": for every parameter, the error of this repo
against the fp32 run must not exceed 2x the error of the reference's bf16 run against it (+ 1.5e-2), for the full gradients of
the first step and for the parameter change after the last step; losses and an eval pass are compared directly (3e-2).  Also
checked: which autograd path ran (``_FusedLinearFn`` only in the root-unit mode) and that FSDP really made the projectors their
own units (per-unit mode, identical to the reference's unit list) or left them to the root unit (root-unit mode).

    --backend nccl   real kernels on one GPU per rank            (pytest -m gpu, needs >= 2 GPUs)
    --backend gloo   CPU, kernels replaced by tests/kernel_emulation.py: covers the host logic only (pytest -m "not gpu")
"""
import argparse
import copy
import functools
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn as nn  # noqa: E402
from torch.distributed.fsdp import FullyShardedDataParallel as FSDP  # noqa: E402
from torch.distributed.fsdp import MixedPrecision, ShardingStrategy  # noqa: E402
from torch.distributed.fsdp.wrap import _module_wrap_policy, _or_policy  # noqa: E402


class _Setattr:  # the part of pytest's monkeypatch kernel_emulation.emulate() uses
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--mode", default="per_unit", choices=["per_unit", "root_unit"])
    ap.add_argument("--mlp-type", default="linear")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--resampler", default="3davg", choices=["3davg", "3dconv"])  # merv.py:135-150
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    gpu = args.backend == "nccl"
    if gpu:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
    else:
        dev = torch.device("cpu")
        dist.init_process_group("gloo")
        from tests import kernel_emulation

        kernel_emulation.emulate(_Setattr())

    import merv_b200 as M
    from merv_b200 import nn_utils as NU
    from oracle.ref_loader import load_reference_nn_utils

    ref = load_reference_nn_utils()
    dims, llm, frames, S, embed = [128, 64, 64], 256, 4, 4, 96
    patches = [64, 49, 64]  # 8 x 8, 7 x 7 (overlapping windows), 8 x 8 -> 4 x 4
    T = frames * S * S

    class FakeMerv(nn.Module):
        """projectors / feature_fusion / a stand-in LLM head, named and wrapped as MERV's (merv.py:152-163,214-216,465-497)."""

        def __init__(self):
            super().__init__()
            torch.manual_seed(dims[0])  # merv.py:87
            Resampler = ref.AveragePooling3DProjector if args.resampler == "3davg" else ref.Convolutional3DProjector
            self.projectors = nn.ModuleList([Resampler(c, llm, output_frames=frames, output_size=S, mlp_type=args.mlp_type) for c in dims])
            self.feature_fusion = ref.CrossAttentionAdapterLearnableQuery(embed_dim=embed, llm_dim=llm, token_length=T, averagetoken=True)
            self.llm_head = nn.Linear(llm, 32)  # stands for the LLM consuming the prefix (its own FSDP unit via the LLM's policy)
            with torch.no_grad():
                self.feature_fusion.Q.mul_(16.0)  # non-degenerate mixing weights

        def get_fsdp_wrapping_policy(self):  # merv.py:465-497 with the LLM policy reduced to the stand-in head
            videolm = functools.partial(_module_wrap_policy, module_classes={
                ref.LinearProjector, ref.MLPProjector, ref.FusedMLPProjector, ref.AveragePoolingProjector, ref.MLPDeepProjector,
                ref.AveragePooling3DProjector, ref.Convolutional3DProjector})
            llm_policy = functools.partial(_module_wrap_policy, module_classes={nn.Linear}) if False else \
                functools.partial(lambda module, recurse, nonwrapped_numel, head: True if recurse else module is head, head=self.llm_head)
            return functools.partial(_or_policy, policies=[llm_policy, videolm])

        def forward(self, feats):  # merv.py:587-589,607-609
            projected = [p(x) for p, x in zip(self.projectors, feats)]
            prefix, weights = self.feature_fusion(projected)
            return self.llm_head(prefix), weights

    base = FakeMerv()
    ours = copy.deepcopy(base)
    M.patch_merv(ours, fused_training=True if args.mode == "root_unit" else None)
    want_cls = M.AveragePooling3DProjector if args.resampler == "3davg" else M.Convolutional3DProjector
    assert isinstance(ours.projectors[0], want_cls) and isinstance(ours.feature_fusion, M.CrossAttentionAdapterLearnableQuery)

    mp = MixedPrecision(param_dtype=torch.bfloat16, reduce_dtype=torch.bfloat16, buffer_dtype=torch.bfloat16)  # fsdp.py:217-224

    def wrap(model):
        return FSDP(model, auto_wrap_policy=model.get_fsdp_wrapping_policy(), mixed_precision=mp, sharding_strategy=ShardingStrategy.FULL_SHARD,
                    device_id=dev if gpu else torch.device("cpu"), limit_all_gathers=True, use_orig_params=True)  # fsdp.py:233-241

    truth = copy.deepcopy(base)
    f_ref, f_ours = wrap(base), wrap(ours)
    mp32 = MixedPrecision(param_dtype=torch.float32, reduce_dtype=torch.float32, buffer_dtype=torch.float32)  # fsdp.py:226-230
    f_true = FSDP(truth, auto_wrap_policy=truth.get_fsdp_wrapping_policy(), mixed_precision=mp32, sharding_strategy=ShardingStrategy.FULL_SHARD,
                  device_id=dev if gpu else torch.device("cpu"), limit_all_gathers=True, use_orig_params=True)
    units = lambda f: sorted(n for n, m in f.named_modules() if isinstance(m, FSDP) and n)  # noqa: E731
    u_ref, u_ours = units(f_ref), units(f_ours)
    if args.mode == "per_unit":
        # the same FSDP units as the reference: every resampler, its inner projector, and the head
        assert u_ours == u_ref and any("projectors.0" in n for n in u_ours), (u_ref, u_ours)
    else:
        assert not any("projectors" in n for n in u_ours), u_ours  # projectors + adapter share the root unit

    opt_ref = torch.optim.SGD(f_ref.parameters(), lr=0.05)
    opt_ours = torch.optim.SGD(f_ours.parameters(), lr=0.05)
    opt_true = torch.optim.SGD(f_true.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(1234 + rank)  # every rank trains on its own micro-batch (base_strategy.py:153-161)
    B = 2
    info = {"rank": rank, "mode": args.mode, "mlp_type": args.mlp_type, "resampler": args.resampler, "units_ours": u_ours, "losses": [], "fused_fn": None}

    def rel(a, b):
        return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12))

    def full_grads(f):
        with FSDP.summon_full_params(f, with_grads=True):
            return {n: p.grad.detach().float().clone() for n, p in f.named_parameters() if p.grad is not None}

    def full_params(f):
        with FSDP.summon_full_params(f):
            return {n: p.detach().float().clone() for n, p in f.named_parameters()}

    p0 = full_params(f_ref)
    autocast = torch.autocast("cuda", dtype=torch.bfloat16) if gpu else torch.autocast("cpu", enabled=False)
    for step in range(args.steps):
        feats = [(torch.randn((B, frames, n, c), generator=g) + mu).to(torch.bfloat16).to(dev) for n, c, mu in zip(patches, dims, (0.0, 0.5, -0.5))]
        tgt = torch.randn((B, T, 32), generator=g).to(dev)
        wt = torch.randn((B, len(dims)), generator=g).to(dev)
        losses = []
        for f, opt in ((f_ref, opt_ref), (f_ours, opt_ours), (f_true, opt_true)):
            opt.zero_grad(set_to_none=True)
            if f is f_true:  # yardstick: the reference in fp32, no autocast
                y, w = f([x.float() for x in feats])
            else:
                with autocast:
                    y, w = f(feats)
            if f is f_ours and step == 0:
                fn, seen, stack = None, set(), [y.grad_fn]
                while stack:  # which autograd Function produced the prefix
                    n = stack.pop()
                    if n is None or n in seen:
                        continue
                    seen.add(n)
                    if type(n).__name__.startswith(("_FusedLinearFn", "_MixFn")):
                        fn = type(n).__name__
                        break
                    stack.extend(x[0] for x in n.next_functions)
                info["fused_fn"] = fn
            loss = (y.float() * tgt).mean() + (w.float() * wt).mean()
            loss.backward()
            losses.append(float(loss))
            if step == 0:
                (info.setdefault("grads", [])).append(full_grads(f))
            opt.step()
        info["losses"].append(losses)
    g_ref, g_ours, g_true = info.pop("grads")
    assert set(g_ref) == set(g_ours) == set(g_true), (sorted(g_ref), sorted(g_ours))

    def errors(ours_d, ref_d, true_d):
        out = {}
        for n, t in true_d.items():
            scale = float(t.abs().max())
            if scale == 0:  # v_proj / out_proj never reach an output: zero gradient in the reference too
                assert float(ours_d[n].abs().max()) == 0, f"{n}: the reference gives a zero gradient / update here"
                continue
            out[n] = (float((ours_d[n] - t).abs().max()) / scale, float((ref_d[n] - t).abs().max()) / scale)
        return out

    grad_err = errors(g_ours, g_ref, g_true)
    p_ref, p_ours, p_true = full_params(f_ref), full_params(f_ours), full_params(f_true)
    delta_err = errors({n: p_ours[n] - p0[n] for n in p0}, {n: p_ref[n] - p0[n] for n in p0}, {n: p_true[n] - p0[n] for n in p0})
    bad = {f"{kind}:{n}": (round(eo, 4), round(er, 4)) for kind, d in (("grad", grad_err), ("delta", delta_err)) for n, (eo, er) in d.items()
           if eo > 2.0 * er + 1.5e-2}
    f_ref.eval(), f_ours.eval()
    with torch.no_grad(), autocast:
        e_ref, w_ref = f_ref(feats)
        e_ours, w_ours = f_ours(feats)
    info["grad_err_ours_vs_ref_bf16"] = {n: (round(a, 4), round(b, 4)) for n, (a, b) in grad_err.items()}
    info.update(loss_err=max(abs(a - b) / max(abs(a), 1e-6) for a, b, _ in info["losses"]), bad=bad, eval_err=rel(e_ours, e_ref),
                eval_w_err=rel(w_ours, w_ref), n_grads=len(grad_err), n_deltas=len(delta_err))
    ok = info["loss_err"] < 3e-2 and not bad and info["eval_err"] < 3e-2 and info["eval_w_err"] < 3e-2 and len(grad_err) >= 2 * len(dims) + 4
    want_fn = "_FusedLinearFn" if (args.mode == "root_unit" and args.mlp_type == "linear" and args.resampler == "3davg") else "_MixFn"
    ok = ok and info["fused_fn"] is not None and info["fused_fn"].startswith(want_fn)
    info["ok"] = bool(ok)
    if gpu:
        info["native_lib"] = NU.ops._lib.LIB_PATH
    print("FSDP_WORKER " + json.dumps(info), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
