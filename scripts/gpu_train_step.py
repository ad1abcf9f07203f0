"""Training-step timing (forward + backward) at merv-full shapes, bf16: module-by-module autograd path vs the fused training path
(fused forward + _FusedLinearFn backward) vs the reference's op sequence in torch eager, with per-entry-point device times.

    python scripts/gpu_train_step.py [B ...]      # default 16 (the reference's per-device batch, conf/models.py:125) and 64
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch

import merv_b200 as M
from merv_b200 import ops
from oracle import torch_port

dev = "cuda:0"
DIMS, PATCHES = [1024, 1024, 768, 768], [256, 256, 196, 196]


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


report = {}
ONLY = os.environ.get("TRAIN_STEP_ONLY")  # "fused_training" | "module_by_module": restrict the run (ncu launch lists)
for Bt in [int(a) for a in sys.argv[1:]] or [16, 64]:
    g = torch.Generator(device=dev).manual_seed(5)
    feats = [torch.randn((Bt, 16, n, c), generator=g, device=dev).to(torch.bfloat16) for n, c in zip(PATCHES, DIMS)]
    G = torch.randn((Bt, 1024, 4096), generator=g, device=dev).to(torch.bfloat16)
    res = {}
    for name, fused_training in (("module_by_module", False), ("fused_training", True)):
        if ONLY and name != ONLY:
            continue
        mod = M.MervFusion.build(DIMS, 4096, [16] * 4, 64, "linear", seed=1024).to(device=dev, dtype=torch.bfloat16).train()
        mod.feature_fusion.fused_training = fused_training

        params = [p for p in mod.parameters()]

        def step():
            out, w = mod(feats)
            out.backward(G)
            mod.zero_grad(set_to_none=True)
            with torch.no_grad():  # what an optimizer step does to the caches: every parameter gets a new version, so the cached
                torch._foreach_add_(params, 0.0)  # compute-dtype copies, u and the per-projector score vectors are rebuilt next step

        res[name] = timeit(step)
        with ops.KernelTimer(timing=True) as kt:
            step()
        per = {k: round(sum(v), 4) for k, v in kt.durations_ms().items()}
        res[name + "_device_ms_by_entry_point"] = dict(sorted(per.items(), key=lambda kv: -kv[1]))
        res[name + "_peak_mem_GB"] = None
        torch.cuda.reset_peak_memory_stats()
        base = torch.cuda.memory_allocated()
        step()
        res[name + "_peak_mem_GB"] = round((torch.cuda.max_memory_allocated() - base) / 1e9, 3)
        del mod
    if Bt <= 16 and not ONLY:
        ref = M.MervFusion.build(DIMS, 4096, [16] * 4, 64, "linear", seed=1024).to(device=dev, dtype=torch.bfloat16)
        pp = [{k: v.detach().clone().requires_grad_(True) for k, v in p.projector.state_dict().items()} for p in ref.projectors]
        fp = {k: v.detach().clone().requires_grad_(True) for k, v in ref.feature_fusion.state_dict().items()}

        def eager_step():
            out, w = torch_port.fusion_forward_autograd(feats, pp, fp, [16] * 4, 8, "linear", 1024)
            out.backward(G)
            for d in pp + [fp]:
                for t in d.values():
                    t.grad = None

        res["torch_eager_reference_ops"] = timeit(eager_step)
    report[f"B{Bt}"] = res
    print(f"B={Bt}: " + json.dumps(res), flush=True)
os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(REPO, "gpurun_out", "train_step.json"), "w"), indent=1)
