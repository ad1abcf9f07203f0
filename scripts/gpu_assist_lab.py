"""Pool assist (merv_b200/csrc/pool_assist.cuh) on a B200: bit-identity against the three-launch path and a timing sweep over the head size.

    python scripts/gpu_assist_lab.py [--out gpurun_out/assist_lab.json] [--heads 4,8,12,16,24,32] [--batches 16,24,64]

Every run is bounded by a watchdog thread (a protocol bug traps the kernel; a hang would otherwise cost the whole gpurun call)."""
import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

import merv_b200 as M  # noqa: E402

DEV = "cuda:0"


def module():
    m = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024, fused=True)
    with torch.no_grad():
        m.feature_fusion.Q.mul_(64.0)
    return m.to(device=DEV, dtype=torch.bfloat16).eval().requires_grad_(False)


def features(B, seed=7):
    g = torch.Generator(device=DEV).manual_seed(seed)
    shapes = [(B, 16, 256, 1024), (B, 16, 256, 1024), (B, 16, 196, 768), (B, 16, 196, 768)]
    mus = [0.0, 0.5, -0.5, 0.25]
    return [(torch.randn(s, generator=g, device=DEV) + mu).to(torch.bfloat16) for s, mu in zip(shapes, mus)]


def run(m, feats, assist, head=None):
    os.environ["MERV_POOL_ASSIST"] = "1" if assist else "0"
    if head is None:
        os.environ.pop("MERV_ASSIST_HEAD", None)
    else:
        os.environ["MERV_ASSIST_HEAD"] = str(head)
    with torch.inference_mode():
        out, w = m(feats)
    torch.cuda.synchronize()
    return out, w


def timed(m, sets, assist, head, steps=20, warmup=5):
    os.environ["MERV_POOL_ASSIST"] = "1" if assist else "0"
    if head is None:
        os.environ.pop("MERV_ASSIST_HEAD", None)
    else:
        os.environ["MERV_ASSIST_HEAD"] = str(head)
    with torch.inference_mode():
        for i in range(warmup):
            m(sets[i % len(sets)])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            m(sets[i % len(sets)])
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/assist_lab.json")
    ap.add_argument("--heads", default="8,16,24,32,40")
    ap.add_argument("--batches", default="24,64")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--dbg", action="store_true", help="profile builds: also time the debug variants")
    args = ap.parse_args()
    threading.Thread(target=lambda: (time.sleep(420), print("watchdog: giving up", flush=True), os._exit(3)), daemon=True).start()
    rec = {"identity": [], "timing": []}
    m = module()
    for B in [int(b) for b in args.batches.split(",")]:
        feats = features(B)
        ref_out, ref_w = run(m, feats, False)
        for head in (None, 1, B // 2, B - 1):
            out, w = run(m, feats, True, head)
            out2, w2 = run(m, feats, True, head)  # cached plan, flags reset by the scores kernel
            ok = bool(torch.equal(out, ref_out) and torch.equal(w, ref_w) and torch.equal(out2, ref_out) and torch.equal(w2, ref_w))
            bad = int((out != ref_out).any(-1).any(-1).sum()) if not ok else 0
            rec["identity"].append({"B": B, "head": head, "bit_identical": ok, "videos_differing": bad,
                                    "max_abs_diff": float((out.float() - ref_out.float()).abs().max())})
            print(rec["identity"][-1], flush=True)
    B = 64
    sets = [features(B, seed=7 + i) for i in range(3)]
    # The GPU drifts (clocks sink under the power cap as it warms up) by more than the effects looked for: every variant is timed
    # right next to the three-launch path, twice, and reported as the ratio of adjacent runs.
    def ab(tag, head, env=None):
        ratios, ts, bs = [], [], []
        for _ in range(args.repeats):
            for k in list(os.environ):
                if k in ("MERV_ASSIST_DBG", "MERV_ASSIST_GROUP"):
                    os.environ.pop(k)
            b = timed(m, sets, False, None, args.steps)
            os.environ.update(env or {})
            t = timed(m, sets, True, head, args.steps)
            ratios.append(t / b); ts.append(t); bs.append(b)
        for k in (env or {}):
            os.environ.pop(k, None)
        rec["timing"].append({"B": B, "tag": tag, "head": head, "env": env, "assist_ms": ts, "base_ms": bs, "ratio": ratios})
        print(rec["timing"][-1], flush=True)
    for head in [int(h) for h in args.heads.split(",")]:
        ab("head sweep", head)
    ab("one video pooled in the GEMM", B - 1)
    if args.dbg:
        ab("no flag polling (invalid results)", B - 1, {"MERV_ASSIST_DBG": "8"})
        ab("no pooling, no polling (invalid results)", B - 1, {"MERV_ASSIST_DBG": "24"})
        ab("group 16", 32, {"MERV_ASSIST_GROUP": "16"})
    # per-warp cycle counters of the pooling warps (profile build only)
    os.environ["MERV_ASSIST_PROFILE"] = "1"
    for head, dbg in ((32, 0),):
        os.environ["MERV_ASSIST_DBG"] = str(dbg)
        run(m, sets[0], True, head)
        plan = m.feature_fusion.__dict__.get("_last_plan")
        fast = m.__dict__.get("_fast", {})
        plans = [v[1] for v in fast.values()] + ([plan] if plan is not None else [])
        for pl in plans:
            if pl.B != B:
                continue
            prof = pl.sync_ws[4 + 2 * B:].view(-1, 8).cpu()
            prof = prof[prof[:, 0] > 0].double()
            if prof.numel() == 0:
                continue
            cyc = prof[:, 1:6] * 16
            rec.setdefault("profile", []).append({
                "head": head, "dbg": dbg, "warps": int(prof.shape[0]), "items_per_warp_mean": float(prof[:, 0].mean()),
                "items_per_warp_min": float(prof[:, 0].min()), "items_per_warp_max": float(prof[:, 0].max()),
                "cycles_per_item": {k: float((cyc[:, i] / prof[:, 0]).mean()) for i, k in enumerate(["wait_tma", "pool", "complete", "finish_video", "total"])},
                "total_cycles_mean": float(cyc[:, 4].mean())})
            print(rec["profile"][-1], flush=True)
            break
    os.environ.pop("MERV_ASSIST_PROFILE", None)
    os.environ.pop("MERV_ASSIST_DBG", None)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main()
