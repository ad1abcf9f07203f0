"""Why is the CTA pair slow on MN-major operands?  The MN-capable instantiation on K-major operands (its code without the layout), and the
four majorness combinations, single CTA vs pair, at the dW shape (M = 4096, N = 1024, K = 65536) and the Z shape (M = 65536, N = 1024, K = 4096)."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
from merv_b200 import ops
dev = "cuda:0"
g0 = torch.Generator(device=dev).manual_seed(1)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
M = 65536
g = torch.randn(M, 4096, generator=g0, device=dev).to(torch.bfloat16)
W = (torch.randn(4096, 1024, generator=g0, device=dev) / 64).to(torch.bfloat16)
P = torch.randn(M, 1024, generator=g0, device=dev).to(torch.bfloat16)
Wt, gT, Pt = W.t().contiguous(), g.t().contiguous(), P.t().contiguous()
fl = 2 * M * 4096 * 1024
rep = {}
for grp in ("1", "2"):
    os.environ["MERV_GEMM_CTA_GROUP"] = grp
    r = {"Z_kmajor": timeit(lambda: ops.gemm_ex(g, Wt)), "dW_kmajor": timeit(lambda: ops.gemm_ex(gT, Pt))}
    rep[f"grp{grp}_kmajor"] = {k: (round(v, 4), round(fl / v / 1e9)) for k, v in r.items()}
    print(f"cta_group={grp} K-major: " + "  ".join(f"{k} {v:.3f}ms/{fl / v / 1e9:.0f}TF" for k, v in r.items()), flush=True)
    r = {"Z_w_t": timeit(lambda: ops.gemm_ex(g, W, w_t=True)), "dW_a_t_w_t": timeit(lambda: ops.gemm_ex(g, P, a_t=True, w_t=True)),
         "dW_a_t_only": timeit(lambda: ops.gemm_ex(g, Pt, a_t=True)), "dW_w_t_only": timeit(lambda: ops.gemm_ex(gT, P, w_t=True))}
    rep[f"grp{grp}_mn"] = {k: (round(v, 4), round(fl / v / 1e9)) for k, v in r.items()}
    print(f"cta_group={grp} MN-major: " + "  ".join(f"{k} {v:.3f}ms/{fl / v / 1e9:.0f}TF" for k, v in r.items()), flush=True)
# the per-video weight gradient (merv_wgrad_video) at 64 and 16 videos, both encoder widths, both CTA groups — with a value check between them
for videos in (64, 16):
    gv, sc = g[:videos * 1024], torch.rand(videos, generator=g0, device=dev) + 0.1
    for C in (1024, 768):
        Wc, Pc = W[:, :C].contiguous(), P[:videos * 1024, :C].contiguous()
        res = {}
        for grp in ("1", "2"):
            os.environ["MERV_GEMM_CTA_GROUP"] = grp
            dW, part = ops.wgrad_video(gv, Pc, sc, Wc, videos)
            res[grp] = (dW.float(), part.double().sum(1))
            t = timeit(lambda: ops.wgrad_video(gv, Pc, sc, Wc, videos))
            flv = 2 * videos * 1024 * 4096 * C
            rep[f"wgrad_video_B{videos}_C{C}_grp{grp}"] = (round(t, 4), round(flv / t / 1e9))
            print(f"wgrad_video videos={videos} C={C} cta_group={grp}: {t:.3f} ms / {flv / t / 1e9:.0f} TF", flush=True)
        ddw = float((res["1"][0] - res["2"][0]).abs().max() / res["1"][0].abs().max())
        ddot = float((res["1"][1] - res["2"][1]).abs().max() / res["1"][1].abs().max())
        print(f"   single vs pair: dW rel diff {ddw:.2e}, dot rel diff {ddot:.2e}", flush=True)
        rep[f"wgrad_video_B{videos}_C{C}_single_vs_pair"] = (ddw, ddot)
os.environ.pop("MERV_GEMM_CTA_GROUP", None)
# split over the videos (work items of a range of videos, fp32 partial tiles folded afterwards): SM utilisation of 96 - 128 tiles on 148 SMs
for videos in (64, 16):
    gv, sc = g[:videos * 1024], torch.rand(videos, generator=g0, device=dev) + 0.1
    for C in (1024, 768):
        Wc, Pc = W[:, :C].contiguous(), P[:videos * 1024, :C].contiguous()
        flv = 2 * videos * 1024 * 4096 * C
        for split in ("1", "2", "3", "4", "6", "8", None):
            if split is None:
                os.environ.pop("MERV_WGRAD_SPLIT", None)
            else:
                os.environ["MERV_WGRAD_SPLIT"] = split
            t = timeit(lambda: ops.wgrad_video(gv, Pc, sc, Wc, videos))
            rep[f"wgrad_video_B{videos}_C{C}_split{split}"] = (round(t, 4), round(flv / t / 1e9))
            print(f"wgrad_video videos={videos} C={C} split={split}: {t:.3f} ms / {flv / t / 1e9:.0f} TF", flush=True)
os.environ.pop("MERV_WGRAD_SPLIT", None)
json.dump(rep, open(os.path.join(REPO, "gpurun_out", "mn_pair_lab.json"), "w"), indent=1)
