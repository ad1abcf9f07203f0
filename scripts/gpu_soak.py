"""Soak test: thousands of launches of the fused and module-by-module paths at varying batch sizes; every result must be
bit-identical to the first one for the same input (catches rare races / barrier-protocol bugs that single runs miss)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import merv_b200 as M

dev = "cuda:0"
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
torch.manual_seed(0)
mods = {}
for mlp in ("linear", "gelu-mlp"):
    for fused in (True, False):
        m = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, mlp, seed=1024, fused=fused)
        with torch.no_grad():
            m.feature_fusion.Q.mul_(64.0)
        mods[(mlp, fused)] = m.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
g = torch.Generator(device=dev).manual_seed(11)
Bmax = 48
feats = [(torch.randn((Bmax, 16, n, c), generator=g, device=dev) + mu).to(torch.bfloat16) for n, c, mu in zip([256, 256, 196, 196], [1024, 1024, 768, 768], [0, .5, -.5, .25])]
ref = {}
t0, it, bad = time.time(), 0, 0
sizes = [1, 2, 3, 5, 8, 13, 16, 24, 32, 48]
with torch.inference_mode():
    while time.time() - t0 < seconds:
        B = sizes[it % len(sizes)]
        for key, m in mods.items():
            if key[0] == "gelu-mlp" and B > 16 and it % 4:
                continue  # keep the expensive configuration rarer
            out, w = m([f[:B] for f in feats])
            k = key + (B,)
            if k not in ref:
                ref[k] = (out.clone(), w.clone())
            elif not (torch.equal(out, ref[k][0]) and torch.equal(w, ref[k][1])):
                bad += 1
                print("MISMATCH", k, "iteration", it, flush=True)
        it += 1
        if it % 50 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    # cross-check: video b alone == video b in a batch, for the last batch size of every module
    for key, m in mods.items():
        full, _ = m([f[:8] for f in feats])
        one, _ = m([f[3:4] for f in feats])
        if not torch.equal(full[3], one[0]):
            bad += 1
            print("BATCH-INVARIANCE MISMATCH", key, flush=True)
print(f"soak: {it} rounds x {len(mods)} modules in {time.time() - t0:.0f} s, mismatches: {bad}", flush=True)
sys.exit(1 if bad else 0)
