"""e2e (host buffers in -> host buffers out, merv_b200.pipeline.HostPipeline) against the chunk size: videos per chunk 4 / 8 / 16 / 32, next to the bare
H2D copy of the same bytes.  One GPU."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import merv_b200 as M
from merv_b200.pipeline import HostPipeline
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
m = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024)
with torch.no_grad():
    m.feature_fusion.Q.mul_(64.0)
m = m.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
B = 64
shapes = [(B, 16, 256, 1024), (B, 16, 256, 1024), (B, 16, 196, 768), (B, 16, 196, 768)]
host_in = [torch.randn(s).to(torch.bfloat16).pin_memory() for s in shapes]
host_out = torch.empty((B, 1024, 4096), dtype=torch.bfloat16).pin_memory()
host_w = torch.empty((B, 4), dtype=torch.bfloat16).pin_memory()
dev_in = [torch.empty(s, dtype=torch.bfloat16, device=dev) for s in shapes]
rep = {}
def bare():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for h, d in zip(host_in, dev_in):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
rep["h2d_only_ms"] = min(bare() for _ in range(4))
print("bare H2D", rep["h2d_only_ms"], flush=True)
for rnd in range(2):
    for chunk in (4, 8, 16, 32):
        pipe = HostPipeline(m, chunk_videos=chunk, device=dev)
        for _ in range(2):
            pipe(host_in, host_out, host_w)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            pipe(host_in, host_out, host_w)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 8 * 1e3
        rep.setdefault(f"chunk{chunk}", []).append(ms)
        print(f"chunk {chunk}: {ms:.2f} ms per 64 videos = {B / ms * 1e3:.0f} videos/s", flush=True)
json.dump(rep, open(os.path.join(REPO, "gpurun_out", "e2e_lab.json"), "w"), indent=1)
