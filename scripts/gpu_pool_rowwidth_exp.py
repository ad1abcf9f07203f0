"""Does the pool kernel gain from longer contiguous DRAM segments?  The same bytes as [B, 16, 256, C] with C from 1024 (128-byte pieces of 2 KB rows)
down to 64 (fully contiguous slabs).  Measured on a B200: 6186 -> 6294 GB/s, i.e. no."""
import sys, torch
sys.path.insert(0, '/root/repo')
from merv_b200 import ops
dev='cuda:0'
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for C,B in ((1024,64),(512,128),(256,256),(128,512),(64,1024)):
    xs=[torch.randn((B,16,256,C),device=dev).to(torch.bfloat16) for _ in range(3)]
    i=[0]
    def f():
        ops.pool3d([xs[i[0]%3]],[16],8); i[0]+=1
    ms=t(f)
    by=B*16*256*C*2*1.25
    print(f"C={C} B={B}: {ms:.4f} ms  {by/ms/1e6:.0f} GB/s", flush=True)
