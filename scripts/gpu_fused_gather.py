"""Multi-GPU check + timing of the fused GEMM + all-gather (torchrun, one process per GPU).

    torchrun --nproc-per-node N --master-addr 127.0.0.1 scripts/gpu_fused_gather.py [batch_per_gpu]
"""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import torch.distributed as dist
import merv_b200 as M
from merv_b200.parallel import SymmetricPrefixBuffer, all_gather_prefix

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
m = M.MervFusion.build([1024, 1024, 768, 768], 4096, [16] * 4, 64, "linear", seed=1024)
with torch.no_grad():
    m.feature_fusion.Q.mul_(64.0)
m = m.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
g = torch.Generator(device=dev).manual_seed(100 + rank)
feats = [(torch.randn((B, 16, n, c), generator=g, device=dev) + mu).to(torch.bfloat16) for n, c, mu in zip([256, 256, 196, 196], [1024, 1024, 768, 768], [0, .5, -.5, .25])]
buf = SymmetricPrefixBuffer(B, 1024, 4096, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


with torch.inference_mode():
    local_out, local_w = m(feats)
    want = all_gather_prefix(local_out, B * world)          # compute, then NCCL all-gather
    got, w = m(feats, gather=buf)                            # fused: the GEMM epilogue stores to every rank's buffer
    torch.cuda.synchronize()
    ok = bool(torch.equal(got, want)) and bool(torch.equal(w, local_w))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    t_compute = timeit(lambda: m(feats))
    t_nccl = timeit(lambda: all_gather_prefix(m(feats)[0], B * world))
    t_fused = timeit(lambda: m(feats, gather=buf))
if rank == 0:
    res = {"world": world, "batch_per_gpu": B, "gathered_bit_identical_on_all_ranks": bool(flag.item()), "compute_only_ms": t_compute,
           "compute_then_nccl_allgather_ms": t_nccl, "fused_gemm_allgather_ms": t_fused, "wide_out_env": os.environ.get("MERV_GEMM_WIDE_OUT"), "gathered_bytes_per_rank": (world - 1) * B * 1024 * 4096 * 2}
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(REPO, "gpurun_out", f"fused_gather_{world}gpu{os.environ.get('MERV_TAG', '')}.json"), "w"), indent=1)
dist.destroy_process_group()
