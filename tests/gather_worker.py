"""One rank of the fused GEMM + all-gather test (launched by tests/test_multi_gpu.py through torch.distributed.run, one GPU per rank).

Every rank fuses its own shard of videos; the fused projector + mix GEMM stores each finished output box into every rank's
symmetric prefix buffer — through the NVSwitch multicast mapping (multimem.st) where the fabric offers one, through unicast TMA
stores to the peer-mapped buffers otherwise (and always under MERV_GATHER_TRANSPORT=unicast).  The gathered result must be
BIT-IDENTICAL, on every rank, to "compute locally, then NCCL all_gather_into_tensor", for a small shape (partial tiles) and for
merv-full shapes, and the same buffer must survive repeated use (no stale data between steps).
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import merv_b200 as M
    from merv_b200.parallel import SymmetricPrefixBuffer, all_gather_prefix

    report = {"rank": rank, "world": world, "cases": []}
    ok = True
    shapes = {
        "small": dict(dims=[128, 64], patches=[64, 49], frames=4, llm=256, tok=64, embed=96, B=3),
        "merv_full": dict(dims=[1024, 1024, 768, 768], patches=[256, 256, 196, 196], frames=16, llm=4096, tok=1024, embed=3072, B=2),
    }
    for name, c in shapes.items():
        S = int((c["tok"] // c["frames"]) ** 0.5)
        m = M.MervFusion.build(c["dims"], c["llm"], [c["frames"]] * len(c["dims"]), S * S, "linear", text_embedding_dim=c["embed"], seed=c["dims"][0])
        with torch.no_grad():
            m.feature_fusion.Q.mul_(32.0)
        m = m.to(device=dev, dtype=torch.bfloat16).eval().requires_grad_(False)
        for transport in ("multicast", "unicast"):  # forced either way: the automatic choice depends on the number of ranks
            os.environ["MERV_GATHER_TRANSPORT"] = transport
            buf = SymmetricPrefixBuffer(c["B"], c["tok"], c["llm"], device=dev)
            same = True
            with torch.inference_mode():
                for it in range(3):  # the buffer is reused: step i must not see step i-1's data
                    g = torch.Generator(device=dev).manual_seed(1000 * it + 17 * rank + 3)
                    feats = [(torch.randn((c["B"], c["frames"], n, d), generator=g, device=dev) + 0.3 * e).to(torch.bfloat16)
                             for e, (n, d) in enumerate(zip(c["patches"], c["dims"]))]
                    local_out, local_w = m(feats)
                    want = all_gather_prefix(local_out, c["B"] * world)
                    got, w = m(feats, gather=buf)
                    torch.cuda.synchronize()
                    same = same and bool(torch.equal(got, want)) and bool(torch.equal(w, local_w))
                    same = same and bool(torch.equal(got[rank * c["B"]:(rank + 1) * c["B"]], local_out))
            report["cases"].append({"shape": name, "transport": buf.transport, "multicast": bool(buf.multicast_ptr), "bit_identical": same})
            ok = ok and same
            del buf
    os.environ.pop("MERV_GATHER_TRANSPORT", None)
    report["ok"] = ok
    print("GATHER_WORKER " + json.dumps(report), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
