"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous batch shards, no data-path collective,
optional all-gather of the prefixes in rank-major == batch order, ragged and empty shards."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _StubFusion(torch.nn.Module):
    """Per-video deterministic stand-in for MervFusion (the CUDA library cannot run on the CPU test box)."""

    class _FF:
        token_length, llm_dim = 3, 5

    feature_fusion = _FF()

    def forward(self, feats):
        b = feats[0].shape[0]
        key = feats[0].reshape(b, -1).sum(1)  # identifies the video
        prefix = key.view(b, 1, 1) + torch.arange(15, dtype=feats[0].dtype).view(1, 3, 5)
        weights = torch.stack([key, -key], 1)
        return prefix, weights


def _worker(rank, world, port, batch, q):
    sys.path.insert(0, REPO)
    from merv_b200 import parallel as P

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feats = [torch.arange(batch * 4, dtype=torch.float32).view(batch, 4) * 1.0, torch.ones(batch, 2)]
        module = _StubFusion()
        full_prefix, full_w = module(feats)
        lo, hi = P.shard_bounds(batch, rank, world)
        local_prefix, local_w = P.fusion_forward_sharded(module, feats, gather=False)
        ok_local = torch.equal(local_prefix, full_prefix[lo:hi]) and torch.equal(local_w, full_w[lo:hi])
        g_prefix, g_w = P.fusion_forward_sharded(module, feats, gather=True)
        ok_gather = torch.equal(g_prefix, full_prefix) and torch.equal(g_w, full_w)
        q.put((rank, lo, hi, ok_local, ok_gather))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5, 1])
def test_sharded_equals_unsharded_world2(batch):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, lo, hi, ok_local, ok_gather in results:
        assert ok_local, f"rank {rank}: local shard differs from the same videos computed unsharded"
        assert ok_gather, f"rank {rank}: gathered prefixes are not in batch order"
        covered += list(range(lo, hi))
    assert covered == list(range(batch)), "shards must tile the batch contiguously, rank-major"


def test_shard_bounds_tile_the_batch():
    from merv_b200.parallel import shard_bounds

    for batch in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [shard_bounds(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
