"""Data-parallel sharding of the fusion path: one process per GPU, videos are independent units.

The reference has no parallelism on this path beyond data parallel: every video's
fusion depends only on its own features and the replicated weights.  So the batch is split into contiguous
blocks, rank r owning videos [r*ceil(B/G), ...), there is NO collective on the compute path, and the only
optional exchange is one all-gather of the fused prefixes when a consumer wants them on every rank
(NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests).
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of videos owned by `rank`: rank-major order == batch order after the gather."""
    per = (batch + world - 1) // world
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def shard_features(features: Sequence[torch.Tensor], rank: int, world: int) -> List[torch.Tensor]:
    lo, hi = shard_bounds(features[0].shape[0], rank, world)
    return [f[lo:hi] for f in features]


def all_gather_prefix(local: torch.Tensor, batch: int, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Gather per-rank prefixes [b_r, T, K] into [batch, T, K] on every rank (rank-major == batch order).

    Ranks may own unequal (even empty) blocks when `batch` is not a multiple of the world size: blocks are
    padded to ceil(batch/world) videos for the collective and the padding is dropped afterwards.
    """
    world = dist.get_world_size(group)
    per = (batch + world - 1) // world
    T, K = local.shape[1], local.shape[2]
    if local.shape[0] != per:
        padded = local.new_zeros((per, T, K))
        padded[: local.shape[0]] = local
    else:
        padded = local.contiguous()
    out = local.new_empty((world * per, T, K))
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, padded, group=group)
    else:
        dist.all_gather(list(out.view(world, per, T, K).unbind(0)), padded, group=group)
    return out[:batch]


def fusion_forward_sharded(module, features: Sequence[torch.Tensor], gather: bool = False,
                           group: Optional[dist.ProcessGroup] = None):
    """Run `module` (a MervFusion) on this rank's block of the GLOBAL batch held in `features`.

    Returns (prefix, weights) for the local block, or for the whole batch on every rank when gather=True.
    """
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    batch = features[0].shape[0]
    local = shard_features(features, rank, world)
    if local[0].shape[0] > 0:
        prefix, weights = module(local)
    else:  # more ranks than videos
        T, K = module.feature_fusion.token_length, module.feature_fusion.llm_dim
        prefix = features[0].new_zeros((0, T, K))
        weights = features[0].new_zeros((0, len(features)))
    if not gather:
        return prefix, weights
    w3 = all_gather_prefix(weights.unsqueeze(-1), batch, group).squeeze(-1)
    return all_gather_prefix(prefix, batch, group), w3


class SymmetricPrefixBuffer:
    """The whole batch's fused prefixes `[world * batch_per_rank, T, K]` (bf16) in NVLink peer-mapped symmetric memory.

    Passed as `gather=` to a linked `MervFusion` / adapter call, the fused projector+mix GEMM stores every finished output
    tile into this rank's block of EVERY rank's buffer with TMA stores to peer addresses (`merv_fused_linear_mix_gather`):
    the all-gather rides on the GEMM's epilogue, tile by tile, instead of running as a separate NCCL collective afterwards.
    Memory comes from `torch.distributed._symmetric_memory` (plumbing: allocation, rendezvous, barriers).
    """

    def __init__(self, batch_per_rank: int, tokens: int, width: int, group: Optional[dist.ProcessGroup] = None,
                 device: Optional[torch.device] = None) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert self.world <= 8, "the fused GEMM stores to at most 7 peers"
        self.batch_per_rank = batch_per_rank
        device = device or torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty((self.world * batch_per_rank, tokens, width), dtype=torch.bfloat16, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.block_bytes = batch_per_rank * tokens * width * 2
        # NVSwitch multicast mapping of the buffers (NVLS): one store is replicated by the switch into every rank's buffer.
        # 0 / absent where the fabric or driver has no multicast support.
        import os

        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        # Measured on B200 x 2 / x 8 (64 videos per rank, compute alone 1.75-1.85 ms): unicast 1.93 / 6.33 ms, multicast 2.15 / 6.68 ms.
        # The all-gather is bound by every GPU's NVLink INGRESS (world - 1 blocks; NCCL's own all-gather of the same bytes takes
        # 5.9 ms at 8 GPUs), not by its egress: unicast loads both directions of the full-duplex links equally, while a multicast
        # store also loops the sender's own block back through the switch (world instead of world - 1 blocks of ingress) and pays
        # the per-thread multimem.st issue.  So unicast TMA stores are the default; MERV_GATHER_TRANSPORT=multicast selects the
        # switch-replicated path (kept tested: it is the right choice where egress, not ingress, is the scarce direction).
        choice = os.environ.get("MERV_GATHER_TRANSPORT", "")
        if choice != "multicast":
            mc = 0
        self.multicast_ptr = mc
        self.transport = ("multimem.st to the NVSwitch multicast mapping (one egress write per byte)" if mc
                          else f"unicast TMA stores to {self.world - 1} peer-mapped buffers")

    def multicast_block_ptr(self) -> int:
        """Multicast address of THIS rank's block (0 when multicast is unavailable)."""
        return self.multicast_ptr + self.rank * self.block_bytes if self.multicast_ptr else 0

    def local_block(self) -> torch.Tensor:
        lo = self.rank * self.batch_per_rank
        return self.buf[lo:lo + self.batch_per_rank]

    def peer_block_ptrs(self) -> List[int]:
        """Addresses of THIS rank's block inside every other rank's buffer (peer-mapped device pointers), in ring order
        starting at the next rank so that at any moment the ranks target different peers."""
        ptrs = self.handle.buffer_ptrs
        return [int(ptrs[(self.rank + k) % self.world]) + self.rank * self.block_bytes for k in range(1, self.world)]

    def barrier(self) -> None:
        """Stream-ordered barrier across the ranks (device-side signals; no host synchronisation)."""
        self.handle.barrier()
