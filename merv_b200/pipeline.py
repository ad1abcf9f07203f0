"""Host-buffer entry point of the fusion path: features in (pinned) host memory -> fused prefix in host memory.

This is the call a consumer without device-resident features makes, and what bench.py times as ``e2e``: every
step includes the host->device copy of the inputs and the device->host copy of the result.  Videos are
independent, so the batch is streamed in chunks: copy-in of chunk i+1, compute of chunk i and copy-out of chunk
i-1 overlap on three CUDA streams with double-buffered device staging.  The chunk sizes ramp up at the start and
down at the end (2, 2, 4, 8, ..., 8, 4, 2, 2 for chunk_videos = 8): the bus sits idle only while the first chunk
is in flight alone (fill) and after the last one landed (drain), so those are the small ones.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def chunk_schedule(batch: int, chunk: int) -> List[Tuple[int, int]]:
    """[lo, hi) video ranges covering `batch`: ramp-up, full-size chunks, ramp-down."""
    ramp = [max(1, chunk // 4), max(1, chunk // 4), max(1, chunk // 2)]
    sizes_head: List[int] = []
    sizes_tail: List[int] = []
    left = batch
    for r in ramp:
        if left >= 2 * chunk + 2 * r:  # keep at least two full chunks in the middle
            sizes_head.append(r)
            sizes_tail.append(r)
            left -= 2 * r
    middle = [chunk] * (left // chunk) + ([left % chunk] if left % chunk else [])
    out, lo = [], 0
    for n in sizes_head + middle + sizes_tail[::-1]:
        out.append((lo, lo + n))
        lo += n
    assert lo == batch
    return out


class HostPipeline:
    def __init__(self, module: torch.nn.Module, chunk_videos: int = 8, device: Optional[torch.device] = None) -> None:
        self.module = module
        self.chunk = chunk_videos
        self.device = device or next(module.parameters()).device
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self._staging = None

    def _staging_for(self, feats: Sequence[torch.Tensor], dtype: torch.dtype):
        key = tuple((tuple(f.shape[1:]), dtype) for f in feats)
        if self._staging is None or self._staging[0] != key:
            bufs = [[torch.empty((self.chunk, *f.shape[1:]), dtype=dtype, device=self.device) for f in feats] for _ in range(2)]
            self._staging = (key, bufs)
        return self._staging[1]

    @torch.inference_mode()
    def __call__(self, feats_host: Sequence[torch.Tensor], out_host: Optional[torch.Tensor] = None,
                 weights_host: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """feats_host[e]: [B, F, N, C] CPU tensors (pinned for full copy bandwidth).  Returns (prefix, weights) on the host."""
        B = feats_host[0].shape[0]
        dtype = feats_host[0].dtype
        ff = self.module.feature_fusion
        T, K, E = ff.token_length, ff.llm_dim, len(feats_host)
        if out_host is None:
            out_host = torch.empty((B, T, K), dtype=dtype).pin_memory()
        if weights_host is None:
            weights_host = torch.empty((B, E), dtype=dtype).pin_memory()
        bufs = self._staging_for(feats_host, dtype)
        compute_done: List[Optional[torch.cuda.Event]] = [None, None]
        start = torch.cuda.Event()
        start.record(torch.cuda.current_stream(self.device))
        for s in (self.s_in, self.s_compute, self.s_out):
            s.wait_event(start)
        for i, (lo, hi) in enumerate(chunk_schedule(B, self.chunk)):
            n, slot = hi - lo, i % 2
            with torch.cuda.stream(self.s_in):
                if compute_done[slot] is not None:
                    self.s_in.wait_event(compute_done[slot])  # staging slot is free again
                for buf, f in zip(bufs[slot], feats_host):
                    buf[:n].copy_(f[lo:hi], non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self.s_in)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(copied)
                prefix, weights = self.module([b[:n] for b in bufs[slot]])
                done = torch.cuda.Event()
                done.record(self.s_compute)
                compute_done[slot] = done
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                prefix.record_stream(self.s_out)
                weights.record_stream(self.s_out)
                out_host[lo:hi].copy_(prefix, non_blocking=True)
                weights_host[lo:hi].copy_(weights, non_blocking=True)
        self.s_out.synchronize()
        return out_host, weights_host
