"""Host-side runtime around the backbone: batch sharding across ranks (the path's only multi-GPU split,
SURVEY.md §8e — images never interact, so there is no data-path collective) and a pinned-memory pipeline
that overlaps H2D copies, the forward and the D2H read-back of the feature maps on separate CUDA streams."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of `n_items` for `rank`: sizes differ by at most one, earlier ranks take
    the remainder (the layout DistributedSampler-style batch sharding produces for an in-order batch)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's images of a global batch [B, 3, H, W] (a view, no copy)."""
    b, e = shard_bounds(batch.shape[0], rank, world)
    return batch[b:e]


def gather_features(local_outs: Sequence[torch.Tensor], world: int, group=None) -> List[torch.Tensor]:
    """Reassemble per-stage feature maps of a sharded batch on every rank (test / evaluation helper; the
    inference path itself needs no collective).  Works with gloo (CPU tensors) and nccl."""
    import torch.distributed as dist
    outs = []
    for o in local_outs:
        sizes = [torch.zeros(1, dtype=torch.int64, device=o.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([o.shape[0]], dtype=torch.int64, device=o.device), group=group)
        mx = int(max(int(s) for s in sizes))
        pad = o if o.shape[0] == mx else torch.cat([o, o.new_zeros((mx - o.shape[0],) + o.shape[1:])], 0)
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad.contiguous(), group=group)
        outs.append(torch.cat([b[:int(s)] for b, s in zip(bufs, sizes)], 0))
    return outs


class HostPipeline:
    """model(images on the host) -> feature maps on the host, chunked and overlapped.

    Three streams: copy-in (H2D of chunk i+1), compute (forward of chunk i), copy-out.  Every stage's feature map
    is copied out as soon as that stage has been computed (`forward_streamed`), so the D2H traffic — 3.75x the
    H2D bytes for this backbone — runs underneath the remaining stages and the following chunks.
    Input must be a pinned fp32 [B, 3, H, W] tensor; outputs are pinned fp32 NCHW maps reused across calls."""

    def __init__(self, model, chunk: int = 4):
        self.model = model
        self.chunk = int(chunk)
        self.dev = next(model.parameters()).device
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._host_out = None
        self._dev_in = None

    def _ensure_host_out(self, B, k, o):
        if self._host_out is None or self._host_out_batch != B:
            self._host_out, self._host_out_batch = {}, B
        h = self._host_out.get(k)
        if h is None or h.shape[1:] != o.shape[1:]:
            h = torch.empty((B,) + tuple(o.shape[1:]), dtype=o.dtype).pin_memory()
            self._host_out[k] = h
        return h

    @torch.no_grad()
    def __call__(self, host_img: torch.Tensor):
        if host_img.is_cuda or not host_img.is_pinned():
            raise ValueError("HostPipeline wants a pinned host tensor")
        B = host_img.shape[0]
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        n_in = min(self.chunk, B)
        if self._dev_in is None or self._dev_in[0].shape[1:] != host_img.shape[1:] or self._dev_in[0].shape[0] < n_in:
            self._dev_in = [torch.empty((n_in,) + tuple(host_img.shape[1:]), device=self.dev) for _ in range(2)]
        in_free = [None, None]                      # event: compute finished reading input buffer k
        live = []                                   # keeps device outputs alive until the final sync
        for ci, b0 in enumerate(range(0, B, self.chunk)):
            b1 = min(B, b0 + self.chunk)
            k = ci & 1
            with torch.cuda.stream(self.s_in):
                if in_free[k] is not None:
                    self.s_in.wait_event(in_free[k])
                xin = self._dev_in[k][: b1 - b0]
                xin.copy_(host_img[b0:b1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.s_in)

            def on_output(idx, o, b0=b0, b1=b1):
                done = torch.cuda.Event()
                done.record(self.s_cmp)
                h = self._ensure_host_out(B, idx, o)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(done)
                    h[b0:b1].copy_(o, non_blocking=True)

            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ready)
                outs = self.model.forward_streamed(xin, on_output)
                done_in = torch.cuda.Event()
                done_in.record(self.s_cmp)
                in_free[k] = done_in
            live.append(outs)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.synchronize()
        return [self._host_out[i] for i in range(len(live[0]))]
