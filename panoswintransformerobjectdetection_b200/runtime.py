"""Host-side runtime around the backbone: batch sharding across ranks (the path's only multi-GPU split,
SURVEY.md §8e — images never interact, so there is no data-path collective), a pinned-memory pipeline
that overlaps H2D copies, the forward and the D2H read-back of the feature maps on separate CUDA streams, and the
data-parallel training step as CUDA-graph replays around one NCCL gradient all-reduce (GraphedTrainStep)."""
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


class GraphedForward:
    """The backbone forward for one fixed input shape captured into CUDA graphs: a replay re-issues the ~100 kernel
    launches of a forward without any Python / ctypes work (small chunks are launch-bound otherwise).
    The forward is cut into one graph per output map (all in one memory pool, replayed in capture order), so a
    caller can start reading output k while the stages after it still run (`replay(on_output)`).
    `static_in` is the graphs' input buffer, `static_out` the tuple of output maps.

    A graph bakes in raw pointers to the model's cached bf16 weight copies, folded stem weights and bias tables, and
    the compute dtype / pano mode of capture time.  The capture therefore records `model.cache_signature()` and keeps
    strong references to every cached tensor it uses; `replay()` compares the signature first and RE-CAPTURES when a
    parameter changed (load_state_dict, optimizer step), or the compute dtype / pano mode was switched — a stale
    replay would silently produce old results or read recycled memory."""

    def __init__(self, model, shape, device):
        self.model = model
        self.device = device
        self.static_in = torch.zeros(shape, device=device)
        self.recaptures = 0
        self._capture()

    def _capture(self):
        model, device = self.model, self.device
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side), torch.no_grad():      # warm-up: builds every cached constant / bf16 weight copy
            for _ in range(2):
                model(self.static_in)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graphs = []
        n_out = len(tuple(model.out_indices))
        pool = torch.cuda.graph_pool_handle()

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool)
            self.graphs.append(g)

        def on_output(k, _fmap):                            # cut after every output map but the last
            if k < n_out - 1:
                self.graphs[-1].capture_end()
                begin()

        with torch.cuda.stream(side), torch.no_grad():
            begin()
            try:
                self.static_out = model.forward_streamed(self.static_in, on_output)
            finally:
                self.graphs[-1].capture_end()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.signature = model.cache_signature()
        # the tensors whose addresses the graphs hold: converted weights and this resolution's constants
        self._keep = (dict(model._weight_cache), dict(model._cur_res))

    def replay(self, on_output=None):
        """Re-run the forward on the current stream; `on_output(k, map)` is called right after the graph that
        produces output k has been launched."""
        if self.model.cache_signature() != self.signature:
            self.recaptures += 1
            self._capture()
        for k, g in enumerate(self.graphs):
            g.replay()
            if on_output is not None:
                on_output(k, self.static_out[k])
        return self.static_out


class HostPipeline:
    """model(images on the host) -> feature maps on the host, chunked and overlapped.

    Three streams: copy-in (H2D of chunk i+1), compute (forward of chunk i), copy-out (the feature maps are 3.75x
    the input bytes, so the D2H stream is the critical resource: the copy of output map k of a chunk starts as soon
    as stage k has run, while the later stages still compute).  With `graphs=True` (default) each of the two chunk
    slots owns the CUDA graphs of the forward, so chunks can be small (early first D2H, short tail) without
    becoming launch-bound.  Input must be a pinned fp32 [B, 3, H, W] tensor; outputs are
    pinned fp32 NCHW maps reused across calls."""

    def __init__(self, model, chunk: int = 8, graphs: bool = True, sizes: Sequence[int] = None, out_dtype=torch.float32):
        """`out_dtype=torch.bfloat16` is an opt-in OUTSIDE the reference contract (its forward returns fp32 maps, :977):
        the maps are rounded to bf16 on the device before the read-back, which halves the bytes on the host link -- the
        resource that bounds the end-to-end rate."""
        self.model = model
        self.out_dtype = out_dtype
        self.sizes = list(sizes) if sizes else None     # explicit chunk sizes (must sum to the batch), else _schedule
        self.chunk = int(chunk)
        self.graphs = graphs
        self.dev = next(model.parameters()).device
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_cmp = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self._host_out = None
        self._host_out_batch = None
        self._slots = {}                              # (slot, chunk shape) -> GraphedForward / input buffer

    def _ensure_host_out(self, B, k, o):
        if self._host_out is None or self._host_out_batch != B:
            self._host_out, self._host_out_batch = {}, B
        h = self._host_out.get(k)
        if h is None or h.shape[1:] != o.shape[1:]:
            h = torch.empty((B,) + tuple(o.shape[1:]), dtype=o.dtype).pin_memory()
            self._host_out[k] = h
        return h

    def _schedule(self, B):
        """Chunk boundaries.  The copy-out stream is the critical resource: it should start early and then never
        starve, so the chunks ramp up (c/2, 3c/4, 3c/4, c, c, ...): a small first chunk puts the first feature map
        on the wire early, and no chunk takes much longer to compute than its predecessor takes to copy out."""
        c = self.chunk
        if self.sizes is not None and sum(self.sizes) == B:
            sizes = list(self.sizes)
        elif B > 2 * c and c >= 4:
            sizes, left = [], B
            for n in (c // 2, (3 * c) // 4, (3 * c) // 4):
                sizes.append(n)
                left -= n
            while left > 0:
                sizes.append(min(c, left))
                left -= sizes[-1]
            if len(sizes) > 1 and sizes[-1] < c // 4:        # fold a tiny tail into the previous chunk
                sizes[-2] += sizes.pop()
        else:
            sizes, left = [], B
            while left > 0:
                sizes.append(min(c, left))
                left -= sizes[-1]
        bounds, b0 = [], 0
        for n in sizes:
            bounds.append((b0, b0 + n))
            b0 += n
        return bounds

    def _slot(self, k, shape):
        key = (k, tuple(shape))
        s = self._slots.get(key)
        if s is None:
            s = GraphedForward(self.model, shape, self.dev) if self.graphs else torch.empty(shape, device=self.dev)
            self._slots[key] = s
        return s

    @torch.no_grad()
    def __call__(self, host_img: torch.Tensor):
        if host_img.is_cuda or not host_img.is_pinned():
            raise ValueError("HostPipeline wants a pinned host tensor")
        B = host_img.shape[0]
        bounds = self._schedule(B)
        slots = [self._slot(ci & 1, (b1 - b0,) + tuple(host_img.shape[1:])) for ci, (b0, b1) in enumerate(bounds)]
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        in_free, out_free = {}, {}                    # per slot object: input consumed / outputs copied out
        live = []
        n_out = 0
        for (b0, b1), slot in zip(bounds, slots):
            xin = slot.static_in if self.graphs else slot
            with torch.cuda.stream(self.s_in):
                if id(slot) in in_free:
                    self.s_in.wait_event(in_free[id(slot)])
                xin.copy_(host_img[b0:b1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ready)
                if id(slot) in out_free:              # the slot's static outputs are still being copied out
                    self.s_cmp.wait_event(out_free[id(slot)])

                def copy_out(idx, o, b0=b0, b1=b1):   # output idx is enqueued: its D2H may start while later stages run
                    ev = torch.cuda.Event()
                    ev.record(self.s_cmp)
                    self.s_out.wait_event(ev)
                    with torch.cuda.stream(self.s_out):
                        if self.out_dtype != o.dtype:
                            from . import ops
                            o = ops.cast(o, self.out_dtype)       # our cast kernel, on the copy-out stream
                        self._ensure_host_out(B, idx, o)[b0:b1].copy_(o, non_blocking=True)

                outs = slot.replay(copy_out) if self.graphs else self.model.forward_streamed(xin, copy_out)
                done = torch.cuda.Event()
                done.record(self.s_cmp)
                in_free[id(slot)] = done
            copied = torch.cuda.Event()
            copied.record(self.s_out)
            out_free[id(slot)] = copied
            live.append(outs)
            n_out = len(outs)
        cur.wait_stream(self.s_out)
        cur.wait_stream(self.s_cmp)
        cur.synchronize()
        return [self._host_out[i] for i in range(n_out)]


class _null_context:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


class GraphedTrainStep:
    """One data-parallel training step as two CUDA-graph replays around one gradient all-reduce:

        graph A   forward, loss, backward, one multi-tensor copy of the gradients into ONE flat buffer
        NCCL      all_reduce(flat gradient, AVG) over the data-parallel group (skipped for a single process)
        graph B   optimizer step (capturable AdamW)

    The counterpart of the reference's training loop (mmdet/apis/train.py:91-99: MMDistributedDataParallel around the
    detector, mmdet/utils/optimizer.py:22-33: the optimizer hook's backward + step) for a fixed input shape.  An eager
    step issues about a thousand launches through autograd and is bound by the host; eight ranks on one box share the
    host cores, so the eager DDP step was measured at 21-27 ms against 13-14 ms of GPU work.  Replaying graphs removes
    the host from the step; the 110 MB gradient buffer crosses NVSwitch in one collective (no bucketing: a single
    all-reduce of that size runs at link rate, and with the host out of the way there is nothing left to overlap it
    with that would be worth a second stream).  Semantics match DistributedDataParallel: parameters are broadcast from
    rank 0 at construction, gradients are averaged, parameters that receive no gradient are left out of the optimizer.

    `loss_fn(outputs) -> scalar`; `make_optimizer(params) -> torch.optim.Optimizer` must build a capturable optimizer.
    `graphs=False` runs the same three phases eagerly (CPU / gloo tests, debugging)."""

    def __init__(self, model, loss_fn, example_input, make_optimizer, group=None, graphs: bool = True, warmup: int = 3):
        import torch.distributed as dist
        self.model, self.loss_fn, self.group, self.graphs = model, loss_fn, group, graphs
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.static_in = example_input.clone()
        if self.world > 1:
            for p in model.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            for b in model.buffers():
                dist.broadcast(b.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        # which parameters take part: one eager pass with .grad unset (on the side stream the warm-up and the capture use,
        # so no AccumulateGrad node remembers another stream)
        for p in model.parameters():
            p.grad = None
        side = None
        if graphs:
            side = torch.cuda.Stream(example_input.device)
            side.wait_stream(torch.cuda.current_stream(example_input.device))
        with torch.cuda.stream(side) if graphs else _null_context():
            self.loss_fn(model(self.static_in)).backward()
        if graphs:
            torch.cuda.current_stream(example_input.device).wait_stream(side)
            torch.cuda.synchronize(example_input.device)
        self.params = [p for p in model.parameters() if p.requires_grad and p.grad is not None]
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(total, dtype=self.params[0].dtype, device=self.params[0].device)
        off = 0
        for p in model.parameters():
            p.grad = None
        self.views = []
        for p in self.params:
            self.views.append(self.flat_grad[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.optimizer = make_optimizer(self.params)
        self.loss = None
        self._ga = self._gb = None
        if graphs:
            dev = self.flat_grad.device
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(warmup):                               # allocator, optimizer state, NCCL communicator
                    self._fwd_bwd()
                    self._allreduce()
                    self.optimizer.step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.loss = None                                          # drop the warm-up's autograd graph before capturing
            self._ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._ga):
                self._fwd_bwd()
            self._gb = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._gb, pool=self._ga.pool()):
                self.optimizer.step()

    def _fwd_bwd(self):
        # .grad unset: autograd hands every parameter its freshly computed gradient (no zero-fill, no read-modify-write
        # accumulation -- 170 small launches per step otherwise), then ONE multi-tensor copy gathers them in the flat buffer
        # and .grad is re-pointed at the buffer's views for the all-reduce and the optimizer
        for p in self.params:
            p.grad = None
        self.loss = self.loss_fn(self.model(self.static_in))
        self.loss.backward()
        torch._foreach_copy_(self.views, [p.grad for p in self.params])
        for p, v in zip(self.params, self.views):
            p.grad = v

    def _allreduce(self):
        if self.world > 1:
            import torch.distributed as dist
            # SUM then scale (gloo has no AVG); the scale is one pass over 110 MB at HBM rate
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
            self.flat_grad.mul_(1.0 / self.world)

    def step(self, batch: torch.Tensor = None, sync_gradients: bool = True) -> torch.Tensor:
        """One optimizer step on `batch` (copied into the static input; None re-uses the last one).  Returns the loss
        tensor of this rank (device-resident; it is overwritten by the next step)."""
        if batch is not None:
            self.static_in.copy_(batch, non_blocking=True)
        if self.graphs:
            self._ga.replay()
        else:
            self._fwd_bwd()
        if sync_gradients:
            self._allreduce()
        if self.graphs:
            self._gb.replay()
        else:
            self.optimizer.step()
        return self.loss
