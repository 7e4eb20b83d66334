"""Copy-only probe of the box's host link: every rank copies the bench step's 755 MB of fp32 feature maps device -> pinned
host (and 201 MB host -> device) in a loop, no compute.  Tells whether the end-to-end rate of bench.py at N GPUs is the
box's aggregate pinned-host DMA ceiling or an artefact of HostPipeline.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py   (or plain python for N = 1)"""
import json
import os

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
D2H, H2D = 754_974_720, 201_326_592
dbuf = torch.empty(D2H, dtype=torch.uint8, device=dev)
hbuf = torch.empty(D2H, dtype=torch.uint8).pin_memory()
hin = torch.empty(H2D, dtype=torch.uint8).pin_memory()
din = torch.empty(H2D, dtype=torch.uint8, device=dev)
s_out, s_in = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(both, iters=10):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_out.wait_stream(torch.cuda.current_stream())
    s_in.wait_stream(torch.cuda.current_stream())
    for _ in range(iters):
        with torch.cuda.stream(s_out):
            hbuf.copy_(dbuf, non_blocking=True)
        if both:
            with torch.cuda.stream(s_in):
                din.copy_(hin, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s_out)
    torch.cuda.current_stream().wait_stream(s_in)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for _ in range(2):
    run(True, 2)
ms_d2h = run(False)
ms_both = run(True)
if rank == 0:
    print(json.dumps({"n_gpus": world, "d2h_only_ms": ms_d2h, "d2h_only_gbs_per_gpu": D2H / ms_d2h / 1e6,
                      "d2h_only_gbs_aggregate": world * D2H / ms_d2h / 1e6, "d2h_plus_h2d_ms": ms_both,
                      "aggregate_gbs_both_directions": world * (D2H + H2D) / ms_both / 1e6,
                      "images_per_s_ceiling_32_per_gpu": world * 32 / (ms_both * 1e-3)}), flush=True)
if world > 1:
    dist.destroy_process_group()
