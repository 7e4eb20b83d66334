"""Kernel-level profile of the graphed training step (runtime.GraphedTrainStep, PanoSwin-T, 4x3x512x1024, bf16):
torch.profiler over a few replays, kernels aggregated by name.  usage: python tools/train_graph_profile.py [batch]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py"] + sys.argv[1:2]
import bench  # noqa: E402
from panoswintransformerobjectdetection_b200.runtime import GraphedTrainStep  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
m = bench.build_model(dev, "fp32")
m.train()
img = torch.rand(B, 3, 512, 1024, device=dev)
ts = GraphedTrainStep(m, lambda outs: sum(o.square().mean() for o in outs), img,
                      lambda ps: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.05, capturable=True, fused=True))
for _ in range(3):
    ts.step()
torch.cuda.synchronize()
N = 5
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        ts.step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:90]]
        a[0] += 1
        a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"graphed training step B={B}: {tot / N / 1e3:.2f} ms of kernels per step")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"  {v[1] / N / 1e3:7.3f} ms  x{v[0] // N:<4d} {k}")
