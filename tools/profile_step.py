"""One warm-up + N forwards of the bench workload (PanoSwin-T, bf16) — the short command that is run plain and
then under ncu (launch list / --set full captures).  usage: python tools/profile_step.py [batch] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
img = torch.rand(batch, 3, bench.IMG_H, bench.IMG_W, device=dev)
for _ in range(1 + steps):
    outs = model(img)
torch.cuda.synchronize()
print("ok", [tuple(o.shape) for o in outs])
