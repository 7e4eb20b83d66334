"""Where a training step spends its time: CUDA-event time per libpanoswin_b200 entry point (ops.set_tracer) against the
whole step.  usage: python tools/train_profile.py [batch]"""
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from panoswintransformerobjectdetection_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
m = bench.build_model(dev)
m.train()
opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0.05)
img = torch.rand(B, 3, 512, 1024, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    loss = sum(o.square().mean() for o in m(img))
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
rec = []
ops.set_tracer(lambda fn, args, e0, e1: rec.append((fn, e0, e1)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
step()
e1.record()
torch.cuda.synchronize()
ops.set_tracer(None)
tot = collections.Counter()
cnt = collections.Counter()
for fn, a, b in rec:
    tot[fn] += a.elapsed_time(b)
    cnt[fn] += 1
whole = e0.elapsed_time(e1)
print(f"training step B={B}: {whole:.1f} ms, of which libpanoswin_b200 entry points {sum(tot.values()):.1f} ms")
for fn, ms in tot.most_common():
    print(f"  {fn:32s} x{cnt[fn]:<4d} {ms:8.2f} ms")
