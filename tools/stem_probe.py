import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panoswintransformerobjectdetection_b200 import ops
dev = "cuda:0"
img = torch.rand(32, 3, 512, 1024, device=dev)
w = torch.randn(32, 27, device=dev) * 0.2
b = torch.randn(32, device=dev) * 0.1
for _ in range(3):
    y = ops.stem_conv3x3_relu(img, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    y = ops.stem_conv3x3_relu(img, w, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"stem conv1: {ms*1e3:.1f} us  {img.numel()*4/1e9/ms*1e3 + y.numel()*2/1e9/ms*1e3:.0f} GB/s")
x2 = y
w2 = (torch.randn(9, 64, 32, device=dev) / 17).bfloat16()
b = torch.randn(64, device=dev) * 0.1
outs = [torch.empty_like(x2) for _ in range(2)]
for _ in range(3):
    z = ops.stem_conv3x3_c32_relu(x2, w2, b)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    z = ops.stem_conv3x3_c32_relu(x2, w2, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"stem conv2: {ms*1e3:.1f} us  {(x2.numel() + z.numel())*2/1e9/ms*1e3:.0f} GB/s")
