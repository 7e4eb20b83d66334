"""PCIe copy rates and the end-to-end pipeline at several chunk sizes."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from panoswintransformerobjectdetection_b200.runtime import HostPipeline

dev = torch.device("cuda", 0)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("H2D GB/s", 0.268435456 / (t(lambda: d.copy_(h, non_blocking=True)) * 1e-3))
print("D2H GB/s", 0.268435456 / (t(lambda: h.copy_(d, non_blocking=True)) * 1e-3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
print("duplex: ms for 256MiB each way", t(both))
model = bench.build_model(dev)
img = torch.rand(32, 3, 512, 1024).pin_memory()
dimg = img.to(dev)
for sizes in ([4, 8, 8, 8, 4], [4] * 8, [2, 4, 6, 8, 8, 4], [4, 6, 6, 8, 8], [3, 5, 8, 8, 8], [8] * 4, [2, 6, 8, 8, 8], [6, 6, 6, 6, 8]):
    pipe = HostPipeline(model, graphs=True, sizes=sizes)
    ms = min(t(lambda: pipe(img), 5) for _ in range(2))
    print(f"pipeline sizes {sizes}: {ms:.2f} ms -> {32 / ms * 1e3:.0f} img/s", flush=True)
    del pipe
    torch.cuda.empty_cache()
