"""PanoSwin-B (E=128, depths 2-2-18-2, heads 4-8-16-32, window 7) at 1024x2048, device-resident
bf16 forward timing (CUDA graphs) at a few batch sizes.  Random-init weights, synthetic images."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import panoswintransformerobjectdetection_b200 as P  # noqa: E402
from panoswintransformerobjectdetection_b200.runtime import GraphedForward  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = P.SimplePanoSwinTransformer(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=7, ape=True,
                                patch_norm=True, pano_mode=True, drop_path_rate=0.0)
m.init_weights(None)
m.to(dev).eval()
m.set_compute_dtype("bf16")
print("params", sum(p.numel() for p in m.parameters()))
for bs in (4, 8, 16):
    img = torch.rand(bs, 3, 1024, 2048, device=dev)
    g = GraphedForward(m, tuple(img.shape), dev)
    g.static_in.copy_(img)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        outs = g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"PanoSwin-B 1024x2048 batch {bs}: {ms:.2f} ms/step  {bs / ms * 1e3:.1f} images/s  "
          f"outputs {[tuple(o.shape) for o in outs]}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    del g
    torch.cuda.empty_cache()
