"""Times psw_window_attn_bwd (bf16) at the four PanoSwin-T stage shapes.  usage: python tools/attn_bwd_probe.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panoswintransformerobjectdetection_b200 import ops  # noqa: E402
from panoswintransformerobjectdetection_b200.backbone import make_uv_hw2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
DEV = "cuda:0"
for (H, W, C, heads) in [(128, 256, 96, 3), (64, 128, 192, 6), (32, 64, 384, 12), (16, 32, 768, 24)]:
    qkv = torch.randn(B, H, W, 3 * C, device=DEV).bfloat16()
    dout = torch.randn(B, H, W, C, device=DEV).bfloat16()
    alpha = torch.randn(169, heads, device=DEV) * 0.1
    beta = torch.randn(169, heads, device=DEV) * 0.1
    qb = torch.randn(3 * C, device=DEV) * 0.1
    uv = make_uv_hw2(H, W).to(DEV)
    for _ in range(3):
        ops.window_attention_bwd(qkv, dout, alpha, beta, qb, uv, None, heads, 7, 3, True, 32 ** -0.5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.window_attention_bwd(qkv, dout, alpha, beta, qb, uv, None, heads, 7, 3, True, 32 ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"attn bwd B{B} {H}x{W} C{C} h{heads}: {us:.1f} us  ({B * H * W * C * 2 * 8 / us / 1e3:.0f} GB/s of qkv + dout + dqkv)", flush=True)
