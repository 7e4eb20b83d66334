"""GPU diagnostics with numbers instead of pass/fail (first contact with new tcgen05 kernels).
usage: python tools/gpu_diag.py {linear|attn|attnsimt}   — dumps tensors to gpurun_out/ for offline analysis."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _expect import attention_core, rel_l2  # noqa: E402
from oracle import panoswin_oracle as O  # noqa: E402
from panoswintransformerobjectdetection_b200 import ops  # noqa: E402

DEV = "cuda:0"
OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def stats(name, got, want):
    got = got.float().cpu()
    nan = int((~torch.isfinite(got)).sum())
    g = torch.nan_to_num(got)
    print(f"{name}: rel_l2={rel_l2(g, want):.3e} max_abs={float((g - want.float()).abs().max()):.3e} nonfinite={nan} "
          f"|got|={float(g.norm()):.3e} |want|={float(want.norm()):.3e}", flush=True)


def linear():
    for (M, N, K) in [(128, 96, 64), (128, 96, 96), (128, 16, 16), (256, 256, 128), (300, 288, 96), (1000, 384, 768)]:
        g = torch.Generator().manual_seed(M + N + K)
        x = torch.randn(M, K, generator=g).bfloat16()
        w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
        b = torch.randn(N, generator=g)
        want = F.linear(x.float(), w.float(), b)
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), out_dtype=torch.float32)
        torch.cuda.synchronize()
        stats(f"linear_tc M{M} N{N} K{K}", got, want)
        if (M, N, K) == (128, 96, 96):
            torch.save({"got": got.cpu(), "want": want, "x": x, "w": w, "b": b}, os.path.join(OUT, "linear_dump.pt"))


def attn(impl):
    for (H, W, heads, shift, pano) in [(7, 14, 2, 0, True), (16, 32, 2, 3, True), (13, 25, 3, 3, True), (12, 31, 2, 3, False)]:
        g = torch.Generator().manual_seed(H + W)
        C = heads * 32
        qkv = torch.randn(2, H, W, 3 * C, generator=g).bfloat16()
        alpha = torch.randn(169, heads, generator=g) * 0.5
        beta = torch.randn(169, heads, generator=g) * 0.5
        qb = torch.randn(3 * C, generator=g) * 0.5
        uv = O.uv_grid(H, W) if pano else torch.zeros(H, W, 2)
        want = attention_core(qkv.float(), alpha, beta, qb.bfloat16().float(), uv, H, W, heads, 7, shift, pano, 32 ** -0.5)
        mask = O.planar_shift_mask(H, W, 7, shift).to(DEV) if (not pano and shift) else None
        got = ops.window_attention(qkv.to(DEV), alpha.to(DEV), beta.to(DEV), qb.to(DEV), uv.to(DEV) if pano else None, mask,
                                   heads, 7, shift, pano, 32 ** -0.5, impl=impl)
        torch.cuda.synchronize()
        stats(f"attn[{impl}] H{H} W{W} heads{heads} shift{shift} pano{pano}", got, want)
        if (H, W) == (7, 14):
            torch.save({"got": got.cpu(), "want": want, "qkv": qkv, "alpha": alpha, "beta": beta, "qb": qb},
                       os.path.join(OUT, f"attn_dump_{impl}.pt"))


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "linear":
        linear()
    else:
        attn({"attn": None, "attnsimt": "simt"}[what])
