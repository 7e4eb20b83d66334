"""Micro-benchmarks of single ops at the bench shapes (CUDA events, L2 flushed between iterations by cycling
through enough distinct buffers).  usage: python tools/microbench.py {attn|linear|ln|all} [iters]"""
import os
import sys

os.environ["PSW_DIAGNOSTICS"] = "1"       # every call goes to the -DPSW_DIAGNOSTICS build (libpanoswin_b200_diag.so)

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from panoswintransformerobjectdetection_b200 import ops  # noqa: E402
from panoswintransformerobjectdetection_b200.backbone import make_uv_hw2  # noqa: E402

DEV = "cuda:0"
B = int(os.environ.get("MB_BATCH", "32"))


_WARM = [False]


def warm_clocks(seconds=1.5):
    """The SM clock of an idle GPU takes a while to ramp up: keep the GPU busy before the first timing."""
    if _WARM[0]:
        return
    import time
    a = torch.randn(4096, 4096, device=DEV, dtype=torch.bfloat16)
    t0 = time.time()
    while time.time() - t0 < seconds:
        for _ in range(20):
            a @ a
        torch.cuda.synchronize()
    _WARM[0] = True


def time_op(fn, n_bufs, iters):
    warm_clocks()
    for i in range(3):
        fn(i % n_bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i % n_bufs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


STAGES = [(128, 256, 96, 3), (64, 128, 192, 6), (32, 64, 384, 12), (16, 32, 768, 24)]


def _attn_setup(H, W, C, heads, shift=3):
    nb = max(2, int(400e6 // (B * H * W * 4 * C * 2)) + 1)
    qkv = [torch.randn(B, H, W, 3 * C, device=DEV).bfloat16() for _ in range(nb)]
    out = [torch.empty(B, H, W, C, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
    alpha = torch.randn(169, heads, device=DEV) * 0.1
    beta = torch.randn(169, heads, device=DEV) * 0.1
    qb = torch.randn(3 * C, device=DEV) * 0.1
    bf = ops.window_bias_full(alpha, beta, make_uv_hw2(H, W).to(DEV), None, H, W, 7, shift, True)
    return nb, qkv, out, qb, bf


def _attn_diag(lib, qkv, out, bf, qb, H, W, C, heads, shift, ph=None, mode=0, variant=0):
    from panoswintransformerobjectdetection_b200 import _lib
    rc = lib.psw_diag_window_attn_full(qkv.data_ptr(), out.data_ptr(), bf.data_ptr(), qb.data_ptr(), B, H, W, C, heads, 7, shift, 1,
                                       32 ** -0.5, None if ph is None else ph.data_ptr(), mode, variant,
                                       torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "psw_diag_window_attn_full")


def attn(iters):
    """The production attention kernel per stage."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (H, W, C, heads) in STAGES:
        for shift in (0, 3):
            nb, qkv, out, qb, bf = _attn_setup(H, W, C, heads, shift)
            byt = B * H * W * C * 8
            res = []
            us = min(time_op(lambda i: ops.window_attention_full(qkv[i], bf, qb, heads, 7, shift, True, 32 ** -0.5, out=out[i]), nb, iters)
                     for _ in range(2))
            res.append(f"default {us:7.1f} us {byt / us / 1e3:6.0f} GB/s")
            print(f"attn B{B} {H}x{W} C{C} h{heads} s{shift}: " + "   ".join(res), flush=True)


def linear(iters):
    shapes = [(32768, 288, 96, 0, 0, "bf16"), (32768, 96, 96, 0, 1, "f32"), (32768, 384, 96, 1, 0, "bf16"), (32768, 96, 384, 0, 1, "f32"),
              (8192, 576, 192, 0, 0, "bf16"), (8192, 768, 192, 1, 0, "bf16"), (8192, 192, 768, 0, 1, "f32"),
              (2048, 1152, 384, 0, 0, "bf16"), (2048, 1536, 384, 1, 0, "bf16"), (2048, 384, 1536, 0, 1, "f32"),
              (512, 2304, 768, 0, 0, "bf16"), (512, 3072, 768, 1, 0, "bf16"), (512, 768, 3072, 0, 1, "f32")]
    if os.environ.get("MB_SHAPES"):
        shapes = [shapes[int(i)] for i in os.environ["MB_SHAPES"].split(",")]
    for (tok, N, K, gelu, res, odt) in shapes:
        M = tok * B
        od = torch.bfloat16 if odt == "bf16" else torch.float32
        nb = max(2, int(400e6 // (M * (K * 2 + N * od.itemsize))) + 1)
        x = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nb)]
        y = [torch.randn(M, N, device=DEV).to(od) for _ in range(nb)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        us = time_op(lambda i: ops.linear(x[i], w, b, residual=y[i] if res else None, gelu=bool(gelu), out=y[i], out_dtype=od),
                     nb, iters)
        byt = M * K * 2 + N * K * 2 + M * N * od.itemsize * (2 if res else 1)
        fl = 2.0 * M * N * K
        print(f"linear M{M} N{N} K{K} gelu{gelu} res{res} {odt}: {us:8.1f} us  {byt / us / 1e3:7.0f} GB/s  {fl / us / 1e6:7.1f} TF/s", flush=True)


def ln(iters):
    for (tok, C) in [(32768, 96), (8192, 192), (2048, 384), (512, 768)]:
        rows = tok * B
        nb = max(2, int(400e6 // (rows * C * 6)) + 1)
        x = [torch.randn(rows, C, device=DEV) for _ in range(nb)]
        y = [torch.empty(rows, C, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
        g, bb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        us = time_op(lambda i: ops.layernorm(x[i], g, bb, 1e-5, torch.bfloat16, out=y[i]), nb, iters)
        print(f"layernorm rows{rows} C{C}: {us:8.1f} us  {rows * C * 6 / us / 1e3:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    if what in ("attn", "all"):
        attn(iters)
    if what in ("linear", "all"):
        linear(iters)
    if what in ("ln", "all"):
        ln(iters)


def attn_phases():
    """Per-phase cycle breakdown of one CTA of the tcgen05 attention kernel (psw_diag_window_attn_full)."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (H, W, C, heads) in STAGES:
        nb, qkv, out, qb, bf = _attn_setup(H, W, C, heads)
        ph = torch.zeros(6, dtype=torch.int64, device=DEV)
        for _ in range(2):
            _attn_diag(lib, qkv[0], out[0], bf, qb, H, W, C, heads, 3, ph=ph)
        torch.cuda.synchronize()
        v = ph.tolist()
        n = max(v[5], 1)
        names = ["wait-S", "softmax", "PV-mma", "S-issue+loads", "store"]
        print(f"attn phases {H}x{W} C{C}: steps {v[5]}  " + "  ".join(f"{nm} {v[i] / n:.0f}" for i, nm in enumerate(names)) +
              f"  total/step {sum(v[:5]) / n:.0f} cyc", flush=True)


def attn_skeleton(iters=20):
    """The attention kernel's memory skeleton (gathers + stores only) vs the full kernel and its ablations."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (H, W, C, heads) in STAGES:
        nb, qkv, out, qb, bf = _attn_setup(H, W, C, heads)
        res = []
        for name, mode in (("full", 0), ("skeleton(pair schedule)", 1), ("no-bias-loads", 2), ("no-qkv-loads", 3)):
            us = min(time_op(lambda i: _attn_diag(lib, qkv[i], out[i], bf, qb, H, W, C, heads, 3, mode=mode), nb, iters) for _ in range(2))
            res.append(f"{name} {us:.1f} us {B * H * W * C * 8 / us / 1e3:.0f} GB/s")
        print(f"attn {H}x{W} C{C}: " + "   ".join(res), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "phases":
    attn_phases()
if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "skeleton":
    attn_skeleton()


def linear_modes(iters=20):
    """GEMM diagnostics: which of loads / MMAs / stores bounds each shape (psw_diag_linear_mode), the stage-count and
    tile-width sensitivity, and the cuBLAS time of the same product (library reference, no epilogue fusion)."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    shapes = [(32768, 288, 96, 0, 0, "bf16"), (32768, 384, 96, 1, 0, "bf16"), (32768, 96, 96, 0, 1, "f32"),
              (8192, 768, 192, 1, 0, "bf16"), (2048, 1152, 384, 0, 0, "bf16"), (2048, 384, 1536, 0, 1, "f32"),
              (512, 3072, 768, 1, 0, "bf16")]
    if os.environ.get("MB_SHAPES"):
        shapes = [shapes[int(i)] for i in os.environ["MB_SHAPES"].split(",")]
    variants = [("full", 0), ("no-store", 1), ("no-load", 2), ("no-mma", 4), ("no-load/mma", 6), ("only-epi", 7), ("only-epi-nofence", 15), ("nofence", 8),
                ("stages2", 2 << 8), ("stages3", 3 << 8), ("bn64", 64 << 16), ("bn96", 96 << 16), ("bn128", 128 << 16),
                ("bn192", 192 << 16), ("bn256", 256 << 16), ("1cta", 1 << 26), ("1cta-bn256", (1 << 26) | (256 << 16)),
                ("pair", 1 << 27), ("pair-bn128", (1 << 27) | (128 << 16)),
                ("pair-bn192", (1 << 27) | (192 << 16)), ("pair-bn256", (1 << 27) | (256 << 16)), ("pair-gw4", (1 << 27) | (1 << 25))]
    for (tok, N, K, gelu, res, odt) in shapes:
        M = tok * B
        od = torch.bfloat16 if odt == "bf16" else torch.float32
        nb = max(2, int(400e6 // (M * (K * 2 + N * od.itemsize))) + 1)
        x = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nb)]
        y = [torch.randn(M, N, device=DEV).to(od) for _ in range(nb)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        line = []
        for name, mode in variants:
            if ((mode >> 16) & 0x1ff) and N % ((mode >> 16) & 0x1ff):
                continue
            lib.psw_diag_linear_mode(mode)
            try:
                us = time_op(lambda i: ops.linear(x[i], w, b, residual=y[i] if res else None, gelu=bool(gelu), out=y[i], out_dtype=od),
                             nb, iters)
                line.append(f"{name} {us:.0f}")
            except Exception as e:  # noqa: BLE001
                line.append(f"{name} ERR")
            lib.psw_diag_linear_mode(0)
        yb = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
        us = time_op(lambda i: torch.matmul(x[i], w.t(), out=yb[i]), nb, iters)
        line.append(f"cublas(no epi, bf16 out) {us:.0f}")
        print(f"linear M{M} N{N} K{K} gelu{gelu} res{res} {odt} [us]: " + "  ".join(line), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "linmodes":
    linear_modes()


def linear_cycles():
    """Per-role SM-cycle breakdown of CTA 0 of the tcgen05 GEMM (psw_diag_linear_mode bit 4)."""
    import ctypes
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    shapes = [(32768, 288, 96, 0, 0, "bf16"), (32768, 384, 96, 1, 0, "bf16"), (32768, 96, 96, 0, 1, "f32"),
              (8192, 768, 192, 1, 0, "bf16"), (2048, 1152, 384, 0, 0, "bf16"), (2048, 384, 1536, 0, 1, "f32"),
              (512, 3072, 768, 1, 0, "bf16")]
    names = ["prod wait-empty", "mma wait-tempty", "mma wait-full", "mma issue", "epi wait-tfull", "epi tmem-ld",
             "epi math", "epi stage+store"]
    for (tok, N, K, gelu, res, odt) in shapes:
        M = tok * B
        od = torch.bfloat16 if odt == "bf16" else torch.float32
        x = torch.randn(M, K, device=DEV).bfloat16()
        y = torch.randn(M, N, device=DEV).to(od)
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        for extra in (1 << 26, 1 << 27):
            lib.psw_diag_linear_mode(16 | extra)
            for _ in range(2):
                ops.linear(x, w, b, residual=y if res else None, gelu=bool(gelu), out=y, out_dtype=od)
            buf = (ctypes.c_longlong * 16)()
            _lib.check(lib.psw_diag_linear_cycles(ctypes.cast(buf, ctypes.c_void_p)), "cycles")
            lib.psw_diag_linear_mode(0)
            v = list(buf)
            n = max(v[8], 1)
            print(f"linear M{M} N{N} K{K} gelu{gelu} res{res} {odt} {'pair' if extra >> 27 else '1cta'}: tiles {v[8]}  per tile (epilogue warp 0 sees every 2nd tile): " +
                  "  ".join(f"{nm} {v[i] / n:.0f}" for i, nm in enumerate(names) if nm), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "lincycles":
    linear_cycles()


def linear_sweep(iters=20):
    """Tile-shape / CTA-pair sweep over the stage 1-3 GEMM shapes (three interleaved repeats, best of)."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    shapes = [(8192, 576, 192, 0, 0, "bf16"), (8192, 192, 192, 0, 1, "f32"), (8192, 768, 192, 1, 0, "bf16"), (8192, 192, 768, 0, 1, "f32"),
              (2048, 1152, 384, 0, 0, "bf16"), (2048, 384, 384, 0, 1, "f32"), (2048, 1536, 384, 1, 0, "bf16"), (2048, 384, 1536, 0, 1, "f32"),
              (512, 2304, 768, 0, 0, "bf16"), (512, 768, 768, 0, 1, "f32"), (512, 3072, 768, 1, 0, "bf16"), (512, 768, 3072, 0, 1, "f32"),
              (8192, 192, 384, 0, 0, "bf16"), (2048, 384, 768, 0, 0, "bf16"), (512, 768, 1536, 0, 0, "bf16")]
    variants = [("auto", 0)]
    for bn in (128, 192, 256, 384):
        variants.append((f"1cta-bn{bn}", (1 << 26) | (bn << 16)))
        variants.append((f"pair-bn{bn}", (1 << 27) | (bn << 16)))
    for (tok, N, K, gelu, res, odt) in shapes:
        M = tok * B
        od = torch.bfloat16 if odt == "bf16" else torch.float32
        nb = max(2, int(400e6 // (M * (K * 2 + N * od.itemsize))) + 1)
        x = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nb)]
        y = [torch.randn(M, N, device=DEV).to(od) for _ in range(nb)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        best = {}
        for rep in range(3):
            for name, mode in variants:
                bn = (mode >> 16) & 0x1ff
                if bn and (N % bn or bn > 256):
                    continue
                lib.psw_diag_linear_mode(mode)
                try:
                    us = time_op(lambda i: ops.linear(x[i], w, b, residual=y[i] if res else None, gelu=bool(gelu), out=y[i], out_dtype=od),
                                 nb, iters)
                    best[name] = min(best.get(name, 1e9), us)
                except Exception:  # noqa: BLE001
                    best[name] = float("nan")
                lib.psw_diag_linear_mode(0)
        print(f"linear M{M} N{N} K{K} gelu{gelu} res{res} {odt} [us]: " + "  ".join(f"{k} {v:.1f}" for k, v in best.items()), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "linsweep":
    linear_sweep()


def attn_hc(iters=20):
    """Batch-innermost schedule vs the window-pair schedule (variant 15) of the tcgen05 attention kernel."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (H, W, C, heads) in STAGES:
        nb, qkv, out, qb, bf = _attn_setup(H, W, C, heads)
        res = []
        for hc in (0, 15):
            us = min(time_op(lambda i: _attn_diag(lib, qkv[i], out[i], bf, qb, H, W, C, heads, 3, variant=hc), nb, iters) for _ in range(3))
            res.append(f"{'pair' if hc else 'batch-inner'} {us:.1f}")
        print(f"attn {H}x{W} C{C} h{heads}: " + "  ".join(res), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "attnhc":
    attn_hc()


def ln_heads(iters=20):
    """Output LayerNorm -> NCHW and patch-merge LayerNorm at the four stage shapes."""
    for (tok, C, H, W) in [(32768, 96, 128, 256), (8192, 192, 64, 128), (2048, 384, 32, 64), (512, 768, 16, 32)]:
        nb = max(2, int(400e6 // (B * tok * C * 8)) + 1)
        x = [torch.randn(B, tok, C, device=DEV) for _ in range(nb)]
        g, bb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        us = time_op(lambda i: ops.layernorm_nchw(x[i], g, bb, H, W), nb, iters)
        print(f"layernorm_nchw B{B} {H}x{W} C{C}: {us:8.1f} us  {B * tok * C * 8 / us / 1e3:7.0f} GB/s", flush=True)
        if C < 768:
            g4, b4 = torch.ones(4 * C, device=DEV), torch.zeros(4 * C, device=DEV)
            us = time_op(lambda i: ops.patch_merge_layernorm(x[i], g4, b4, H, W, out_dtype=torch.bfloat16), nb, iters)
            print(f"patch_merge_ln B{B} {H}x{W} C{C}: {us:8.1f} us  {B * tok * C * 6 / us / 1e3:7.0f} GB/s", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "lnheads":
    ln_heads()


def linear_lnf(iters=20):
    """proj / fc2 with the fused LayerNorm epilogue vs the plain GEMM + standalone LayerNorm."""
    for (tok, N, K) in [(32768, 96, 96), (32768, 96, 384), (8192, 192, 192), (8192, 192, 768)]:
        M = tok * B
        nb = max(2, int(400e6 // (M * (K * 2 + N * 10))) + 1)
        x = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nb)]
        y = [torch.randn(M, N, device=DEV) for _ in range(nb)]
        xn = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        g, bb = torch.ones(N, device=DEV), torch.zeros(N, device=DEV)
        us0 = time_op(lambda i: ops.linear(x[i], w, b, residual=y[i], out=y[i], out_dtype=torch.float32), nb, iters)
        us1 = time_op(lambda i: ops.layernorm(y[i], g, bb, 1e-5, torch.bfloat16, out=xn[i]), nb, iters)
        us2 = time_op(lambda i: ops.linear_layernorm(x[i], w, b, y[i], g, bb, 1e-5, out=y[i]), nb, iters)
        byt = M * (K * 2 + N * 10)
        print(f"M{M} N{N} K{K}: linear {us0:.1f} us + layernorm {us1:.1f} us = {us0 + us1:.1f}   fused {us2:.1f} us  {byt / us2 / 1e3:.0f} GB/s", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "lnf":
    linear_lnf()


def mlp(iters=20):
    """Fused MLP kernel (C = 96, hidden = 384) vs fc1(+GELU) and fc2(+residual) as two GEMMs."""
    M, C, Hd = 32768 * B, 96, 384
    nb = 2
    xn = [torch.randn(M, C, device=DEV).bfloat16() for _ in range(nb)]
    x = [torch.randn(M, C, device=DEV) for _ in range(nb)]
    hid = torch.empty(M, Hd, device=DEV, dtype=torch.bfloat16)
    w1 = (torch.randn(Hd, C, device=DEV) / C ** 0.5).bfloat16()
    w2 = (torch.randn(C, Hd, device=DEV) / Hd ** 0.5).bfloat16()
    b1, b2 = torch.randn(Hd, device=DEV), torch.randn(C, device=DEV)
    us1 = time_op(lambda i: ops.linear(xn[i], w1, b1, gelu=True, out=hid), nb, iters)
    us2 = time_op(lambda i: ops.linear(hid, w2, b2, residual=x[i], out=x[i], out_dtype=torch.float32), nb, iters)
    us3 = time_op(lambda i: ops.mlp_fused(xn[i], w1, b1, w2, b2, x[i]), nb, iters)
    byt = M * C * (2 + 4 + 4)
    print(f"MLP M{M}: fc1 {us1:.1f} us + fc2 {us2:.1f} us = {us1 + us2:.1f}   fused {us3:.1f} us  {byt / us3 / 1e3:.0f} GB/s", flush=True)
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for mode, name in ((2, "no-final-epilogue"), (8, "no-gelu-math"), (10, "no-gelu-math no-final-epilogue")):
        lib.psw_diag_mlp_mode(mode)
        us = time_op(lambda i: ops.mlp_fused(xn[i], w1, b1, w2, b2, x[i]), nb, iters)
        lib.psw_diag_mlp_mode(0)
        print(f"   fused {name}: {us:.1f} us", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mlp":
    mlp()


def ln_stem(iters=20):
    """Stem LayerNorm (bf16 tokens -> fp32 residual stream, + position rows) alone and with the first norm1 fused."""
    tok, C = 32768, 96
    rows = tok * B
    nb = 3
    x = [torch.randn(rows, C, device=DEV).bfloat16() for _ in range(nb)]
    g, bb = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    pos = torch.randn(tok, C, device=DEV)
    for pp, name in ((None, "no pos"), (pos, "+pos")):
        us = time_op(lambda i: ops.layernorm(x[i], g, bb, 1e-5, torch.float32, pp), nb, iters)
        print(f"stem layernorm {name}: {us:8.1f} us  {rows * C * 6 / us / 1e3:7.0f} GB/s", flush=True)
        us = time_op(lambda i: ops.layernorm2(x[i], g, bb, 1e-5, pp, g, bb, 1e-5), nb, iters)
        print(f"stem layernorm x2 {name}: {us:8.1f} us  {rows * C * 8 / us / 1e3:7.0f} GB/s", flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "lnstem":
    ln_stem()


def attn_ctas():
    """Per-CTA residency and phase cycles of the batch-innermost attention kernel."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (H, W, C, heads) in STAGES:
        nb, qkv, out, qb, bf = _attn_setup(H, W, C, heads)
        ph = torch.zeros(8 + 11 * 1024, dtype=torch.int64, device=DEV)
        warm_clocks()
        for i in range(30):
            _attn_diag(lib, qkv[i % nb], out[i % nb], bf, qb, H, W, C, heads, 3, ph=ph, variant=32)
        torch.cuda.synchronize()
        v = ph[8:8 + 3 * 1024].view(-1, 3)
        n = int((v[:, 2] > 0).sum())
        v = v[:n]
        pc = ph[8 + 3 * 1024:].view(-1, 8)[:n].float()
        t0 = int(v[:, 1].min())
        st, en = (v[:, 1] - t0).float() / 1e3, (v[:, 2] - t0).float() / 1e3
        dur = en - st
        steps = pc[:, 5].clamp(min=1)
        cyc = pc[:, :5].sum(1)
        names = ["wait-S", "softmax", "PV", "S-issue+loads", "store"]
        # the per-phase sums below cover only ~60% of the loop's cycles (the clock reads are not ordered against the
        # asynchronous work around them): read them as proportions, the lifetime / loop cycle counts as absolutes
        print(f"  clock64 / globaltimer over the CTA lifetime: {float((pc[:, 6] / dur).median()) / 1e3:.3f} GHz; loop share of the "
              f"lifetime cycles {float((pc[:, 7] / pc[:, 6]).median()):.3f}; phase-sum share {float((cyc / pc[:, 6]).median()):.3f}; "
              f"loop cycles per step {float((pc[:, 7] / steps).median()):.0f}")
        print(f"attn ctas {H}x{W} C{C}: {n} CTAs, start max {float(st.max()):.1f} us, duration min {float(dur.min()):.1f} median "
              f"{float(dur.median()):.1f} max {float(dur.max()):.1f} us, end max {float(en.max()):.1f} us; steps {float(steps.mean()):.1f}; "
              f"loop cycles median {float(cyc.median()):.0f} -> {float((cyc / dur).median()) / 1e3:.2f} GHz apparent; per step: "
              + "  ".join(f"{nm} {float((pc[:, i] / steps).mean()):.0f}" for i, nm in enumerate(names)), flush=True)
        for c in (0, 1, 2, n // 2, n - 1):
            print(f"   CTA {c}: SM {int(v[c, 0])} start {float(st[c]):.1f} end {float(en[c]):.1f} us, cycles/step "
                  + " ".join(f"{float(pc[c, i] / steps[c]):.0f}" for i in range(5)), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "attnctas":
    attn_ctas()


def linear_stage0(iters=20):
    """block_n sweep of the stage-0 qkv GEMM (M = 32 x 32768, N = 288, K = 96) and the stage-1 qkv / fc1 shapes."""
    from panoswintransformerobjectdetection_b200 import _lib
    lib = _lib.load()
    for (M, N, K, gelu) in [(B * 32768, 288, 96, 0), (B * 8192, 576, 192, 0), (B * 8192, 768, 192, 1)]:
        nb = max(2, int(400e6 // (M * (K + N) * 2)) + 1)
        x = [torch.randn(M, K, device=DEV).bfloat16() for _ in range(nb)]
        y = [torch.empty(M, N, device=DEV, dtype=torch.bfloat16) for _ in range(nb)]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=DEV)
        res = []
        for bn in (0, 64, 96, 128, 160, 192, 224, 256):
            lib.psw_diag_linear_mode(((1 << 26) | (bn << 16)) if bn else 0)
            try:
                us = min(time_op(lambda i: ops.linear(x[i], w, b, gelu=bool(gelu), out=y[i]), nb, iters) for _ in range(2))
                res.append(f"bn{bn if bn else '-auto'} {us:.1f}")
            except Exception as e:  # noqa: BLE001
                res.append(f"bn{bn} failed")
            lib.psw_diag_linear_mode(0)
        print(f"linear M{M} N{N} K{K} gelu{gelu} [us]: " + "  ".join(res), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "linstage0":
    linear_stage0()
