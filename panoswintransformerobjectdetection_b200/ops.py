"""Thin torch-tensor wrappers over the C ABI (include/panoswin_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every op below forwards raw
device pointers to libpanoswin_b200.so.  Nothing in this module computes on the CPU or through
torch kernels — if the tensors are not CUDA tensors, or the library is missing, the call raises.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import PSW_BF16, PSW_EPI_GELU, PSW_F32, PanoSwinB200Error

_launches = 0          # kernels enqueued through this module (bench.py reports it as gpu_launches)


def launch_count() -> int:
    return _launches


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return PSW_F32
    if t.dtype == torch.bfloat16:
        return PSW_BF16
    raise PanoSwinB200Error(f"unsupported tensor dtype {t.dtype} (fp32 or bf16 only)")


def _chk(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise PanoSwinB200Error("libpanoswin_b200 runs on CUDA tensors only (there is no CPU fallback)")
        if not t.is_contiguous():
            raise PanoSwinB200Error("libpanoswin_b200 needs contiguous tensors")
        dev = dev or t.device
        if t.device != dev:
            raise PanoSwinB200Error("all tensors of one call must live on the same device")
    return dev


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _f32(t, name):
    if t is not None and t.dtype != torch.float32:
        raise PanoSwinB200Error(f"{name} must be fp32")
    return t


_tracer = None         # optional callable(fn_name, args, start_event, end_event): per-launch CUDA-event timing


def set_tracer(tracer):
    """Install / remove (None) a per-launch tracer.  When set, every launch is bracketed by CUDA events
    recorded on the launching stream (bench.py uses this for the live roofline numbers)."""
    global _tracer
    _tracer = tracer


def _call(fn_name, *args):
    global _launches
    lib = _lib.load()
    if _tracer is None:
        rc = getattr(lib, fn_name)(*args)
    else:
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = getattr(lib, fn_name)(*args)
        e1.record(stream)
        _tracer(fn_name, args, e0, e1)
    _lib.check(rc, fn_name)
    _launches += 1


def layernorm(x, gamma, beta, eps=1e-5, out_dtype=None, pos=None, out=None):
    """LayerNorm over the last dim (+ optional fp32 position table `pos` [pos_rows, C] added per row
    modulo pos_rows).  Reference: norm1/norm2 (simple_panoswin_transformer.py:504,:534), :768-772, :962."""
    dev = _chk(x, gamma, beta, pos, out)
    C = x.shape[-1]
    rows = x.numel() // C
    out_dtype = out_dtype or x.dtype
    if out is None:
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    pos_rows = 0 if pos is None else pos.numel() // C
    with torch.cuda.device(dev):
        _call("psw_layernorm_fwd", _ptr(x), _ptr(out), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")),
              _ptr(_f32(pos, "pos")), rows, C, pos_rows, float(eps), _dt(x), _dt(out), _stream(dev))
    return out


def layernorm2(x, gamma, beta, eps, pos, gamma2, beta2, eps2):
    """y = LN(x) * gamma + beta (+ pos) in fp32 and y2 = LN(y) * gamma2 + beta2 in bf16, one pass over the rows."""
    dev = _chk(x, gamma, beta, pos, gamma2, beta2)
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    y2 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    pos_rows = 0 if pos is None else pos.numel() // C
    with torch.cuda.device(dev):
        _call("psw_layernorm2_fwd", _ptr(x), _ptr(y), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")), _ptr(_f32(pos, "pos")),
              _ptr(y2), _ptr(_f32(gamma2, "gamma2")), _ptr(_f32(beta2, "beta2")), rows, C, pos_rows, float(eps), float(eps2),
              _dt(x), _stream(dev))
    return y, y2


def linear(x, w, bias=None, residual=None, gelu=False, out_dtype=None, out=None):
    """act(x @ w.T + bias) (+ residual).  x [..., K], w [N, K]; fp32 tensors run the CUDA-core parity
    kernel, bf16 tensors the tcgen05 kernel.  Reference: nn.Linear at :287, :309(+:533), :55-61(+:534), :575."""
    dev = _chk(x, w, bias, residual, out)
    K = x.shape[-1]
    N = w.shape[0]
    if w.shape[1] != K:
        raise PanoSwinB200Error(f"linear: weight {tuple(w.shape)} does not match input features {K}")
    M = x.numel() // K
    if x.dtype != w.dtype:
        raise PanoSwinB200Error("linear: x and w must share a dtype")
    out_dtype = out_dtype or (residual.dtype if residual is not None else x.dtype)
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=out_dtype, device=x.device)
    if residual is not None and (residual.dtype != out.dtype or residual.numel() != out.numel()):
        raise PanoSwinB200Error("linear: residual must match the output's dtype and shape")
    with torch.cuda.device(dev):
        _call("psw_linear_fwd", _ptr(x), _ptr(w), _ptr(_f32(bias, "bias")), _ptr(residual), _ptr(out), M, N, K,
              PSW_EPI_GELU if gelu else 0, _dt(x), _dt(out), _stream(dev))
    return out


def linear_layernorm(x, w, bias, residual, ln_gamma, ln_beta, ln_eps, out=None):
    """y = x @ w^T + bias + residual (fp32; `out` may be the residual tensor itself) and LayerNorm(y) (bf16) in one
    kernel (bf16 x / w; N % 32 == 0, N <= 256).  Returns (y, ln_out)."""
    dev = _chk(x, w, bias, residual, ln_gamma, ln_beta, out)
    M, K = x.numel() // x.shape[-1], x.shape[-1]
    N = w.shape[0]
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or residual.dtype != torch.float32:
        raise PanoSwinB200Error("linear_layernorm wants bf16 x / w and an fp32 residual")
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
    ln_out = torch.empty(x.shape[:-1] + (N,), dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_linear_ln_fwd", _ptr(x), _ptr(w), _ptr(_f32(bias, "bias")), _ptr(residual), _ptr(out),
              _ptr(_f32(ln_gamma, "ln_gamma")), _ptr(_f32(ln_beta, "ln_beta")), float(ln_eps), _ptr(ln_out), M, N, K, _stream(dev))
    return out, ln_out


def linear_layernorm_nchw(x, w, bias, residual, ln_gamma, ln_beta, ln_eps, H, W, out=None):
    """y = x @ w^T + bias + residual (fp32 [B, H*W, N]) and LayerNorm(y) as the fp32 NCHW map [B, N, H, W] in one kernel
    (the last fc2 of a stage + the stage's output norm; N % 32 == 0, N <= 256).  Returns (y, map)."""
    dev = _chk(x, w, bias, residual, ln_gamma, ln_beta, out)
    K = x.shape[-1]
    M = x.numel() // K
    N = w.shape[0]
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or residual.dtype != torch.float32 or M % (H * W) != 0:
        raise PanoSwinB200Error("linear_layernorm_nchw wants bf16 x / w, an fp32 residual and M = B * H * W rows")
    if out is None:
        out = torch.empty(x.shape[:-1] + (N,), dtype=torch.float32, device=x.device)
    fmap = torch.empty((M // (H * W), N, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_linear_ln_nchw_fwd", _ptr(x), _ptr(w), _ptr(_f32(bias, "bias")), _ptr(residual), _ptr(out),
              _ptr(_f32(ln_gamma, "ln_gamma")), _ptr(_f32(ln_beta, "ln_beta")), float(ln_eps), _ptr(fmap), H * W, M, N, K,
              _stream(dev))
    return out, fmap


def mlp_fused(xn, w1, b1, w2, b2, x):
    """x <- x + fc2(GELU(fc1(xn))) in one kernel; xn bf16 [..., C], x fp32 [..., C] updated in place (C = 96, hidden = 384)."""
    dev = _chk(xn, w1, b1, w2, b2, x)
    C = xn.shape[-1]
    M = xn.numel() // C
    hidden = w1.shape[0]
    if xn.dtype != torch.bfloat16 or w1.dtype != torch.bfloat16 or w2.dtype != torch.bfloat16 or x.dtype != torch.float32:
        raise PanoSwinB200Error("mlp_fused wants bf16 xn / w1 / w2 and an fp32 x")
    if tuple(w1.shape) != (hidden, C) or tuple(w2.shape) != (C, hidden) or x.numel() != M * C:
        raise PanoSwinB200Error("mlp_fused: shape mismatch")
    with torch.cuda.device(dev):
        _call("psw_mlp_fused_fwd", _ptr(xn), _ptr(w1), _ptr(_f32(b1, "b1")), _ptr(w2), _ptr(_f32(b2, "b2")), _ptr(x), M, C, hidden,
              _stream(dev))
    return x


def window_grid(H, W, window, pano_mode):
    """(windows per column, windows per row) of the map the attention runs on (host-only helper)."""
    import ctypes
    nwh, nww = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.load().psw_window_grid(H, W, window, 1 if pano_mode else 0, ctypes.byref(nwh), ctypes.byref(nww)),
               "psw_window_grid")
    return nwh.value, nww.value


def window_attention_full_supported(window, head_dim) -> bool:
    """True when the bf16 tcgen05 attention kernel is instantiated for this window / head_dim (window 7, head_dim 32)."""
    return bool(_lib.load().psw_window_attn_full_supported(int(window), int(head_dim)))


def window_bias_full(alpha, beta, uv, mask, H, W, window, shift, pano_mode):
    """Every additive term of the logits of one block at one resolution (great-circle + relative-position bias, planar
    shift mask; times log2 e) for all windows of one image and all heads: fp32 [windows, heads, 13, 64, 4] for
    window_attention_full."""
    dev = _chk(alpha, beta, uv, mask)
    heads = alpha.shape[1]
    nbytes = _lib.load().psw_window_bias_full_bytes(H, W, heads, window, 1 if pano_mode else 0)
    out = torch.empty((nbytes // (heads * 13 * 64 * 16), heads, 13, 64, 4), dtype=torch.float32, device=alpha.device)
    with torch.cuda.device(dev):
        _call("psw_window_bias_full", _ptr(_f32(alpha, "alpha")), _ptr(_f32(beta, "beta")), _ptr(_f32(uv, "uv")),
              _ptr(_f32(mask, "mask")), _ptr(out), H, W, heads, window, shift, 1 if pano_mode else 0, _stream(dev))
    return out


def window_attention_full(qkv, bias_full, qkv_bias, heads, window, shift, pano_mode, scale, out=None):
    """The bf16 tcgen05 attention kernel with the precomputed bias table of window_bias_full() (production path).
    qkv [B, H, W, 3C] bf16 -> [B, H, W, C] bf16."""
    dev = _chk(qkv, bias_full, qkv_bias, out)
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if qkv.dtype != torch.bfloat16:
        raise PanoSwinB200Error("window_attention_full is the bf16 path")
    if out is None:
        out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(dev):
        _call("psw_window_attn_full_fwd", _ptr(qkv), _ptr(out), _ptr(bias_full), _ptr(_f32(qkv_bias, "qkv_bias")), B, H, W,
              C, heads, window, shift, 1 if pano_mode else 0, float(scale), _stream(dev))
    return out


def window_attention(qkv, alpha, beta, qkv_bias, uv, mask, heads, window, shift, pano_mode, scale, out=None):
    """Fused shift + partition + W-MSA core + reverse + un-shift on the generic CUDA-core kernel (any window size and
    head_dim).  qkv [B, H, W, 3C] -> [B, H, W, C], fp32 (parity path) or bf16 (the route for shapes the tcgen05 kernel
    of window_attention_full is not instantiated for); `uv` is required in pano mode."""
    dev = _chk(qkv, alpha, beta, qkv_bias, uv, mask, out)
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if out is None:
        out = torch.empty((B, H, W, C), dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(dev):
        _call("psw_window_attn_fwd", _ptr(qkv), _ptr(out), _ptr(_f32(alpha, "alpha")), _ptr(_f32(beta, "beta")),
              _ptr(_f32(qkv_bias, "qkv_bias")), _ptr(_f32(uv, "uv")), _ptr(_f32(mask, "mask")), B, H, W, C, heads, window,
              shift, 1 if pano_mode else 0, float(scale), _dt(qkv), _stream(dev))
    return out


def patch_merge_layernorm(x, gamma, beta, H, W, eps=1e-5, out_dtype=None):
    """x [B, H*W, C] -> LN(concat 2x2) [B, ceil(H/2)*ceil(W/2), 4C] (reference :563-574)."""
    dev = _chk(x, gamma, beta)
    B, S, C = x.shape
    if S != H * W:
        raise PanoSwinB200Error("input feature has wrong size")
    out = torch.empty((B, ((H + 1) // 2) * ((W + 1) // 2), 4 * C), dtype=out_dtype or x.dtype, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_patch_merge_ln_fwd", _ptr(x), _ptr(out), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")),
              B, H, W, C, float(eps), _dt(x), _dt(out), _stream(dev))
    return out


def layernorm_nchw(x, gamma, beta, H, W, eps=1e-5):
    """x [B, H*W, C] -> LayerNorm -> fp32 [B, C, H, W] contiguous (reference :974-978)."""
    dev = _chk(x, gamma, beta)
    B, S, C = x.shape
    if S != H * W:
        raise PanoSwinB200Error("input feature has wrong size")
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_layernorm_nchw_fwd", _ptr(x), _ptr(out), _ptr(_f32(gamma, "gamma")), _ptr(_f32(beta, "beta")),
              B, S, C, float(eps), _dt(x), _stream(dev))
    return out


def stem_conv3x3_relu(img, w_folded, bias_folded):
    """conv3x3(pad 1) + folded BatchNorm + ReLU: fp32 NCHW image [B, 3, H, W] -> bf16 NHWC [B, H, W, cout]
    (reference PatchEmbed.proj[0..2], :743-745).  w_folded [cout, 27], bias_folded [cout] fp32, cout 32 or 64 (a
    narrower layer is zero-padded by the caller)."""
    dev = _chk(img, w_folded, bias_folded)
    B, cin, H, W = img.shape
    cout = w_folded.shape[0]
    if img.dtype != torch.float32:
        raise PanoSwinB200Error("stem_conv3x3_relu wants an fp32 image")
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=img.device)
    with torch.cuda.device(dev):
        _call("psw_stem_conv3x3_relu_fwd", _ptr(img), _ptr(_f32(w_folded, "w_folded")), _ptr(_f32(bias_folded, "bias_folded")),
              _ptr(out), B, H, W, cin, cout, _stream(dev))
    return out


def stem_conv3x3_c32_relu(x_nhwc, w_taps, bias):
    """conv3x3(32 -> cout, pad 1) + folded BatchNorm + ReLU on NHWC bf16 [B, H, W, 32] -> [B, H, W, cout] (reference
    PatchEmbed.proj[3..5], :746-748).  w_taps [9, cout, 32 in] bf16 (tap = ky*3 + kx), bias [cout] fp32; cout 32 or 64."""
    dev = _chk(x_nhwc, w_taps, bias)
    B, H, W, C = x_nhwc.shape
    cout = w_taps.shape[1]
    if C != 32 or x_nhwc.dtype != torch.bfloat16 or tuple(w_taps.shape) != (9, cout, 32) or w_taps.dtype != torch.bfloat16:
        raise PanoSwinB200Error("stem_conv3x3_c32_relu wants bf16 NHWC x [B,H,W,32] and bf16 w_taps [9,cout,32]")
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    with torch.cuda.device(dev):
        _call("psw_stem_conv3x3_c32_relu_fwd", _ptr(x_nhwc), _ptr(w_taps), _ptr(_f32(bias, "bias")), _ptr(out), B, H, W,
              cout, _stream(dev))
    return out


def conv3x3_nhwc(x_nhwc, w_ohwi, bias, relu=True):
    """conv3x3(stride 1, pad 1) + bias (+ ReLU) of an NHWC bf16 image as a GEMM over shifted TMA views (reference
    PatchEmbed.proj[3..5], :746-748, for stem widths other than embed_dim 96's).  x [B, H, W, cin] -> [B, H, W, cout];
    w_ohwi [cout, 3, 3, cin] bf16, bias [cout] fp32; cin % 64 == 0 and cout % 16 == 0 (zero-pad the channels)."""
    dev = _chk(x_nhwc, w_ohwi, bias)
    B, H, W, cin = x_nhwc.shape
    cout = w_ohwi.shape[0]
    if x_nhwc.dtype != torch.bfloat16 or w_ohwi.dtype != torch.bfloat16 or tuple(w_ohwi.shape) != (cout, 3, 3, cin):
        raise PanoSwinB200Error("conv3x3_nhwc wants bf16 NHWC x and bf16 w [cout, 3, 3, cin]")
    out = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    with torch.cuda.device(dev):
        _call("psw_conv3x3_nhwc_fwd", _ptr(x_nhwc), _ptr(w_ohwi), _ptr(_f32(bias, "bias")), _ptr(out), B, H, W, cin, cout,
              1 if relu else 0, _stream(dev))
    return out


def patch_conv(x_nhwc, w_ohwi, bias, patch):
    """conv(kernel = stride = patch) of an NHWC bf16 image as a GEMM (reference PatchEmbed.proj[6], :749):
    x [B, H, W, cin] -> tokens [B, H/ph, W/pw, cout] bf16.  w_ohwi [cout, ph, pw, cin] bf16, bias [cout] fp32."""
    dev = _chk(x_nhwc, w_ohwi, bias)
    B, H, W, cin = x_nhwc.shape
    ph, pw = patch
    cout = w_ohwi.shape[0]
    if x_nhwc.dtype != torch.bfloat16 or w_ohwi.dtype != torch.bfloat16 or tuple(w_ohwi.shape) != (cout, ph, pw, cin):
        raise PanoSwinB200Error("patch_conv wants bf16 NHWC x and bf16 w [cout, ph, pw, cin]")
    out = torch.empty((B, H // ph, W // pw, cout), dtype=torch.bfloat16, device=x_nhwc.device)
    with torch.cuda.device(dev):
        _call("psw_patch_conv_fwd", _ptr(x_nhwc), _ptr(w_ohwi), _ptr(_f32(bias, "bias")), _ptr(out), B, H, W, cin, cout, ph, pw,
              _stream(dev))
    return out


def conv2d_f32(x, w, bias=None, bn_scale=None, bn_shift=None, stride=1, padding=0, relu=False, out_nhwc=False):
    """fp32 direct convolution (+ BatchNorm affine + ReLU) on CUDA cores: the stem of the fp32 parity path (reference
    PatchEmbed.proj :742-750).  x [B, cin, H, W] NCHW, w [cout, cin, k, k] -> NCHW, or NHWC tokens with out_nhwc."""
    dev = _chk(x, w, bias, bn_scale, bn_shift)
    B, cin, H, W = x.shape
    cout, _, k, k2 = w.shape
    if x.dtype != torch.float32 or w.dtype != torch.float32 or k != k2 or w.shape[1] != cin:
        raise PanoSwinB200Error("conv2d_f32 wants fp32 NCHW x and a square fp32 kernel [cout, cin, k, k]")
    Ho, Wo = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
    out = torch.empty((B, Ho, Wo, cout) if out_nhwc else (B, cout, Ho, Wo), dtype=torch.float32, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_conv2d_f32_fwd", _ptr(x), _ptr(w), _ptr(_f32(bias, "bias")), _ptr(_f32(bn_scale, "bn_scale")),
              _ptr(_f32(bn_shift, "bn_shift")), _ptr(out), B, cin, H, W, cout, k, stride, padding, 1 if relu else 0,
              1 if out_nhwc else 0, _stream(dev))
    return out


def cast(x, dtype):
    dev = _chk(x)
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_cast", _ptr(x), _ptr(out), x.numel(), _dt(x), _dt(out), _stream(dev))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# backward ops (training path; include/panoswin_b200.h "Backward entry points")
# ---------------------------------------------------------------------------------------------------------------------
def layernorm_bwd(x, dy, gamma, eps=1e-5, need_params=True):
    """Gradients of layernorm(): returns (dx [like x], dgamma, dbeta) — the last two None when not needed."""
    dev = _chk(x, dy, gamma)
    C = x.shape[-1]
    rows = x.numel() // C
    dx = torch.empty_like(x)
    dg = torch.empty(C, dtype=torch.float32, device=x.device) if need_params else None
    db = torch.empty(C, dtype=torch.float32, device=x.device) if need_params else None
    stats = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_layernorm_bwd", _ptr(x), _ptr(dy), _ptr(_f32(gamma, "gamma")), _ptr(dx), _ptr(dg), _ptr(db), _ptr(stats), rows, C,
              float(eps), _dt(x), _dt(dy), _stream(dev))
    return dx, dg, db


def patch_merge_layernorm_bwd(x, dy, gamma, H, W, eps=1e-5, need_params=True):
    """Gradients of patch_merge_layernorm(): x [B, H*W, C], dy [B, H2*W2, 4C] -> (dx [like x], dgamma [4C], dbeta [4C])."""
    dev = _chk(x, dy, gamma)
    B, S, C = x.shape
    if S != H * W:
        raise PanoSwinB200Error("input feature has wrong size")
    rows = B * ((H + 1) // 2) * ((W + 1) // 2)
    dx = torch.empty_like(x)
    dg = torch.empty(4 * C, dtype=torch.float32, device=x.device) if need_params else None
    db = torch.empty(4 * C, dtype=torch.float32, device=x.device) if need_params else None
    stats = torch.empty(2 * rows, dtype=torch.float32, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_patch_merge_ln_bwd", _ptr(x), _ptr(dy), _ptr(_f32(gamma, "gamma")), _ptr(dx), _ptr(dg), _ptr(db), _ptr(stats),
              B, H, W, C, float(eps), _dt(x), _dt(dy), _stream(dev))
    return dx, dg, db


def linear_bwd(x, w, dy, need_dx=True, need_dw=True, need_db=True, dx_dtype=None):
    """Gradients of linear() without epilogue: x [..., K], w [N, K], dy [..., N] (one dtype) ->
    (dx [..., K] in dx_dtype, dw [N, K] fp32, db [N] fp32); entries not needed are None."""
    dev = _chk(x, w, dy)
    K = w.shape[1]
    N = w.shape[0]
    M = dy.numel() // N
    if x is not None and x.dtype != w.dtype or dy.dtype != w.dtype:
        raise PanoSwinB200Error("linear_bwd: x, w and dy must share a dtype")
    dx_dtype = dx_dtype or w.dtype
    dx = torch.empty(dy.shape[:-1] + (K,), dtype=dx_dtype, device=dy.device) if need_dx else None
    dw = torch.empty((N, K), dtype=torch.float32, device=dy.device) if need_dw else None
    db = torch.empty(N, dtype=torch.float32, device=dy.device) if need_db else None
    nbytes = _lib.load().psw_linear_bwd_workspace_bytes(M, N, K, _dt(w)) if need_dx else 0
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=dy.device) if nbytes else None
    with torch.cuda.device(dev):
        _call("psw_linear_bwd", _ptr(x), _ptr(w), _ptr(dy), _ptr(dx), _ptr(dw), _ptr(db), M, N, K, _dt(w),
              PSW_F32 if dx_dtype == torch.float32 else PSW_BF16, _ptr(ws), nbytes, _stream(dev))
    return dx, dw, db


def gelu(h):
    """Exact (erf) GELU, elementwise, keeping h for gelu_bwd (training path)."""
    dev = _chk(h)
    y = torch.empty_like(h)
    with torch.cuda.device(dev):
        _call("psw_gelu_fwd", _ptr(h), _ptr(y), h.numel(), _dt(h), _stream(dev))
    return y


def gelu_bwd(h, dy):
    dev = _chk(h, dy)
    dh = torch.empty_like(h)
    with torch.cuda.device(dev):
        _call("psw_gelu_bwd", _ptr(h), _ptr(dy), _ptr(dh), h.numel(), _dt(h), _stream(dev))
    return dh


def bn_relu_train(x_nhwc, gamma, beta, eps):
    """Train-mode BatchNorm2d + ReLU on an NHWC bf16 tensor [..., C] (batch statistics, biased variance; reference
    PatchEmbed.proj[1:3] / [4:6] under model.train()).  Returns (y, mean, rstd, var_biased): y bf16 like x, the rest fp32 [C]."""
    dev = _chk(x_nhwc, gamma, beta)
    C = x_nhwc.shape[-1]
    npix = x_nhwc.numel() // C
    if x_nhwc.dtype != torch.bfloat16 or not x_nhwc.is_contiguous():
        raise PanoSwinB200Error("bn_relu_train wants a contiguous bf16 NHWC tensor")
    sums = torch.empty(2, C, dtype=torch.float32, device=x_nhwc.device)
    with torch.cuda.device(dev):
        _call("psw_bn_stats_fwd", _ptr(x_nhwc), _ptr(sums[0]), _ptr(sums[1]), npix, C, _stream(dev))
    mean = sums[0] / npix
    var = (sums[1] / npix - mean * mean).clamp_min_(0.0)
    rstd = torch.rsqrt(var + eps)
    scale = (_f32(gamma, "gamma") * rstd).contiguous()
    shift = (_f32(beta, "beta") - mean * scale).contiguous()
    y = torch.empty_like(x_nhwc)
    with torch.cuda.device(dev):
        _call("psw_bn_apply_relu_fwd", _ptr(x_nhwc), _ptr(y), _ptr(scale), _ptr(shift), npix, C, _stream(dev))
    return y, mean, rstd, var


def bn_relu_bwd(x_nhwc, y, dy, mean, rstd, gamma):
    """Gradients of bn_relu_train: (dx bf16 like x, dgamma, dbeta fp32 [C])."""
    dev = _chk(x_nhwc, y, dy, mean, rstd, gamma)
    C = x_nhwc.shape[-1]
    npix = x_nhwc.numel() // C
    dx = torch.empty_like(x_nhwc)
    dg = torch.empty(C, dtype=torch.float32, device=x_nhwc.device)
    db = torch.empty(C, dtype=torch.float32, device=x_nhwc.device)
    with torch.cuda.device(dev):
        _call("psw_bn_relu_bwd", _ptr(x_nhwc), _ptr(y), _ptr(dy), _ptr(mean), _ptr(rstd), _ptr(_f32(gamma, "gamma")), _ptr(dx),
              _ptr(dg), _ptr(db), npix, C, _stream(dev))
    return dx, dg, db


def transpose(x):
    """[R, C] -> [C, R] (contiguous)."""
    dev = _chk(x)
    R, Cc = x.shape
    out = torch.empty((Cc, R), dtype=x.dtype, device=x.device)
    with torch.cuda.device(dev):
        _call("psw_transpose", _ptr(x), _ptr(out), R, Cc, _dt(x), _stream(dev))
    return out


def window_attention_bwd(qkv, dout, alpha, beta, qkv_bias, uv, mask, heads, window, shift, pano_mode, scale):
    """Gradients of window_attention() / window_attention_full(): returns (dqkv [like qkv], dalpha, dbeta [fp32, like
    alpha], dqkv_bias [3C] fp32 or None: the gradient reaching the qkv bias through the padding cells)."""
    dev = _chk(qkv, dout, alpha, beta, qkv_bias, uv, mask)
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    if dout.dtype != qkv.dtype:
        raise PanoSwinB200Error("window_attention_bwd: qkv and dout must share a dtype")
    dqkv = torch.empty_like(qkv)
    dalpha = torch.zeros_like(alpha)
    dbeta = torch.empty_like(beta)
    dqb = torch.empty(C3, dtype=torch.float32, device=qkv.device) if qkv_bias is not None else None
    with torch.cuda.device(dev):
        _call("psw_window_attn_bwd", _ptr(qkv), _ptr(dout), _ptr(_f32(alpha, "alpha")), _ptr(_f32(beta, "beta")),
              _ptr(_f32(qkv_bias, "qkv_bias")), _ptr(_f32(uv, "uv")), _ptr(_f32(mask, "mask")), _ptr(dqkv), _ptr(dalpha), _ptr(dbeta),
              _ptr(dqb), B, H, W, C, heads, window, shift, 1 if pano_mode else 0, float(scale), _dt(qkv), _stream(dev))
    return dqkv, dalpha, dbeta, dqb
