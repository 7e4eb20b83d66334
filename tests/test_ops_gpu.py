"""Op-level parity of libpanoswin_b200 (through the C ABI, via ops.py) against the CPU oracle.
Tolerances: fp32 path <= 1e-5 rel-L2 (BASELINE.json north_star); bf16 path compared with an fp32
evaluation of the SAME bf16-rounded inputs: <= 4e-3 where only the output rounding differs (GEMM, LN),
<= 1e-2 for the attention core (bf16 P operand), the 'per-block' tolerance of SURVEY.md §8d."""
import pytest
import torch
import torch.nn.functional as F

from _expect import attention_core, rel_l2
from oracle import panoswin_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from panoswintransformerobjectdetection_b200 import ops as o
    return o


def _g(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.parametrize("C", [24, 96, 192, 384, 768, 1536, 3072])
@pytest.mark.parametrize("io", ["f32f32", "f32bf16", "bf16bf16"])
def test_layernorm(ops, C, io):
    g = _g(C)
    rows = 517
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    pos = torch.randn(47, C, generator=g)
    xin = x if io.startswith("f32") else x.bfloat16()
    odt = torch.float32 if io.endswith("f32") else torch.bfloat16
    want = F.layer_norm(xin.float(), (C,), gamma, beta, 1e-5)
    got = ops.layernorm(xin.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-5, odt)
    assert got.dtype == odt
    assert rel_l2(got.float(), want) <= (1e-6 if odt == torch.float32 else 4e-3)
    want_pos = want + pos[torch.arange(rows) % 47]
    got = ops.layernorm(xin.to(DEV), gamma.to(DEV), beta.to(DEV), 1e-5, odt, pos.to(DEV))
    assert rel_l2(got.float(), want_pos) <= (1e-6 if odt == torch.float32 else 4e-3)


@pytest.mark.parametrize("shape", [(2, 16, 32, 96), (1, 13, 25, 32), (2, 7, 13, 64), (1, 4, 7, 128), (1, 32, 64, 384)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_patch_merge_layernorm(ops, shape, dt):
    B, H, W, C = shape
    g = _g(H * W)
    x = torch.randn(B, H * W, C, generator=g).to(dt)
    gamma, beta = torch.randn(4 * C, generator=g), torch.randn(4 * C, generator=g)
    img = F.pad(x.float().view(B, H, W, C), (0, 0, 0, W % 2, 0, H % 2))
    quad = torch.cat([img[:, 0::2, 0::2], img[:, 1::2, 0::2], img[:, 0::2, 1::2], img[:, 1::2, 1::2]], -1)
    want = F.layer_norm(quad.reshape(B, -1, 4 * C), (4 * C,), gamma, beta, 1e-5)
    got = ops.patch_merge_layernorm(x.to(DEV), gamma.to(DEV), beta.to(DEV), H, W, 1e-5, dt)
    assert got.shape == want.shape
    assert rel_l2(got.float(), want) <= (1e-6 if dt == torch.float32 else 4e-3)


@pytest.mark.parametrize("shape", [(2, 16, 32, 96), (1, 13, 25, 192), (3, 5, 9, 768), (1, 2, 4, 1024),
                                   (2, 40, 70, 384), (2, 33, 67, 96),            # ragged tiles of the swizzled-tile kernel
                                   (2, 7, 9, 24), (1, 6, 11, 100)])              # C % 32 != 0: the scalar-tile kernel
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_nchw(ops, shape, dt):
    B, H, W, C = shape
    g = _g(C + H)
    x = torch.randn(B, H * W, C, generator=g).to(dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    want = F.layer_norm(x.float(), (C,), gamma, beta, 1e-5).view(B, H, W, C).permute(0, 3, 1, 2).contiguous()
    got = ops.layernorm_nchw(x.to(DEV), gamma.to(DEV), beta.to(DEV), H, W, 1e-5)
    assert got.dtype == torch.float32 and got.is_contiguous() and got.shape == want.shape
    assert rel_l2(got, want) <= 1e-6


@pytest.mark.parametrize("mnk", [(300, 96, 96), (1000, 288, 96), (64, 32, 5), (513, 72, 24), (257, 96, 384), (129, 384, 1536)])
@pytest.mark.parametrize("gelu,res", [(False, False), (True, False), (False, True)])
def test_linear_fp32(ops, mnk, gelu, res):
    M, N, K = mnk
    g = _g(M + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g) if res else None
    want = F.linear(x.double(), w.double(), b.double())
    if gelu:
        want = F.gelu(want)
    if res:
        want = want + r.double()
    got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), None if r is None else r.to(DEV), gelu)
    assert rel_l2(got, want) <= 2e-6


TC_SHAPES = [(300, 96, 96), (1000, 288, 96), (4096, 384, 96), (777, 96, 384), (512, 2304, 768), (256, 768, 3072),
             (130, 192, 1536), (128, 1152, 384), (20000, 192, 192), (333, 128, 128), (100, 32, 32), (65, 576, 192),
             (200, 48, 64), (130, 144, 96), (257, 400, 128), (40000, 96, 96)]


@pytest.mark.parametrize("mnk", TC_SHAPES)
@pytest.mark.parametrize("odt,gelu,res,bias", [(torch.bfloat16, False, False, True), (torch.bfloat16, True, False, True),
                                               (torch.float32, False, True, True), (torch.float32, False, False, False),
                                               (torch.bfloat16, False, True, True)])
def test_linear_bf16_tcgen05(ops, mnk, odt, gelu, res, bias):
    M, N, K = mnk
    g = _g(M * 3 + N)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g) if bias else None
    r = torch.randn(M, N, generator=g).to(odt) if res else None
    want = F.linear(x.double(), w.double(), None if b is None else b.double())
    if gelu:
        want = F.gelu(want)
    if res:
        want = want + r.double()
    got = ops.linear(x.to(DEV), w.to(DEV), None if b is None else b.to(DEV), None if r is None else r.to(DEV), gelu,
                     out_dtype=odt)
    torch.cuda.synchronize()
    assert got.dtype == odt and got.shape == (M, N)
    assert torch.isfinite(got).all()
    assert rel_l2(got.float(), want) <= (4e-3 if odt == torch.bfloat16 else 2e-5)


PAIR_SHAPES = [(333, 96, 64), (512, 256, 128), (1000, 384, 1536), (4100, 768, 768), (20000, 1152, 384), (257, 64, 3072)]


@pytest.mark.parametrize("mnk", PAIR_SHAPES)
@pytest.mark.parametrize("odt,gelu,res", [(torch.bfloat16, False, False), (torch.bfloat16, True, False), (torch.float32, False, True)])
@pytest.mark.parametrize("mode", [1 << 27, (1 << 27) | (1 << 25), 1 << 26])
def test_linear_bf16_cta_pair(ops, monkeypatch, mnk, odt, gelu, res, mode):
    """The cta_group::2 (CTA-pair, 256-row tiles) variant of the tcgen05 GEMM forced on ragged / small / deep-K
    shapes, with 16 and 8 epilogue warps, against the forced single-CTA variant and the fp64 reference."""
    from panoswintransformerobjectdetection_b200 import _lib
    monkeypatch.setattr(_lib, "_DEFAULT_DIAG", True)        # route this test's calls to the -DPSW_DIAGNOSTICS build:
    lib = _lib.load()                                       # the forcing switch does not exist in the product library
    M, N, K = mnk
    g = _g(M + 7 * N + K)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g).to(odt) if res else None
    want = F.linear(x.double(), w.double(), b.double())
    if gelu:
        want = F.gelu(want)
    if res:
        want = want + r.double()
    old = lib.psw_diag_linear_mode(mode)
    try:
        got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), None if r is None else r.to(DEV), gelu, out_dtype=odt)
        torch.cuda.synchronize()
    finally:
        lib.psw_diag_linear_mode(old)
    assert torch.isfinite(got).all()
    assert rel_l2(got.float(), want) <= (4e-3 if odt == torch.bfloat16 else 2e-5)


def test_linear_bf16_gelu_extremes(ops):
    """The packed-fp16 GELU epilogue stays finite and exact in the tails (|x| far beyond the fp16 range)."""
    M, N, K = 256, 64, 64
    x = torch.zeros(M, K).bfloat16()
    w = torch.zeros(N, K).bfloat16()
    b = torch.linspace(-3e5, 3e5, N)
    got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), None, True, out_dtype=torch.bfloat16).float().cpu()
    want = F.gelu(b.double()).float().bfloat16().float().expand(M, N)
    assert torch.isfinite(got).all()
    big = b.abs() < 6.0e4                                   # inside the fp16 range: exact tails
    assert torch.equal(got[:, big], want[:, big])
    assert (got[:, b < -6.0e4] == 0).all() and (got[:, b > 6.0e4] >= 6.0e4).all()


def _attn_case(H, W, heads, hd, shift, pano, B=2, seed=0, ws=7):
    g = _g(seed + H * 7 + W)
    C = heads * hd
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    qkv[..., :C] *= 1.5
    alpha = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    beta = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    qb = torch.randn(3 * C, generator=g) * 0.5
    uv = O.uv_grid(H, W) if pano else torch.zeros(H, W, 2)
    return qkv, alpha, beta, qb, uv, C


ATTN_CASES = [  # H, W, heads, shift, pano
    (16, 32, 2, 0, True), (16, 32, 2, 3, True), (13, 25, 3, 3, True), (13, 25, 1, 0, True), (7, 13, 4, 3, True),
    (4, 7, 8, 3, True), (25, 50, 1, 3, True), (12, 31, 2, 3, False), (12, 31, 2, 0, False), (20, 16, 3, 3, False),
]


@pytest.mark.parametrize("case", ATTN_CASES)
@pytest.mark.parametrize("hd", [32, 24])
def test_window_attention_fp32(ops, case, hd):
    H, W, heads, shift, pano = case
    qkv, alpha, beta, qb, uv, C = _attn_case(H, W, heads, hd, shift, pano)
    scale = hd ** -0.5
    want = attention_core(qkv, alpha, beta, qb, uv, H, W, heads, 7, shift, pano, scale)
    mask = O.planar_shift_mask(H, W, 7, shift).to(DEV) if (not pano and shift) else None
    got = ops.window_attention(qkv.to(DEV), alpha.to(DEV), beta.to(DEV), qb.to(DEV), uv.to(DEV) if pano else None, mask,
                               heads, 7, shift, pano, scale)
    torch.cuda.synchronize()
    assert rel_l2(got, want) <= 1e-5


@pytest.mark.parametrize("case", ATTN_CASES)
@pytest.mark.parametrize("impl", ["generic", "full"])
def test_window_attention_bf16(ops, case, impl):
    """bf16 storage: the tcgen05 kernel with the precomputed bias table ("full", window-pair schedule at this batch)
    and the generic CUDA-core route of psw_window_attn_fwd(PSW_BF16) against the oracle primitives."""
    H, W, heads, shift, pano = case
    qkv, alpha, beta, qb, uv, C = _attn_case(H, W, heads, 32, shift, pano, seed=5)
    qkv = qkv.bfloat16()
    scale = 32 ** -0.5
    # padding tokens use the bf16-rounded bias on the tensor-core path; fp32 on the CUDA-core path
    want = attention_core(qkv.float(), alpha, beta, qb.bfloat16().float() if impl == "full" else qb, uv, H, W, heads, 7,
                          shift, pano, scale)
    mask = O.planar_shift_mask(H, W, 7, shift).to(DEV) if (not pano and shift) else None
    if impl == "full":                                           # production path: all additive terms precomputed (fp32)
        bf = ops.window_bias_full(alpha.to(DEV), beta.to(DEV), uv.to(DEV) if pano else None, mask, H, W, 7, shift, pano)
        got = ops.window_attention_full(qkv.to(DEV), bf, qb.to(DEV), heads, 7, shift, pano, scale)
    else:
        got = ops.window_attention(qkv.to(DEV), alpha.to(DEV), beta.to(DEV), qb.to(DEV), uv.to(DEV) if pano else None, mask,
                                   heads, 7, shift, pano, scale)
    torch.cuda.synchronize()
    assert got.dtype == torch.bfloat16 and torch.isfinite(got.float()).all()
    assert rel_l2(got.float(), want) <= 1e-2


@pytest.mark.parametrize("ws,hd,H,W,heads,shift,pano", [(12, 32, 24, 48, 4, 6, True), (12, 32, 24, 48, 2, 0, True),
                                                        (5, 16, 13, 25, 3, 2, True), (8, 64, 16, 32, 2, 4, False),
                                                        (12, 32, 30, 40, 2, 6, False)])
@pytest.mark.parametrize("dt", ["fp32", "bf16"])
def test_window_attention_other_windows_and_head_dims(ops, ws, hd, H, W, heads, shift, pano, dt):
    """BASELINE.json configs[3] 'larger windows/heads': window 12 (Swin-B/384 style, 144 tokens per window) and other
    head dims run on the generic route of psw_window_attn_fwd in both storage types."""
    qkv, alpha, beta, qb, uv, C = _attn_case(H, W, heads, hd, shift, pano, seed=21, ws=ws)
    scale = hd ** -0.5
    if dt == "bf16":
        qkv = qkv.bfloat16()
    want = attention_core(qkv.float(), alpha, beta, qb, uv, H, W, heads, ws, shift, pano, scale)
    mask = O.planar_shift_mask(H, W, ws, shift).to(DEV) if (not pano and shift) else None
    assert not ops.window_attention_full_supported(ws, hd)
    got = ops.window_attention(qkv.to(DEV), alpha.to(DEV), beta.to(DEV), qb.to(DEV), uv.to(DEV) if pano else None, mask,
                               heads, ws, shift, pano, scale)
    torch.cuda.synchronize()
    assert rel_l2(got.float(), want) <= (1e-5 if dt == "fp32" else 1e-2)


@pytest.mark.parametrize("case", ATTN_CASES)
@pytest.mark.parametrize("B", [4, 5, 9])
def test_window_attention_bf16_batch_inner(ops, case, B):
    """Batches of >= 4 images take the batch-innermost kernel (same window position of two images per tile, bias row
    and token map kept across the image pairs of an item); odd batches leave half of the last tile empty."""
    H, W, heads, shift, pano = case
    qkv, alpha, beta, qb, uv, C = _attn_case(H, W, heads, 32, shift, pano, B=B, seed=11)
    qkv = qkv.bfloat16()
    scale = 32 ** -0.5
    want = attention_core(qkv.float(), alpha, beta, qb.bfloat16().float(), uv, H, W, heads, 7, shift, pano, scale)
    mask = O.planar_shift_mask(H, W, 7, shift).to(DEV) if (not pano and shift) else None
    bf = ops.window_bias_full(alpha.to(DEV), beta.to(DEV), uv.to(DEV) if pano else None, mask, H, W, 7, shift, pano)
    got = ops.window_attention_full(qkv.to(DEV), bf, qb.to(DEV), heads, 7, shift, pano, scale)
    torch.cuda.synchronize()
    assert torch.isfinite(got.float()).all()
    assert rel_l2(got.float(), want) <= 1e-2
    for b in range(B):                                           # per-image check: no cross-talk between the tile halves
        assert rel_l2(got[b].float(), want[b]) <= 1.5e-2


def test_window_attention_no_qkv_bias(ops):
    H, W, heads, shift, pano = 13, 25, 2, 3, True
    qkv, alpha, beta, _, uv, C = _attn_case(H, W, heads, 32, shift, pano, seed=9)
    want = attention_core(qkv, alpha, beta, None, uv, H, W, heads, 7, shift, pano, 32 ** -0.5)
    got = ops.window_attention(qkv.to(DEV), alpha.to(DEV), beta.to(DEV), None, uv.to(DEV), None, heads, 7, shift, pano, 32 ** -0.5)
    assert rel_l2(got, want) <= 1e-5
    got = ops.window_attention(qkv.bfloat16().to(DEV), alpha.to(DEV), beta.to(DEV), None, uv.to(DEV), None, heads, 7, shift,
                               pano, 32 ** -0.5)
    want = attention_core(qkv.bfloat16().float(), alpha, beta, None, uv, H, W, heads, 7, shift, pano, 32 ** -0.5)
    assert rel_l2(got.float(), want) <= 1e-2


def test_errors_are_loud(ops):
    from panoswintransformerobjectdetection_b200.ops import PanoSwinB200Error
    x = torch.randn(4, 96)
    with pytest.raises(PanoSwinB200Error):
        ops.layernorm(x, torch.ones(96), torch.zeros(96))                      # CPU tensors: no fallback
    qkv = torch.randn(1, 14, 28, 3 * 48, device=DEV).bfloat16()               # head_dim 48 on the tcgen05 path
    t = torch.zeros(169, 1, device=DEV)
    with pytest.raises(PanoSwinB200Error):
        ops.window_attention_full(qkv, torch.zeros(8, 1, 13, 64, 4, device=DEV), None, 1, 7, 0, True, 1.0)
    with pytest.raises(PanoSwinB200Error):                                     # pano mode without the uv table
        ops.window_attention(qkv, t, t, None, None, None, 1, 7, 0, True, 1.0)
    with pytest.raises(PanoSwinB200Error):
        ops.window_attention(qkv.float(), t, t, None, torch.zeros(14, 28, 2, device=DEV), None, 1, 7, 7, True, 1.0)   # shift >= window


@pytest.mark.parametrize("cout", [32, 64])
@pytest.mark.parametrize("shape", [(2, 3, 64, 128), (1, 3, 52, 100), (1, 3, 9, 68), (3, 3, 16, 64), (1, 3, 20, 260)])
def test_stem_conv3x3_relu(ops, shape, cout):
    """tcgen05 stem conv (3 -> 32 / 64) + folded BN + ReLU against fp32 conv on the bf16-rounded operands."""
    B, _, H, W = shape
    g = _g(H + W + cout)
    img = torch.rand(shape, generator=g)
    w = torch.randn(cout, 3, 3, 3, generator=g) / 27 ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    wq = w.bfloat16().float()
    want = F.relu(F.conv2d(img.bfloat16().float(), wq, b, padding=1)).permute(0, 2, 3, 1)
    got = ops.stem_conv3x3_relu(img.to(DEV), w.reshape(cout, 27).to(DEV), b.to(DEV))
    torch.cuda.synchronize()
    assert got.shape == (B, H, W, cout) and got.dtype == torch.bfloat16
    assert rel_l2(got.float(), want) <= 4e-3


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 8, 128, 64, 96), (2, 13, 100, 64, 96), (1, 4, 260, 64, 64), (3, 33, 129, 64, 48),
                                             (2, 64, 256, 64, 96), (1, 5, 40, 128, 128)])
def test_conv3x3_nhwc_gemm_view(ops, B, H, W, cin, cout, relu):
    """conv3x3 (stride 1, pad 1) + bias (+ ReLU) of an NHWC bf16 image through the tcgen05 GEMM over shifted 4-D TMA views
    (PanoSwin-B's second stem layer, channels zero-padded to 64 -> 96) vs torch fp32; ragged heights / widths and the image
    borders rely on TMA's out-of-range zero fill."""
    g = _g(B * 1000 + H * 10 + W + cout)
    x = torch.randn(B, H, W, cin, generator=g).bfloat16()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5).bfloat16()
    b = torch.randn(cout, generator=g)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    if relu:
        want = F.relu(want)
    got = ops.conv3x3_nhwc(x.to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), relu=relu)
    torch.cuda.synchronize()
    assert got.shape == (B, H, W, cout) and got.dtype == torch.bfloat16
    assert rel_l2(got.float(), want) <= 4e-3
    assert (got.float().cpu() - want).abs().max() <= 0.06


def test_stem_conv_rejects_unsupported_shapes(ops):
    from panoswintransformerobjectdetection_b200.ops import PanoSwinB200Error
    w, b = torch.zeros(32, 27, device=DEV), torch.zeros(32, device=DEV)
    with pytest.raises(PanoSwinB200Error):                      # W must be a multiple of 4
        ops.stem_conv3x3_relu(torch.zeros(1, 3, 8, 70, device=DEV), w, b)
    with pytest.raises(PanoSwinB200Error):                      # built for 3 -> 32 or 64 channels
        ops.stem_conv3x3_relu(torch.zeros(1, 3, 8, 64, device=DEV), torch.zeros(48, 27, device=DEV), torch.zeros(48, device=DEV))


@pytest.mark.parametrize("cout", [32, 64])
@pytest.mark.parametrize("B,H,W", [(1, 8, 128), (2, 13, 100), (1, 4, 260), (3, 33, 129), (2, 64, 256)])
def test_stem_conv2_tcgen05(ops, B, H, W, cout):
    """conv3x3 32->32 + folded BN + ReLU on NHWC bf16 (implicit GEMM over shifted patch views) vs torch fp32,
    including ragged heights / widths (zero padding comes from TMA out-of-bounds fill)."""
    g = _g(B * 1000 + H * 10 + W)
    x = torch.randn(B, H, W, 32, generator=g).bfloat16()
    w = (torch.randn(cout, 32, 3, 3, generator=g) / 17.0).bfloat16()
    b = torch.randn(cout, generator=g)
    want = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)).permute(0, 2, 3, 1)
    w_taps = w.permute(2, 3, 0, 1).reshape(9, cout, 32).contiguous()
    got = ops.stem_conv3x3_c32_relu(x.to(DEV), w_taps.to(DEV), b.to(DEV))
    torch.cuda.synchronize()
    assert got.shape == (B, H, W, cout) and got.dtype == torch.bfloat16
    assert rel_l2(got.float(), want) <= 4e-3
    assert (got.float().cpu() - want).abs().max() <= 0.05


@pytest.mark.parametrize("B,H,W,cin,cout,patch", [(2, 16, 512, 64, 96, (4, 4)), (1, 8, 1024, 64, 96, (4, 4)), (3, 12, 100, 64, 96, (4, 4)),
                                                   (2, 8, 64, 32, 48, (2, 2)), (1, 4, 2048, 64, 128, (4, 4))])
def test_patch_conv_as_gemm(ops, B, H, W, cin, cout, patch):
    """conv(kernel = stride = patch) through the tcgen05 GEMM over a 3-D TMA view of the NHWC image (token rows that are
    not a multiple of the 128-row tile included) vs torch fp32."""
    g = _g(B + H + W + cout)
    x = torch.randn(B, H, W, cin, generator=g).bfloat16()
    w = (torch.randn(cout, cin, patch[0], patch[1], generator=g) / (cin * patch[0] * patch[1]) ** 0.5).bfloat16()
    b = torch.randn(cout, generator=g)
    want = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=patch).permute(0, 2, 3, 1)
    got = ops.patch_conv(x.to(DEV), w.permute(0, 2, 3, 1).contiguous().to(DEV), b.to(DEV), patch)
    torch.cuda.synchronize()
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    assert rel_l2(got.float(), want) <= 4e-3


@pytest.mark.parametrize("mnk", [(300, 96, 96), (4096, 96, 384), (1000, 192, 192), (777, 192, 768), (130, 256, 64), (20000, 96, 96)])
@pytest.mark.parametrize("alias", [False, True])
def test_linear_layernorm_fused(ops, mnk, alias):
    """proj / fc2 with the following LayerNorm fused into the GEMM epilogue: y = x w^T + b + residual (fp32) and
    LN(y) (bf16), against fp64; `alias` writes y over the residual tensor as the backbone does."""
    M, N, K = mnk
    g = _g(M + N + K)
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g) * 3 + 1.5                   # a mean offset exercises the two-pass variance
    gam, bet = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    want_y = F.linear(x.double(), w.double(), b.double()) + r.double()
    want_ln = F.layer_norm(want_y, (N,), gam.double(), bet.double(), 1e-5)
    rd = r.to(DEV)
    y, ln = ops.linear_layernorm(x.to(DEV), w.to(DEV), b.to(DEV), rd, gam.to(DEV), bet.to(DEV), 1e-5, out=rd if alias else None)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and ln.dtype == torch.bfloat16 and ln.shape == (M, N)
    assert (y.data_ptr() == rd.data_ptr()) == alias
    assert rel_l2(y, want_y) <= 2e-5
    assert rel_l2(ln.float(), want_ln) <= 4e-3


@pytest.mark.parametrize("B,H,W,N,K", [(2, 8, 16, 96, 384), (3, 5, 9, 192, 768), (1, 16, 32, 96, 96), (4, 7, 11, 256, 64)])
def test_linear_layernorm_nchw_fused(ops, B, H, W, N, K):
    """Last fc2 of a stage with the stage's output LayerNorm -> fp32 NCHW fused into the epilogue (token counts that are
    not multiples of the 128-row tile, so tiles straddle images)."""
    g = _g(B + H + W + N + K)
    M = B * H * W
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g)
    r = torch.randn(M, N, generator=g) * 2 - 0.7
    gam, bet = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    want_y = F.linear(x.double(), w.double(), b.double()) + r.double()
    want_map = F.layer_norm(want_y, (N,), gam.double(), bet.double(), 1e-5).view(B, H, W, N).permute(0, 3, 1, 2)
    rd = r.to(DEV).view(B, H * W, N)
    y, fmap = ops.linear_layernorm_nchw(x.to(DEV).view(B, H * W, K), w.to(DEV), b.to(DEV), rd, gam.to(DEV), bet.to(DEV), 1e-5,
                                        H, W, out=rd)
    torch.cuda.synchronize()
    assert fmap.shape == (B, N, H, W) and fmap.dtype == torch.float32 and fmap.is_contiguous()
    assert rel_l2(y.view(M, N), want_y) <= 2e-5
    assert rel_l2(fmap, want_map) <= 2e-5


@pytest.mark.parametrize("M", [128, 300, 4096, 20000, 148 * 128 * 3 + 77])
def test_mlp_fused(ops, M):
    """fc1 + GELU + fc2 + shortcut in one kernel (C = 96, hidden = 384) against fp64 with the bf16-rounded hidden."""
    _check_mlp_fused(ops, M)


def _check_mlp_fused(ops, M):
    g = _g(M)
    C, Hd = 96, 384
    xn = torch.randn(M, C, generator=g).bfloat16()
    w1 = (torch.randn(Hd, C, generator=g) / C ** 0.5).bfloat16()
    w2 = (torch.randn(C, Hd, generator=g) / Hd ** 0.5).bfloat16()
    b1, b2 = torch.randn(Hd, generator=g) * 0.5, torch.randn(C, generator=g) * 0.5
    x = torch.randn(M, C, generator=g) * 2
    hid = F.gelu(F.linear(xn.double(), w1.double(), b1.double()))
    want = x.double() + F.linear(hid, w2.double(), b2.double())
    got = ops.mlp_fused(xn.to(DEV), w1.to(DEV), b1.to(DEV), w2.to(DEV), b2.to(DEV), x.to(DEV).clone())
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert rel_l2(got, want) <= 3e-3                          # the hidden activation is rounded to bf16 once
    assert rel_l2(got - x.to(DEV), want - x.double()) <= 6e-3


@pytest.mark.parametrize("rows,C", [(1000, 96), (333, 192), (4096, 128), (77, 768)])
@pytest.mark.parametrize("with_pos", [True, False])
def test_layernorm2_dual(ops, rows, C, with_pos):
    """Stem patch_norm (+ position add) and the first block's norm1 in one pass over the rows."""
    g = _g(rows + C)
    x = (torch.randn(rows, C, generator=g) * 1.7 + 0.3).bfloat16()
    g1, b1 = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    g2, b2 = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    pos = torch.randn(rows // 3 if rows % 3 == 0 else rows, C, generator=g) if with_pos else None
    want = F.layer_norm(x.double(), (C,), g1.double(), b1.double(), 1e-5)
    if with_pos:
        want = want + pos.double().repeat(rows // pos.shape[0], 1)
    want2 = F.layer_norm(want, (C,), g2.double(), b2.double(), 1e-5)
    y, y2 = ops.layernorm2(x.to(DEV), g1.to(DEV), b1.to(DEV), 1e-5, None if pos is None else pos.to(DEV), g2.to(DEV), b2.to(DEV), 1e-5)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and y2.dtype == torch.bfloat16
    assert rel_l2(y, want) <= 2e-6
    assert rel_l2(y2.float(), want2) <= 4e-3


@pytest.mark.parametrize("B,cin,H,W,cout,k,s,p", [(2, 3, 20, 36, 32, 3, 1, 1), (1, 32, 17, 33, 64, 3, 1, 1), (2, 64, 16, 24, 96, 4, 4, 0),
                                                  (1, 42, 9, 11, 84, 3, 1, 1), (1, 5, 8, 8, 7, 2, 2, 0)])
@pytest.mark.parametrize("nhwc", [False, True])
def test_conv2d_f32_parity_path(ops, B, cin, H, W, cout, k, s, p, nhwc):
    """The fp32 stem convolution (CUDA-core direct conv + BatchNorm affine + ReLU) against torch fp64."""
    g = _g(B + cin + H + W + cout)
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    sc, sh = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    want = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=s, padding=p) * sc.double()[None, :, None, None] +
                  sh.double()[None, :, None, None])
    got = ops.conv2d_f32(x.to(DEV), w.to(DEV), b.to(DEV), sc.to(DEV), sh.to(DEV), s, p, relu=True, out_nhwc=nhwc)
    torch.cuda.synchronize()
    if nhwc:
        got = got.permute(0, 3, 1, 2)
    assert got.shape == want.shape and rel_l2(got, want) <= 2e-6
    plain = ops.conv2d_f32(x.to(DEV), w.to(DEV), None, None, None, s, p)
    assert rel_l2(plain, F.conv2d(x.double(), w.double(), None, stride=s, padding=p)) <= 2e-6
