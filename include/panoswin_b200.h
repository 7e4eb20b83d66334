/*
 * panoswin_b200.h — C ABI of libpanoswin_b200.so: the B200 (sm_100a) kernels behind the PanoSwin
 * pano-style shifted-window attention path.
 *
 * The reference (1069066484/PanoSwinTransformerObjectDetection) has NO FFI: the path is pure Python
 * (mmdet/models/backbones/simple_panoswin_transformer.py).  Each entry point below therefore cites the
 * reference Python code it replaces; INTEGRATION.md shows the ctypes stub a reference maintainer adds.
 *
 * Conventions (every function):
 *   - plain pointers and sizes only; all pointers are DEVICE pointers owned by the caller;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work: no allocation, no sync;
 *   - returns 0 on success, a NEGATIVE psw_status on argument errors, a POSITIVE cudaError_t when the
 *     CUDA runtime refused the launch; psw_last_error_string() describes the last failure of the
 *     calling thread;
 *   - stateless and thread-safe (kernel selection depends on the arguments only; the profiling switches of
 *     include/panoswin_b200_debug.h exist only in a -DPSW_DIAGNOSTICS build of the library);
 *   - there is NO CPU fallback: without a B200 the launch fails loudly.
 *   - dtype arguments: PSW_F32 selects the fp32 parity path (CUDA-core FMA, 1e-5 vs the reference),
 *     PSW_BF16 the throughput path (bf16 storage, tcgen05 tensor cores, fp32 accumulate/softmax).
 */
#ifndef PANOSWIN_B200_H_
#define PANOSWIN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSW_ABI_VERSION 6   /* bumped with every change of a prototype below */

#if defined(__GNUC__)
#define PSW_API __attribute__((visibility("default")))
#else
#define PSW_API
#endif

enum psw_dtype { PSW_F32 = 0, PSW_BF16 = 1 };

enum psw_status {
  PSW_OK = 0,
  PSW_ERR_BAD_ARG = -1,      /* null pointer, non-positive size, misaligned pointer            */
  PSW_ERR_UNSUPPORTED = -2,  /* shape outside what the sm_100a kernel was built for            */
  PSW_ERR_NO_DEVICE = -3,    /* no CUDA device / not compute capability 10.x                   */
  PSW_ERR_DRIVER = -4        /* driver entry point (tensor-map encode) unavailable             */
};

/* epilogue flags of psw_linear_fwd */
#define PSW_EPI_GELU 1       /* GELU after the bias add (nn.GELU, reference :51,:57).  PSW_F32: exact erff.
                              * PSW_BF16: 0.5 x (1 + tanh(x p(x^2))) with a fitted cubic p, evaluated in packed fp16;
                              * max abs deviation from the erf GELU 5e-5, i.e. far below the bf16 output rounding */

PSW_API int psw_abi_version(void);
PSW_API const char* psw_last_error_string(void);
/* Returns 0 when device `dev` is an sm_100 (B200) part the kernels can run on. */
PSW_API int psw_check_device(int dev);

/*
 * LayerNorm over the last dimension: y[r,:] = (x[r,:] - mean) * rstd * gamma + beta.
 * Replaces norm1 / norm2 of PanoSwinTransformerBlock.forward
 * (simple_panoswin_transformer.py:504, :534) and patch_embed.norm (:768-772).
 * x [rows, C] (in_dtype), y [rows, C] (out_dtype), gamma/beta [C] fp32.  C % 4 == 0, C <= 4096.
 * `pos` (nullable, fp32 [pos_rows, C]) is added after the affine, row r using pos[r % pos_rows]: the
 * pano absolute position embedding of forward() (:960-962).
 */
PSW_API int psw_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, const float* pos,
                      int64_t rows, int C, int64_t pos_rows, float eps, int in_dtype, int out_dtype, void* stream);

/*
 * Two LayerNorms in one pass over the row: y = LN(x) * gamma + beta (+ pos), fp32, and y2 = LN(y) * gamma2 + beta2, bf16.
 * In the backbone: the stem's patch_norm + absolute position add (the residual stream) followed by the first block's
 * norm1 (reference :768-771, :925-936, :506).  x may be fp32 or bf16.
 */
PSW_API int psw_layernorm2_fwd(const void* x, float* y, const float* gamma, const float* beta, const float* pos,
                               void* y2, const float* gamma2, const float* beta2, int64_t rows, int C,
                               int64_t pos_rows, float eps, float eps2, int in_dtype, void* stream);

/*
 * y[M,N] = act(x[M,K] . w[N,K]^T + bias[N]) (+ residual[M,N]).
 * Replaces nn.Linear of attn.qkv (:287), attn.proj (:309) with the block's first residual add (:533),
 * Mlp fc1+GELU / fc2 (:55-61) with the second residual add (:534), and PatchMerging.reduction (:575).
 * PSW_F32 : x, w, residual, y fp32 (CUDA-core FMA).
 * PSW_BF16: x, w bf16; bias fp32 (nullable); accumulate fp32 on tcgen05/TMEM;
 *           y is `out_dtype`; residual (nullable) has dtype `out_dtype`.  K % 32 == 0, N % 16 == 0.
 * `flags`: PSW_EPI_GELU.
 */
PSW_API int psw_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                   int64_t M, int N, int K, int flags, int dtype, int out_dtype, void* stream);

/*
 * psw_linear_fwd (bf16 operands, fp32 output + fp32 residual) with the LayerNorm of the result fused into the epilogue:
 * y = x . w^T + bias + residual (fp32 [M, N]; y may alias residual) and ln_out = LayerNorm(y) * ln_gamma + ln_beta
 * (bf16 [M, N]).  In a PanoSwin block: attn.proj + shortcut -> norm2, and mlp.fc2 + shortcut -> the next block's norm1
 * (reference :520-535).  One tile must hold complete rows: N % 32 == 0 and N <= 256.
 */
PSW_API int psw_linear_ln_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                              const float* ln_gamma, const float* ln_beta, float ln_eps, void* ln_out,
                              int64_t M, int N, int K, void* stream);
/* Same, for the last fc2 of a stage: LayerNorm(y) is the stage's output map, written as fp32 NCHW [M / HW, N, HW]
 * (SimplePanoSwinTransformer.forward :974-978) straight from the epilogue (HW = tokens per image, M % HW == 0). */
PSW_API int psw_linear_ln_nchw_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                   const float* ln_gamma, const float* ln_beta, float ln_eps, float* out_nchw,
                                   int64_t HW, int64_t M, int N, int K, void* stream);

/*
 * Fused (shifted-)window multi-head self-attention on an equirectangular token map.
 * Replaces, in ONE pass over HBM, WindowTransition.forward (:376-409, pano shift with longitude
 * wrap-around), pad_x (:486-491), window_partition (:64-75), the core of
 * BasicWindowAttention.forward (:290-308: q*scale, q.k^T, great-circle bias of _sphere_bias :241-260
 * with haversine22 of lzx/models/great_circle.py:71-86, optional planar shift mask, softmax, P.v),
 * window_reverse (:78-92), the crop (:516) and the reverse transition (:394-397).
 *
 *   qkv      [B, H, W, 3C]  output of the qkv linear on UN-shifted tokens, channel order (3, heads, hd)
 *   out      [B, H, W, C]   attention output (before proj), un-shifted token order
 *   alpha, beta [(2*window-1)^2, heads] fp32 tables (sphere_position_{alpha,beta}_table_Te)
 *   qkv_bias [3C] fp32 or NULL — q/k/v of a zero (padding) token: padded cells take part as keys/values
 *   uv       [H, W, 2] fp32 token coordinates (make_uv_hw2 :153-189); required in pano mode
 *   mask     [nW, window^2, window^2] fp32 additive mask or NULL (planar mode, shifted blocks only)
 *   pano_mode 1: attention runs on the (2H, ceil(W/2)) north-south layout; 0: planar Swin roll(-s,-s)
 *
 * This entry is the GENERIC route: CUDA-core kernel, one CTA per (window, head), ANY window size and head_dim
 * (shared memory permitting: window 12 x head_dim 32 fits).  dtype = PSW_F32 is the parity path (<= 1e-5 of the
 * reference); dtype = PSW_BF16 (bf16 qkv / out, fp32 math) serves the configurations the tcgen05 kernel below is not
 * instantiated for (psw_window_attn_full_supported() == 0), e.g. window 12 of Swin-B/384-style models.
 */
PSW_API int psw_window_attn_fwd(const void* qkv, void* out, const float* alpha, const float* beta,
                                const float* qkv_bias, const float* uv, const float* mask,
                                int B, int H, int W, int C, int heads, int window, int shift, int pano_mode,
                                float scale, int dtype, void* stream);

/*
 * Windows per column / row of the shifted, padded map the attention runs on (pano: ceil(2H/window) x
 * ceil(ceil(W/2)/window); planar: ceil(H/window) x ceil(W/window)).  Host-only helper.
 */
PSW_API int psw_window_grid(int H, int W, int window, int pano_mode, int* nwh, int* nww);

/*
 * Host-only: the gather map of WindowTransition.forward + pad_x as every kernel of this library evaluates it
 * (closed form, psw::source_token): map[i * wp + j] = flat source token h*W + w of cell (i, j) of the shifted, padded
 * map, or -1 for a zero-padding cell; *hp / *wp = padded height / width.  map may be NULL to query the size.
 */
PSW_API int psw_window_source_map(int H, int W, int window, int shift, int pano_mode, int* map, int capacity,
                                  int* hp, int* wp);

/*
 * The whole MLP of a block in one kernel (Mlp.forward + shortcut, reference :44-61, :534):
 * x <- x + fc2(GELU(fc1(xn))), xn [M, C] bf16 (= norm2(x)), x [M, C] fp32 updated in place; w1 [hidden, C], w2 [C, hidden]
 * bf16, b1 [hidden], b2 [C] fp32 (both required).  The hidden activation never leaves the SM: both weight matrices stay in
 * shared memory, fc1 chunks accumulate in tensor memory, GELU writes bf16 back to tensor memory, fc2 consumes it from
 * there.  Instantiated for C = 96, hidden = 384 (stage 0 of the embed_dim-96 models); other widths return
 * PSW_ERR_UNSUPPORTED and the caller uses two psw_linear_fwd calls.
 */
PSW_API int psw_mlp_fused_fwd(const void* xn, const void* w1, const float* b1, const void* w2, const float* b2,
                              void* x, int64_t M, int C, int hidden, void* stream);

/*
 * Production bf16 path (tcgen05 / TMEM).  psw_window_bias_full() evaluates EVERY additive term of the attention logits
 * of one block at one resolution — hav(uv_i, uv_j) * alpha[idx] + beta[idx] (_sphere_bias, reference :241-272, fp32
 * math) and, in planar mode, the shifted-window mask (:621-643; mask may be NULL) — for every window of ONE image and
 * every head, multiplied by log2(e) (the kernel's softmax uses exp2): table fp32 [windows][heads][13][64][4], element
 * (i, j) at chunk j/4, row i, lane j%4; psw_window_bias_full_bytes() is its size.  It depends on the geometry and the
 * block's alpha / beta only (not on the batch): build it once per block and resolution; the kernel reads it with 13
 * coalesced 16-byte loads per row (it stays L2-resident across the images of a batch).
 * psw_window_attn_full_fwd() is psw_window_attn_fwd(PSW_BF16) taking that table instead of alpha / beta / uv / mask.
 * Instantiated for window 7 and head_dim 32 (every shipped PanoSwin config): psw_window_attn_full_supported() tells;
 * other shapes return PSW_ERR_UNSUPPORTED and belong on psw_window_attn_fwd.
 */
PSW_API int psw_window_attn_full_supported(int window, int head_dim);
PSW_API int64_t psw_window_bias_full_bytes(int H, int W, int heads, int window, int pano_mode);
PSW_API int psw_window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask, void* table,
                                 int H, int W, int heads, int window, int shift, int pano_mode, void* stream);
PSW_API int psw_window_attn_full_fwd(const void* qkv, void* out, const void* bias_full, const float* qkv_bias,
                                     int B, int H, int W, int C, int heads, int window, int shift,
                                     int pano_mode, float scale, void* stream);

/*
 * PatchMerging front half: 2x2 gather in the order (0,0),(1,0),(0,1),(1,1) with zero padding of odd
 * H/W, then LayerNorm(4C) (:563-574).  x [B, H, W, C] (in_dtype) -> y [B, ceil(H/2)*ceil(W/2), 4C]
 * (out_dtype).  The reduction Linear(4C -> 2C) (:575) is psw_linear_fwd.
 */
PSW_API int psw_patch_merge_ln_fwd(const void* x, void* y, const float* gamma, const float* beta,
                           int B, int H, int W, int C, float eps, int in_dtype, int out_dtype, void* stream);

/*
 * Per-stage output head: LayerNorm then NHWC -> NCHW, fp32 contiguous output
 * (SimplePanoSwinTransformer.forward :974-978).  x [B, HW, C] (in_dtype) -> y [B, C, HW] fp32.
 */
PSW_API int psw_layernorm_nchw_fwd(const void* x, float* y, const float* gamma, const float* beta,
                           int B, int64_t HW, int C, float eps, int in_dtype, void* stream);

/*
 * Stem, first layer: conv3x3(cin -> cout, padding 1) + eval-mode BatchNorm + ReLU (PatchEmbed.proj[0..2], reference
 * :743-745) on tcgen05.  img [B, cin, H, W] fp32 NCHW -> out [B, H, W, cout] bf16 NHWC.  w_folded [cout, cin*9]
 * (k = c*9 + ky*3 + kx) and bias_folded [cout] are the fp32 weights / bias with BatchNorm folded in:
 * w * g/sqrt(var+eps), (b - mean) * g/sqrt(var+eps) + beta.  Built for cin = 3 and cout = 32 (embed_dim 96) or 64; a
 * narrower layer is run with its weights / bias zero-padded to 64 output channels (the extra channels come out as 0,
 * e.g. PanoSwin-B: 42 -> 64).  Other shapes return PSW_ERR_UNSUPPORTED.  W must be a multiple of 4.
 */
PSW_API int psw_stem_conv3x3_relu_fwd(const float* img, const float* w_folded, const float* bias_folded, void* out,
                                      int B, int H, int W, int cin, int cout, void* stream);

/*
 * Stem, second layer: conv3x3(32 -> cout, padding 1) + eval-mode BatchNorm + ReLU (PatchEmbed.proj[3..5], reference
 * :746-748; cout = 64 for embed_dim 96) on tcgen05, NHWC bf16 in and out: x [B, H, W, 32] -> out [B, H, W, cout].
 * w_taps [9][cout][32 in] bf16 with tap = ky*3 + kx holds the BatchNorm-folded weights, bias [cout] fp32 the folded
 * bias.  cout must be 32 or 64.
 */
PSW_API int psw_stem_conv3x3_c32_relu_fwd(const void* x, const void* w_taps, const float* bias, void* out,
                                          int B, int H, int W, int cout, void* stream);

/*
 * conv3x3(cin -> cout, stride 1, padding 1) + bias (+ ReLU) of an NHWC bf16 image as a tcgen05 GEMM over shifted 4-D
 * TMA views of the input (no im2col, the zero padding is TMA's out-of-range fill): the second stem layer of the models
 * the dedicated kernel above is not built for (PanoSwin-B: 42 -> 84 channels, zero-padded by the caller to 64 -> 96).
 * x [B, H, W, cin] -> out [B, H, W, cout]; w [cout][3][3][cin] bf16 (BatchNorm folded in), bias [cout] fp32 or NULL.
 * cin must be a multiple of 64, cout of 16.
 */
PSW_API int psw_conv3x3_nhwc_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                                 int cin, int cout, int relu, void* stream);

/*
 * Stem, last layer: the non-overlapping patch convolution conv(cin -> cout, kernel = stride = patch) (PatchEmbed.proj[6],
 * reference :749) as a tcgen05 GEMM over a 3-D TMA view of the NHWC bf16 input (no im2col; bias in the epilogue).
 * x [B, H, W, cin] bf16 -> out [B * H/ph * W/pw, cout] bf16 tokens.  w [cout][ph][pw][cin] bf16 (the reference's
 * [cout, cin, ph, pw] weight permuted to (0, 2, 3, 1)), bias [cout] fp32 or NULL.  H, W must be multiples of the patch.
 */
PSW_API int psw_patch_conv_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                               int cin, int cout, int patch_h, int patch_w, void* stream);

/*
 * fp32 parity path of the stem: direct convolution (square kernel, stride, zero padding) on CUDA-core FMAs with the
 * eval-mode BatchNorm affine (y * bn_scale[c] + bn_shift[c], both or neither) and ReLU in the epilogue — so that the
 * "<= 1e-5 of the reference" path contains no library kernel (PatchEmbed.proj, reference :742-750).
 * in [B, cin, H, W] fp32 NCHW, w [cout, cin, k, k], bias [cout] or NULL -> out fp32, NCHW or (out_nhwc = 1) NHWC tokens.
 */
PSW_API int psw_conv2d_f32_fwd(const float* in, const float* w, const float* bias, const float* bn_scale,
                               const float* bn_shift, float* out, int B, int cin, int H, int W, int cout,
                               int kernel, int stride, int padding, int relu, int out_nhwc, void* stream);

/* dtype conversion helper for activations entering / leaving the bf16 path: n elements. */
PSW_API int psw_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);


/* ================================================================================================================
 * Backward entry points (training: the reference obtains all of these from torch autograd).
 * Same conventions as above.  Gradients that are REDUCTIONS over the batch (dgamma, dbeta, dw, db, dalpha, dbeta
 * tables, dqkv_bias) are zeroed by the call (cudaMemsetAsync on `stream`) and then accumulated with fp32 atomics: the
 * summation order, hence the last bits, varies from run to run.
 * ================================================================================================================ */

/*
 * Gradient of psw_layernorm_fwd (without the position add, whose gradient is the row sum of dy):
 * x [rows, C] (x_dtype), dy [rows, C] (dy_dtype) -> dx [rows, C] (x_dtype), dgamma / dbeta [C] fp32 (both or neither).
 * stats_ws: scratch of 2 * rows floats (mean, rstd per row).  Reference: nn.LayerNorm at :504, :534, :768-772, :975-976.
 */
PSW_API int psw_layernorm_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma, float* dbeta,
                              float* stats_ws, int64_t rows, int C, float eps, int x_dtype, int dy_dtype, void* stream);

/*
 * Gradient of psw_patch_merge_ln_fwd: x [B, H, W, C] (x_dtype), dy [B, ceil(H/2)*ceil(W/2), 4C] (dy_dtype) ->
 * dx [B, H, W, C] (x_dtype: the 2x2 gather's scatter is folded into the store), dgamma / dbeta [4C] fp32.
 * stats_ws: 2 * B * ceil(H/2) * ceil(W/2) floats.  Reference: PatchMerging.forward :563-574.
 */
PSW_API int psw_patch_merge_ln_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma,
                                   float* dbeta, float* stats_ws, int B, int H, int W, int C, float eps, int x_dtype,
                                   int dy_dtype, void* stream);

/*
 * Gradients of y = x . w^T + bias (psw_linear_fwd without epilogue flags): x [M, K], w [N, K], dy [M, N] (all `dtype`)
 * -> dx [M, K] (dx_dtype; NULL to skip), dw [N, K] fp32 (NULL to skip), db [N] fp32 (NULL to skip).
 * PSW_F32: CUDA-core FMAs.  PSW_BF16: dx on the tcgen05 GEMM against the transposed weight, which is written into
 * `workspace` (psw_linear_bwd_workspace_bytes(); without it, or for K % 16 != 0 / N % 8 != 0, the CUDA-core kernel
 * serves); dw on tcgen05 as well, with both operands MN-major straight from the row-major activations (no transposed
 * copies), fp32 accumulation, the rows of the batch split over the grid (N % 8 == 0 and K % 8 == 0, else CUDA cores).
 * Reference: nn.Linear at :287, :309, :55-61, :575.
 */
PSW_API int64_t psw_linear_bwd_workspace_bytes(int64_t M, int N, int K, int dtype);
PSW_API int psw_linear_bwd(const void* x, const void* w, const void* dy, void* dx, float* dw, float* db,
                           int64_t M, int N, int K, int dtype, int dx_dtype, void* workspace, int64_t workspace_bytes,
                           void* stream);

/* Exact (erf) GELU of the training path, which keeps the pre-activation h: y = gelu(h); dh = dy * gelu'(h).
 * n elements (n % 4 == 0), all tensors `dtype`.  Reference: Mlp.act (:51, :57). */
PSW_API int psw_gelu_fwd(const void* h, void* y, int64_t n, int dtype, void* stream);
PSW_API int psw_gelu_bwd(const void* h, const void* dy, void* dh, int64_t n, int dtype, void* stream);

/* dst [cols, rows] = src [rows, cols]^T (helper of the bf16 backward path). */
PSW_API int psw_transpose(const void* src, void* dst, int64_t rows, int64_t cols, int dtype, void* stream);

/*
 * Train-mode BatchNorm2d + ReLU of the stem on NHWC bf16 activations (PatchEmbed.proj[1:3] / [4:6], reference :743-748
 * under model.train(): batch statistics), x / y / dy / dx [pixels, C] bf16, C % 8 == 0 and C <= 256.
 *   psw_bn_stats_fwd        sum[c], sumsq[c] (fp32) over the pixels
 *   psw_bn_apply_relu_fwd   y = relu(x * scale[c] + shift[c]) with scale = gamma * rstd, shift = beta - mean * scale
 *   psw_bn_relu_bwd         dx and dgamma / dbeta [C] fp32 from x, the forward output y (its sign is the ReLU mask), dy, the
 *                           batch mean / rstd and gamma
 */
PSW_API int psw_bn_stats_fwd(const void* x, float* sum, float* sumsq, int64_t npix, int C, void* stream);
PSW_API int psw_bn_apply_relu_fwd(const void* x, void* y, const float* scale, const float* shift, int64_t npix, int C,
                                  void* stream);
PSW_API int psw_bn_relu_bwd(const void* x, const void* y, const void* dy, const float* mean, const float* rstd,
                            const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t npix, int C, void* stream);

/*
 * Gradient of the fused window attention (psw_window_attn_fwd and psw_window_attn_full_fwd: same function of qkv,
 * alpha, beta): qkv [B, H, W, 3C], dout [B, H, W, C] (both `dtype`) -> dqkv [B, H, W, 3C] (`dtype`, every element
 * written), dalpha / dbeta [(2*window-1)^2, heads] fp32, dqkv_bias [3C] fp32 = the gradient that reaches the qkv bias
 * through the PADDING cells (zero tokens whose q / k / v are the bias, reference :486-491; the bias gradient of the
 * real tokens is psw_linear_bwd's db).  The probabilities are recomputed from qkv; uv / mask carry no gradient.
 * qkv_bias and dqkv_bias are both NULL or both given; dalpha is untouched in planar mode.  Any window / head_dim
 * (CUDA-core kernel); PSW_BF16 with window 7 and head_dim 32 runs on tensor cores (psw_attn_bwd_mma.cu), where padding
 * cells take the bf16-rounded bias like psw_window_attn_full_fwd.
 */
PSW_API int psw_window_attn_bwd(const void* qkv, const void* dout, const float* alpha, const float* beta,
                                const float* qkv_bias, const float* uv, const float* mask, void* dqkv, float* dalpha,
                                float* dbeta, float* dqkv_bias, int B, int H, int W, int C, int heads, int window,
                                int shift, int pano_mode, float scale, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PANOSWIN_B200_H_ */
