// Window-attention entry points: argument validation and dispatch.
//   psw_window_attn_fwd       generic CUDA-core kernel (psw_attn_simt.cu): the fp32 parity path, and the bf16 route
//                             for any window / head_dim the tcgen05 kernel is not instantiated for
//   psw_window_attn_full_fwd  bf16 tcgen05 kernel (psw_attn_tc.cu), window 7 / head_dim 32, bias table precomputed
//   psw_window_attn_bwd       gradients of the fused op (psw_attn_bwd.cu)
#include "psw_common.cuh"

namespace psw {
template <typename T>
int window_attn_simt(const T* qkv, T* out, const float* alpha, const float* beta, const float* qkv_bias,
                     const float* uv, const float* mask, int B, int H, int W, int C, int heads, int window, int shift,
                     int pano, float scale, cudaStream_t st);
int window_attn_tc(const bf16* qkv, bf16* out, const float* qkv_bias, const void* bias_full, int B, int H, int W, int C,
                   int heads, int window, int shift, int pano, float scale, long long* dbg, int mode, int variant,
                   cudaStream_t st);
int window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask, void* table, int H, int W,
                     int heads, int window, int shift, int pano, cudaStream_t st);
}  // namespace psw

using namespace psw;

static int check_attn_args(const char* fn, const void* qkv, const void* out, int B, int H, int W, int C, int heads, int window,
                           int shift) {
  PSW_REQUIRE(qkv && out, PSW_ERR_BAD_ARG, "%s: null pointer", fn);
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && heads > 0 && window > 0, PSW_ERR_BAD_ARG,
              "%s: bad dims B=%d H=%d W=%d C=%d heads=%d window=%d", fn, B, H, W, C, heads, window);
  PSW_REQUIRE(C % heads == 0, PSW_ERR_BAD_ARG, "%s: channels %d not divisible by heads %d", fn, C, heads);
  PSW_REQUIRE(shift >= 0 && shift < window, PSW_ERR_BAD_ARG, "%s: shift_size must be in [0, window)", fn);
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31), PSW_ERR_UNSUPPORTED, "%s: too many tokens", fn);
  return 0;
}

extern "C" PSW_API int psw_window_attn_fwd(const void* qkv, void* out, const float* alpha, const float* beta,
                                           const float* qkv_bias, const float* uv, const float* mask, int B, int H, int W,
                                           int C, int heads, int window, int shift, int pano_mode, float scale, int dtype,
                                           void* stream) {
  int rc = check_attn_args("psw_window_attn_fwd", qkv, out, B, H, W, C, heads, window, shift);
  if (rc) return rc;
  PSW_REQUIRE(alpha && beta, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: null alpha / beta table");
  PSW_REQUIRE(!pano_mode || uv, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: pano mode needs the uv table");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32)
    return window_attn_simt<float>((const float*)qkv, (float*)out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads,
                                   window, shift, pano_mode, scale, st);
  PSW_REQUIRE(dtype == PSW_BF16, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: unknown dtype %d", dtype);
  return window_attn_simt<bf16>((const bf16*)qkv, (bf16*)out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads, window,
                                shift, pano_mode, scale, st);
}

extern "C" PSW_API int psw_window_attn_full_fwd(const void* qkv, void* out, const void* bias_full, const float* qkv_bias,
                                                int B, int H, int W, int C, int heads, int window, int shift,
                                                int pano_mode, float scale, void* stream) {
  int rc = check_attn_args("psw_window_attn_full_fwd", qkv, out, B, H, W, C, heads, window, shift);
  if (rc) return rc;
  PSW_REQUIRE(bias_full, PSW_ERR_BAD_ARG, "psw_window_attn_full_fwd: null bias table");
  PSW_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(bias_full), PSW_ERR_BAD_ARG,
              "psw_window_attn_full_fwd: pointers must be 16-byte aligned");
  return window_attn_tc((const bf16*)qkv, (bf16*)out, qkv_bias, bias_full, B, H, W, C, heads, window, shift, pano_mode,
                        scale, nullptr, 0, 0, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_window_attn_full_supported(int window, int head_dim) {
  return (window == 7 && head_dim == 32) ? 1 : 0;
}

extern "C" PSW_API int64_t psw_window_bias_full_bytes(int H, int W, int heads, int window, int pano_mode) {
  if (H <= 0 || W <= 0 || heads <= 0 || window <= 0) return 0;
  WinGeom g = make_geom(H, W, window, 0, pano_mode);
  return (int64_t)g.nWh * g.nWw * heads * 13 * 64 * 16;
}

extern "C" PSW_API int psw_window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask,
                                            void* table, int H, int W, int heads, int window, int shift, int pano_mode,
                                            void* stream) {
  PSW_REQUIRE(alpha && beta && table, PSW_ERR_BAD_ARG, "psw_window_bias_full: null pointer");
  PSW_REQUIRE(H > 0 && W > 0 && heads > 0 && window > 0 && shift >= 0 && shift < window, PSW_ERR_BAD_ARG,
              "psw_window_bias_full: bad dims H=%d W=%d heads=%d window=%d shift=%d", H, W, heads, window, shift);
  PSW_REQUIRE(!pano_mode || uv, PSW_ERR_BAD_ARG, "psw_window_bias_full: pano mode needs the uv table");
  PSW_REQUIRE(aligned16(table), PSW_ERR_BAD_ARG, "psw_window_bias_full: table must be 16-byte aligned");
  return window_bias_full(alpha, beta, uv, mask, table, H, W, heads, window, shift, pano_mode, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_window_grid(int H, int W, int window, int pano_mode, int* nwh, int* nww) {
  PSW_REQUIRE(H > 0 && W > 0 && window > 0 && nwh && nww, PSW_ERR_BAD_ARG, "psw_window_grid: bad arguments");
  WinGeom g = make_geom(H, W, window, 0, pano_mode);
  *nwh = g.nWh;
  *nww = g.nWw;
  return 0;
}

// Host-only: the window geometry used by every attention kernel (psw::source_token).  Fills
// map[(nWh*ws) * (nWw*ws)] with the flat source token of every cell of the padded shifted map (-1 = zero padding)
// and returns the padded height / width through hp / wp.  No GPU involved.
extern "C" PSW_API int psw_window_source_map(int H, int W, int window, int shift, int pano_mode, int* map, int capacity,
                                             int* hp, int* wp) {
  PSW_REQUIRE(H > 0 && W > 0 && window > 0 && shift >= 0 && shift < window && hp && wp, PSW_ERR_BAD_ARG,
              "psw_window_source_map: bad arguments");
  WinGeom g = make_geom(H, W, window, shift, pano_mode);
  *hp = g.nWh * window;
  *wp = g.nWw * window;
  if (!map) return 0;
  PSW_REQUIRE(capacity >= *hp * *wp, PSW_ERR_BAD_ARG, "psw_window_source_map: capacity %d < %d", capacity, *hp * *wp);
  for (int i = 0; i < *hp; ++i)
    for (int j = 0; j < *wp; ++j) map[i * *wp + j] = source_token(g, i, j);
  return 0;
}

#ifdef PSW_DIAGNOSTICS
// Diagnostics build only (include/panoswin_b200_debug.h): psw_window_attn_full_fwd plus per-phase SM-cycle totals of
// CTA 0 in phase_cycles[6] (device memory, may be NULL), a `mode` (1 memory skeleton, 2 no bias loads, 3 no q/k/v
// loads) and a `variant` (schedule / exp2 evaluation, see window_attn_tc).
extern "C" PSW_API int psw_diag_window_attn_full(const void* qkv, void* out, const void* bias_full, const float* qkv_bias,
                                                 int B, int H, int W, int C, int heads, int window, int shift,
                                                 int pano_mode, float scale, long long* phase_cycles, int mode,
                                                 int variant, void* stream) {
  int rc = check_attn_args("psw_diag_window_attn_full", qkv, out, B, H, W, C, heads, window, shift);
  if (rc) return rc;
  return window_attn_tc((const bf16*)qkv, (bf16*)out, qkv_bias, bias_full, B, H, W, C, heads, window, shift, pano_mode,
                        scale, phase_cycles, mode, variant, (cudaStream_t)stream);
}
#endif
