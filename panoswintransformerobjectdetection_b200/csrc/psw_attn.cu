// psw_window_attn_fwd: argument validation and dispatch between the fp32 parity kernel
// (psw_attn_simt.cu) and the bf16 tcgen05 kernel (psw_attn_tc.cu).
#include "psw_common.cuh"

namespace psw {
template <typename T>
int window_attn_simt(const T* qkv, T* out, const float* alpha, const float* beta, const float* qkv_bias,
                     const float* uv, const float* mask, int B, int H, int W, int C, int heads, int window, int shift,
                     int pano, float scale, cudaStream_t st);
void attn_debug_set_hc(int hc);
int window_attn_tc(const bf16* qkv, bf16* out, const float* alpha, const float* beta, const void* tables,
                   const float* qkv_bias, const void* hav_table, const float* mask, const void* bias_full,
                   bool bias_row_present, int B, int H, int W, int C, int heads, int window, int shift, int pano, float scale,
                   long long* dbg, int mode, cudaStream_t st);
int window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask, void* table, int H, int W,
                     int heads, int window, int shift, int pano, cudaStream_t st);
int window_bias_tables(const float* alpha, const float* beta, void* tables, int heads, int window, cudaStream_t st);
int window_hav_table(const float* uv, void* table, int H, int W, int window, int shift, cudaStream_t st);
}  // namespace psw

using namespace psw;

static int check_attn_args(const void* qkv, void* out, const float* alpha, const float* beta, const void* uv, int B,
                           int H, int W, int C, int heads, int window, int shift, int pano_mode) {
  PSW_REQUIRE(qkv && out && alpha && beta, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && heads > 0 && window > 0, PSW_ERR_BAD_ARG,
              "psw_window_attn_fwd: bad dims B=%d H=%d W=%d C=%d heads=%d window=%d", B, H, W, C, heads, window);
  PSW_REQUIRE(C % heads == 0, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: channels %d not divisible by heads %d", C, heads);
  PSW_REQUIRE(shift >= 0 && shift < window, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: shift_size must be in [0, window)");
  PSW_REQUIRE(!pano_mode || uv, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: pano mode needs the uv table (fp32) / the great-circle table (bf16)");
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_fwd: too many tokens");
  return 0;
}

extern "C" PSW_API int psw_window_attn_fwd(const void* qkv, void* out, const float* alpha, const float* beta,
                                   const void* bias_tables, const float* qkv_bias, const float* uv,
                                   const void* hav_table, const float* mask, int B, int H, int W, int C, int heads,
                                   int window, int shift, int pano_mode, float scale, int dtype, void* stream) {
  int rc = check_attn_args(qkv, out, alpha, beta, dtype == PSW_BF16 ? hav_table : (const void*)uv, B, H, W, C, heads,
                           window, shift, pano_mode);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32)
    return window_attn_simt<float>((const float*)qkv, (float*)out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads,
                                   window, shift, pano_mode, scale, st);
  PSW_REQUIRE(dtype == PSW_BF16, PSW_ERR_BAD_ARG, "psw_window_attn_fwd: unknown dtype %d", dtype);
  PSW_REQUIRE(window * window <= 64 && C / heads == 32, PSW_ERR_UNSUPPORTED,
              "psw_window_attn_fwd(bf16): tcgen05 kernel needs window^2 <= 64 and head_dim == 32 (window=%d head_dim=%d)",
              window, C / heads);
  PSW_REQUIRE(aligned16(qkv) && aligned16(out), PSW_ERR_BAD_ARG, "psw_window_attn_fwd(bf16): pointers must be 16-byte aligned");
  PSW_REQUIRE(aligned16(hav_table), PSW_ERR_BAD_ARG, "psw_window_attn_fwd(bf16): great-circle table must be 16-byte aligned");
  PSW_REQUIRE(aligned16(bias_tables), PSW_ERR_BAD_ARG, "psw_window_attn_fwd(bf16): bias tables must be 16-byte aligned");
  return window_attn_tc((const bf16*)qkv, (bf16*)out, alpha, beta, bias_tables, qkv_bias, hav_table, mask, nullptr, false, B,
                        H, W, C, heads, window, shift, pano_mode, scale, nullptr, 0, st);
}

// bf16 production path: every additive term of the logits (great-circle bias, relative-position bias, planar shift
// mask) comes from the table psw_window_bias_full() built for this block and resolution.
extern "C" PSW_API int psw_window_attn_full_fwd(const void* qkv, void* out, const void* bias_full, const float* qkv_bias,
                                                int64_t qkv_rows, int B, int H, int W, int C, int heads, int window,
                                                int shift, int pano_mode, float scale, void* stream) {
  PSW_REQUIRE(qkv && out && bias_full, PSW_ERR_BAD_ARG, "psw_window_attn_full_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && heads > 0 && window > 0, PSW_ERR_BAD_ARG,
              "psw_window_attn_full_fwd: bad dims B=%d H=%d W=%d C=%d heads=%d window=%d", B, H, W, C, heads, window);
  PSW_REQUIRE(C % heads == 0, PSW_ERR_BAD_ARG, "psw_window_attn_full_fwd: channels %d not divisible by heads %d", C, heads);
  PSW_REQUIRE(shift >= 0 && shift < window, PSW_ERR_BAD_ARG, "psw_window_attn_full_fwd: shift_size must be in [0, window)");
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_full_fwd: too many tokens");
  PSW_REQUIRE(window * window <= 64 && C / heads == 32, PSW_ERR_UNSUPPORTED,
              "psw_window_attn_full_fwd: tcgen05 kernel needs window^2 <= 64 and head_dim == 32 (window=%d head_dim=%d)",
              window, C / heads);
  PSW_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(bias_full), PSW_ERR_BAD_ARG,
              "psw_window_attn_full_fwd: pointers must be 16-byte aligned");
  const int64_t T = (int64_t)B * H * W;
  PSW_REQUIRE(qkv_rows == T || qkv_rows == T + 1, PSW_ERR_BAD_ARG,
              "psw_window_attn_full_fwd: qkv_rows must be B*H*W (or B*H*W + 1 with the bias row), got %lld", (long long)qkv_rows);
  return window_attn_tc((const bf16*)qkv, (bf16*)out, nullptr, nullptr, nullptr, qkv_bias, nullptr, nullptr, bias_full,
                        qkv_rows == T + 1, B, H, W, C, heads, window, shift, pano_mode, scale, nullptr, 0, (cudaStream_t)stream);
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

extern "C" PSW_API int psw_window_bias_tables(const float* alpha, const float* beta, void* tables, int heads, int window,
                                              void* stream) {
  PSW_REQUIRE(alpha && beta && tables && heads > 0 && window > 0, PSW_ERR_BAD_ARG, "psw_window_bias_tables: bad arguments");
  PSW_REQUIRE(aligned16(tables), PSW_ERR_BAD_ARG, "psw_window_bias_tables: tables must be 16-byte aligned");
  return window_bias_tables(alpha, beta, tables, heads, window, (cudaStream_t)stream);
}

// Diagnostics: same as the PSW_BF16 path of psw_window_attn_fwd, plus per-phase SM-cycle totals of CTA 0 written to
// phase_cycles[6] (device memory, may be NULL): {wait-for-loads, S MMA, softmax, P.V MMA, store, steps}.
// mode 1 runs the memory skeleton only (same gathers and stores, no MMA / softmax; output = q rows); mode bits
// [8,12) force the number of heads per work item / image pairs per unit (15: window-pair kernel, 14: no TMA gather);
// bit 12: the qkv tensor has the extra bias row (enables the TMA gather loader).
extern "C" PSW_API int psw_window_attn_fwd_profile(const void* qkv, void* out, const float* alpha, const float* beta,
                                                   const void* bias_tables, const float* qkv_bias,
                                                   const void* hav_table, const void* bias_full, int B, int H, int W,
                                                   int C, int heads, int window, int shift, float scale,
                                                   long long* phase_cycles, int mode, void* stream) {
  int rc = check_attn_args(qkv, out, alpha, beta, bias_full ? bias_full : hav_table, B, H, W, C, heads, window, shift, 1);
  if (rc) return rc;
  PSW_REQUIRE(C / heads == 32 && (mode & 0xff) <= 4, PSW_ERR_BAD_ARG, "psw_window_attn_fwd_profile: bad arguments");
  attn_debug_set_hc((mode >> 8) & 15);                       // bits [8,12): force the heads-per-item choice
  rc = window_attn_tc((const bf16*)qkv, (bf16*)out, alpha, beta, bias_tables, qkv_bias, hav_table, nullptr, bias_full,
                      ((mode >> 12) & 1) != 0, B, H, W, C, heads, window, shift, 1, scale, phase_cycles, mode & 0xff,
                      (cudaStream_t)stream);
  attn_debug_set_hc(0);
  return rc;
}

extern "C" PSW_API int psw_window_grid(int H, int W, int window, int pano_mode, int* nwh, int* nww) {
  PSW_REQUIRE(H > 0 && W > 0 && window > 0 && nwh && nww, PSW_ERR_BAD_ARG, "psw_window_grid: bad arguments");
  WinGeom g = make_geom(H, W, window, 0, pano_mode);
  *nwh = g.nWh;
  *nww = g.nWw;
  return 0;
}

extern "C" PSW_API int psw_window_hav_table(const float* uv, void* table, int H, int W, int window, int shift,
                                            void* stream) {
  PSW_REQUIRE(uv && table, PSW_ERR_BAD_ARG, "psw_window_hav_table: null pointer");
  PSW_REQUIRE(H > 0 && W > 0 && window > 0 && window * window <= 64 && shift >= 0 && shift < window, PSW_ERR_BAD_ARG,
              "psw_window_hav_table: bad dims H=%d W=%d window=%d shift=%d", H, W, window, shift);
  PSW_REQUIRE(aligned16(table), PSW_ERR_BAD_ARG, "psw_window_hav_table: table must be 16-byte aligned");
  return window_hav_table(uv, table, H, W, window, shift, (cudaStream_t)stream);
}

// Debug / cross-check entry (not part of the reference-facing contract): the CUDA-core kernel on bf16 storage.
extern "C" PSW_API int psw_window_attn_fwd_simt_bf16(const void* qkv, void* out, const float* alpha, const float* beta,
                                             const float* qkv_bias, const float* uv, const float* mask, int B, int H,
                                             int W, int C, int heads, int window, int shift, int pano_mode,
                                             float scale, void* stream) {
  int rc = check_attn_args(qkv, out, alpha, beta, uv, B, H, W, C, heads, window, shift, pano_mode);
  if (rc) return rc;
  return window_attn_simt<bf16>((const bf16*)qkv, (bf16*)out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads, window,
                                shift, pano_mode, scale, (cudaStream_t)stream);
}
