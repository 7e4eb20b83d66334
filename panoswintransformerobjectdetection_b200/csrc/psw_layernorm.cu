// K1 / K4: vectorised bandwidth kernels — LayerNorm (+ optional position add), PatchMerging
// gather + LayerNorm(4C), and the per-stage output LayerNorm + NHWC->NCHW transpose.
//
// One warp owns one row; the row lives in registers (VPL 4-element vectors per lane), so every input
// byte is read from HBM exactly once and statistics are two-pass (mean, then centred variance) in fp32
// like torch.nn.functional.layer_norm.  Algorithmic bytes per row: C*(sizeof(in)+sizeof(out)).
#include "psw_common.cuh"

namespace psw {

constexpr int LN_WARPS = 8;

// Row gather policies -----------------------------------------------------------------------------
struct PlainRows {
  int64_t rows;
  int C;
  __device__ __forceinline__ int64_t num_rows() const { return rows; }
  // element offset of 4-vector `vec` of row `r`, or -1 for an implicit zero vector
  __device__ __forceinline__ int64_t offset(int64_t r, int vec) const { return r * C + (int64_t)vec * 4; }
};

// PatchMerging (reference :563-573): output row (b, i2, j2) = concat of x[b, 2*i2+dh, 2*j2+dw, :] for
// quadrant q = 0..3 with dh = q & 1, dw = q >> 1; cells beyond an odd H / W are zero.
struct MergeRows {
  int B, H, W, C, H2, W2;
  __device__ __forceinline__ int64_t num_rows() const { return (int64_t)B * H2 * W2; }
  __device__ __forceinline__ int64_t offset(int64_t r, int vec) const {
    int vpc = C >> 2;
    int q = vec / vpc;
    int within = vec - q * vpc;
    int j2 = (int)(r % W2);
    int64_t t = r / W2;
    int i2 = (int)(t % H2);
    int b = (int)(t / H2);
    int h = 2 * i2 + (q & 1);
    int w = 2 * j2 + (q >> 1);
    if (h >= H || w >= W) return -1;
    return (((int64_t)b * H + h) * W + w) * C + (int64_t)within * 4;
  }
};

template <int VPL, typename TI, typename TO, typename Rows>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_rows_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ pos, int64_t pos_rows,
                      Rows rows, int Cout, float eps) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = Cout >> 2;
  const int64_t nrows = rows.num_rows();
  const float inv_c = 1.0f / (float)Cout;
  for (int64_t r = (int64_t)blockIdx.x * LN_WARPS + warp; r < nrows; r += (int64_t)gridDim.x * LN_WARPS) {
    float v[VPL][4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int vec = lane + 32 * k;
      v[k][0] = v[k][1] = v[k][2] = v[k][3] = 0.f;
      if (vec < nvec) {
        int64_t off = rows.offset(r, vec);
        if (off >= 0) load4(x + off, v[k]);
      }
      s += (v[k][0] + v[k][1]) + (v[k][2] + v[k][3]);
    }
    const float mean = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (lane + 32 * k < nvec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float d = v[k][e] - mean;
          q += d * d;
        }
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    const float* prow = pos ? pos + (r % pos_rows) * Cout : nullptr;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      int vec = lane + 32 * k;
      if (vec < nvec) {
        float g[4], b[4], o[4];
        load4(gamma + vec * 4, g);
        load4(beta + vec * 4, b);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = (v[k][e] - mean) * rstd * g[e] + b[e];
        if (prow) {
          float pp[4];
          load4(prow + vec * 4, pp);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += pp[e];
        }
        store4(y + r * Cout + (int64_t)vec * 4, o);
      }
    }
  }
}

template <typename TI, typename TO, typename Rows>
static int launch_ln(const TI* x, TO* y, const float* gamma, const float* beta, const float* pos, int64_t pos_rows,
                     Rows rows, int64_t nrows, int Cout, float eps, cudaStream_t st) {
  int nvec = Cout / 4;
  int vpl = (nvec + 31) / 32;
  int64_t want = (nrows + LN_WARPS - 1) / LN_WARPS;
  int64_t cap = (int64_t)num_sms() * 16;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
  static const int kInst[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};   // instantiated vectors-per-lane
  int inst = 32;
  for (int k = 9; k >= 0; --k)
    if (kInst[k] >= vpl) inst = kInst[k];
#define PSW_LN_CASE(V)                                                                                     \
  case V:                                                                                                  \
    layernorm_rows_kernel<V, TI, TO, Rows><<<blocks, LN_WARPS * 32, 0, st>>>(x, y, gamma, beta, pos, pos_rows, \
                                                                             rows, Cout, eps);             \
    break;
  switch (inst) {
    PSW_LN_CASE(1) PSW_LN_CASE(2) PSW_LN_CASE(3) PSW_LN_CASE(4) PSW_LN_CASE(6) PSW_LN_CASE(8) PSW_LN_CASE(12)
    PSW_LN_CASE(16) PSW_LN_CASE(24) PSW_LN_CASE(32)
  }
#undef PSW_LN_CASE
  return launch_status("layernorm_rows_kernel");
}

template <typename Rows>
static int dispatch_ln(const void* x, void* y, const float* gamma, const float* beta, const float* pos,
                       int64_t pos_rows, Rows rows, int64_t nrows, int Cout, float eps, int in_dtype, int out_dtype,
                       cudaStream_t st) {
  if (in_dtype == PSW_F32 && out_dtype == PSW_F32)
    return launch_ln((const float*)x, (float*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_F32 && out_dtype == PSW_BF16)
    return launch_ln((const float*)x, (bf16*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_BF16 && out_dtype == PSW_BF16)
    return launch_ln((const bf16*)x, (bf16*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_BF16 && out_dtype == PSW_F32)
    return launch_ln((const bf16*)x, (float*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  set_error("layernorm: unknown dtype combination %d -> %d", in_dtype, out_dtype);
  return PSW_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm + NHWC -> NCHW (fp32 out).  A CTA owns 32 consecutive tokens of one image: 8 warps normalise
// 4 rows each into a [C][33] shared tile, then every warp streams whole channels out as 128-byte rows.
// ---------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void __launch_bounds__(256)
layernorm_nchw_kernel(const TI* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int64_t HW, int C, float eps) {
  extern __shared__ float tile[];                 // [C][33]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t tiles_per_img = (HW + 31) / 32;
  const int b = (int)(blockIdx.x / tiles_per_img);
  const int64_t t0 = (blockIdx.x % tiles_per_img) * 32;
  const float inv_c = 1.0f / (float)C;
  for (int rr = warp; rr < 32; rr += 8) {
    int64_t t = t0 + rr;
    if (t >= HW) break;
    const TI* row = x + ((int64_t)b * HW + t) * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += to_f32(row[c]);
    float mean = warp_sum(s) * inv_c;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) {
      float d = to_f32(row[c]) - mean;
      q += d * d;
    }
    float rstd = rsqrtf(warp_sum(q) * inv_c + eps);
    for (int c = lane; c < C; c += 32) tile[c * 33 + rr] = (to_f32(row[c]) - mean) * rstd * gamma[c] + beta[c];
  }
  __syncthreads();
  int64_t t = t0 + lane;
  if (t < HW) {
    for (int c = warp; c < C; c += 8) y[((int64_t)b * C + c) * HW + t] = tile[c * 33 + lane];
  }
}

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, const float* pos,
                                 int64_t rows, int C, int64_t pos_rows, float eps, int in_dtype, int out_dtype,
                                 void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: null pointer");
  PSW_REQUIRE(rows > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: rows=%lld C=%d", (long long)rows, C);
  PSW_REQUIRE(C % 4 == 0 && C <= 4096, PSW_ERR_UNSUPPORTED, "psw_layernorm_fwd: C=%d must be a multiple of 4, <= 4096", C);
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && aligned16(pos), PSW_ERR_BAD_ARG,
              "psw_layernorm_fwd: pointers must be 16-byte aligned");
  PSW_REQUIRE(pos == nullptr || pos_rows > 0, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: pos given but pos_rows <= 0");
  PlainRows pr{rows, C};
  return dispatch_ln(x, y, gamma, beta, pos, pos_rows > 0 ? pos_rows : 1, pr, rows, C, eps, in_dtype, out_dtype,
                     (cudaStream_t)stream);
}

extern "C" PSW_API int psw_patch_merge_ln_fwd(const void* x, void* y, const float* gamma, const float* beta, int B, int H,
                                      int W, int C, float eps, int in_dtype, int out_dtype, void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_fwd: bad dims");
  PSW_REQUIRE(C % 4 == 0 && 4 * C <= 4096, PSW_ERR_UNSUPPORTED, "psw_patch_merge_ln_fwd: C=%d must be a multiple of 4, <= 1024", C);
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), PSW_ERR_BAD_ARG,
              "psw_patch_merge_ln_fwd: pointers must be 16-byte aligned");
  MergeRows mr{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  int64_t nrows = (int64_t)B * mr.H2 * mr.W2;
  return dispatch_ln(x, y, gamma, beta, nullptr, 1, mr, nrows, 4 * C, eps, in_dtype, out_dtype, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_layernorm_nchw_fwd(const void* x, float* y, const float* gamma, const float* beta, int B,
                                      int64_t HW, int C, float eps, int in_dtype, void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: null pointer");
  PSW_REQUIRE(B > 0 && HW > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: bad dims");
  size_t smem = (size_t)C * 33 * sizeof(float);
  PSW_REQUIRE(smem <= 200 * 1024, PSW_ERR_UNSUPPORTED, "psw_layernorm_nchw_fwd: C=%d too large", C);
  int64_t blocks = (int64_t)B * ((HW + 31) / 32);
  PSW_REQUIRE(blocks < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_layernorm_nchw_fwd: too many tiles");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == PSW_F32) {
    PSW_CUDA(cudaFuncSetAttribute(layernorm_nchw_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layernorm_nchw_kernel<float><<<(unsigned)blocks, 256, smem, st>>>((const float*)x, y, gamma, beta, HW, C, eps);
  } else if (in_dtype == PSW_BF16) {
    PSW_CUDA(cudaFuncSetAttribute(layernorm_nchw_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    layernorm_nchw_kernel<bf16><<<(unsigned)blocks, 256, smem, st>>>((const bf16*)x, y, gamma, beta, HW, C, eps);
  } else {
    PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: unknown dtype %d", in_dtype);
  }
  return launch_status("layernorm_nchw_kernel");
}
