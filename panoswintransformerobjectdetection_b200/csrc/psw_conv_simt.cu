// fp32 parity path of the stem (PatchEmbed.proj, reference simple_panoswin_transformer.py:742-750): a direct
// convolution on CUDA-core FMAs with the eval-mode BatchNorm affine and ReLU applied in the epilogue, so that the
// "<= 1e-5 of the reference" path contains no library kernel.  Any channel counts, square kernel k, stride s, zero
// padding p.  Accumulation order per output: input channel, then kernel row, then kernel column (fp32 FMA).
// The throughput path is psw_stem.cu / psw_stem2.cu / the tcgen05 patch-conv GEMM.
#include "psw_common.cuh"

namespace psw {

constexpr int CV_TX = 32, CV_TY = 8, CV_CO = 8;      // 32 x 8 output pixels per block, 8 output channels per thread

// in [B, cin, H, W] fp32 NCHW; w [cout, cin, k, k]; out NCHW [B, cout, Ho, Wo] or NHWC [B, Ho, Wo, cout]
__global__ void __launch_bounds__(CV_TX * CV_TY)
conv2d_f32_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ bn_scale, const float* __restrict__ bn_shift, float* __restrict__ out,
                  int B, int cin, int H, int W, int cout, int k, int stride, int pad, int Ho, int Wo, int relu, int nhwc) {
  extern __shared__ float sw[];                       // [CV_CO][cin * k * k] weights of this block's output channels
  const int co_groups = (cout + CV_CO - 1) / CV_CO;
  const int b = blockIdx.z / co_groups;
  const int co0 = (blockIdx.z - b * co_groups) * CV_CO;
  const int kk = cin * k * k;
  for (int i = threadIdx.y * CV_TX + threadIdx.x; i < CV_CO * kk; i += CV_TX * CV_TY) {
    const int c = i / kk;
    sw[i] = (co0 + c < cout) ? w[(size_t)(co0 + c) * kk + (i - c * kk)] : 0.f;
  }
  __syncthreads();
  const int ox = blockIdx.x * CV_TX + threadIdx.x;
  const int oy = blockIdx.y * CV_TY + threadIdx.y;
  if (ox >= Wo || oy >= Ho) return;
  float acc[CV_CO];
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) acc[c] = 0.f;
  const float* inb = in + (size_t)b * cin * H * W;
  for (int ci = 0; ci < cin; ++ci)
    for (int ky = 0; ky < k; ++ky) {
      const int iy = oy * stride - pad + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < k; ++kx) {
        const int ix = ox * stride - pad + kx;
        if (ix < 0 || ix >= W) continue;
        const float v = inb[((size_t)ci * H + iy) * W + ix];
        const float* wp = sw + (ci * k + ky) * k + kx;
#pragma unroll
        for (int c = 0; c < CV_CO; ++c) acc[c] = fmaf(v, wp[c * kk], acc[c]);
      }
    }
#pragma unroll
  for (int c = 0; c < CV_CO; ++c) {
    const int co = co0 + c;
    if (co >= cout) break;
    float v = acc[c] + (bias ? bias[co] : 0.f);
    if (bn_scale) v = v * bn_scale[co] + bn_shift[co];
    if (relu) v = fmaxf(v, 0.f);
    if (nhwc) out[(((size_t)b * Ho + oy) * Wo + ox) * cout + co] = v;
    else out[(((size_t)b * cout + co) * Ho + oy) * Wo + ox] = v;
  }
}

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_conv2d_f32_fwd(const float* in, const float* w, const float* bias, const float* bn_scale,
                                          const float* bn_shift, float* out, int B, int cin, int H, int W, int cout,
                                          int kernel, int stride, int padding, int relu, int out_nhwc, void* stream) {
  PSW_REQUIRE(in && w && out, PSW_ERR_BAD_ARG, "psw_conv2d_f32_fwd: null pointer");
  PSW_REQUIRE((bn_scale == nullptr) == (bn_shift == nullptr), PSW_ERR_BAD_ARG, "psw_conv2d_f32_fwd: bn_scale and bn_shift go together");
  PSW_REQUIRE(B > 0 && cin > 0 && H > 0 && W > 0 && cout > 0 && kernel > 0 && stride > 0 && padding >= 0, PSW_ERR_BAD_ARG,
              "psw_conv2d_f32_fwd: bad dims");
  const int Ho = (H + 2 * padding - kernel) / stride + 1, Wo = (W + 2 * padding - kernel) / stride + 1;
  PSW_REQUIRE(Ho > 0 && Wo > 0, PSW_ERR_BAD_ARG, "psw_conv2d_f32_fwd: empty output");
  const size_t smem = sizeof(float) * CV_CO * (size_t)cin * kernel * kernel;
  PSW_REQUIRE(smem <= 200 * 1024, PSW_ERR_UNSUPPORTED, "psw_conv2d_f32_fwd: cin * k * k = %d too large", cin * kernel * kernel);
  const int64_t gz = (int64_t)B * ((cout + CV_CO - 1) / CV_CO);
  PSW_REQUIRE(gz <= 65535 && (Ho + CV_TY - 1) / CV_TY <= 65535, PSW_ERR_UNSUPPORTED, "psw_conv2d_f32_fwd: grid too large (B * cout / 8 <= 65535)");
  PSW_CUDA(cudaFuncSetAttribute(conv2d_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((Wo + CV_TX - 1) / CV_TX, (Ho + CV_TY - 1) / CV_TY, (unsigned)gz);
  conv2d_f32_kernel<<<grid, dim3(CV_TX, CV_TY), smem, (cudaStream_t)stream>>>(in, w, bias, bn_scale, bn_shift, out, B, cin, H, W, cout,
                                                                             kernel, stride, padding, Ho, Wo, relu, out_nhwc);
  return launch_status("conv2d_f32_kernel");
}
