// Train-mode BatchNorm2d + ReLU of the stem on NHWC bf16 activations (PatchEmbed.proj[1:3], [4:6] of the reference,
// simple_panoswin_transformer.py:743-748, under model.train(): batch statistics), forward and backward -- SURVEY.md §8
// f-3.  Pure bandwidth kernels: a thread owns 8 adjacent channels (one 16-byte load) of a pixel, a CTA walks over pixels
// with a grid stride, per-channel sums are reduced in shared memory and leave with one atomic per channel and CTA.
//   forward   stats:  sum[c], sumsq[c] over the B*H*W pixels            (x read once)
//             apply:  y = relu((x - mean) * rstd * gamma + beta)        mean / rstd from the sums, biased variance
//   backward  reduce: g = dy * (y > 0); sdy[c] = sum g, sdyx[c] = sum g * xhat          (= d beta, d gamma)
//             dx:     dx = gamma * rstd * (g - sdy / n - xhat * sdyx / n)
// C must be a multiple of 8 and at most 256; all sums are fp32 (callers zero them).
#include "psw_common.cuh"

namespace psw {

constexpr int BN_THREADS = 256;
constexpr int BN_MAXC = 256;

__device__ __forceinline__ void bn_unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
    f[2 * k] = __low2float(h);
    f[2 * k + 1] = __high2float(h);
  }
}
__device__ __forceinline__ uint4 bn_pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// a[c] += sum over the pixels of va, b[c] += sum of vb; every thread contributes 8 channels starting at ch0
__device__ __forceinline__ void bn_block_reduce(const float (&va)[8], const float (&vb)[8], int ch0, int C, float* ga, float* gb) {
  __shared__ float sa[BN_MAXC], sb[BN_MAXC];
  for (int c = threadIdx.x; c < C; c += BN_THREADS) { sa[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) { atomicAdd(&sa[ch0 + k], va[k]); atomicAdd(&sb[ch0 + k], vb[k]); }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BN_THREADS) { atomicAdd(ga + c, sa[c]); atomicAdd(gb + c, sb[c]); }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const bf16* __restrict__ x, float* __restrict__ sum, float* __restrict__ sumsq, int64_t npix, int C) {
  const int tpp = C / 8;                                     // threads per pixel
  const int ch0 = (threadIdx.x % tpp) * 8;
  const int ppb = BN_THREADS / tpp;                          // pixels per block and iteration
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
  if ((int)threadIdx.x < ppb * tpp) {
    for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp; p < npix; p += (int64_t)gridDim.x * ppb) {
      float f[8];
      bn_unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * C + ch0)), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
    }
  }
  bn_block_reduce(s, q, ch0, C, sum, sumsq);
}

// scale[c] = gamma * rstd, shift[c] = beta - mean * scale
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_relu_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ scale,
                     const float* __restrict__ shift, int64_t npix, int C) {
  const int tpp = C / 8;
  const int ch0 = (threadIdx.x % tpp) * 8;
  const int ppb = BN_THREADS / tpp;
  if ((int)threadIdx.x >= ppb * tpp) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[ch0 + k]; sh[k] = shift[ch0 + k]; }
  for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp; p < npix; p += (int64_t)gridDim.x * ppb) {
    float f[8];
    bn_unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * C + ch0)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
    *reinterpret_cast<uint4*>(y + p * C + ch0) = bn_pack8(f);
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_reduce_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const bf16* __restrict__ dy,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ sdy,
                     float* __restrict__ sdyx, int64_t npix, int C) {
  const int tpp = C / 8;
  const int ch0 = (threadIdx.x % tpp) * 8;
  const int ppb = BN_THREADS / tpp;
  float a[8], b[8], mu[8], rs[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; mu[k] = mean[ch0 + k]; rs[k] = rstd[ch0 + k]; }
  if ((int)threadIdx.x < ppb * tpp) {
    for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp; p < npix; p += (int64_t)gridDim.x * ppb) {
      float fx[8], fy[8], fg[8];
      bn_unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * C + ch0)), fx);
      bn_unpack8(__ldg(reinterpret_cast<const uint4*>(y + p * C + ch0)), fy);
      bn_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + p * C + ch0)), fg);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float g = fy[k] > 0.f ? fg[k] : 0.f;
        a[k] += g;
        b[k] = fmaf(g, (fx[k] - mu[k]) * rs[k], b[k]);
      }
    }
  }
  bn_block_reduce(a, b, ch0, C, sdy, sdyx);
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_dx_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, const bf16* __restrict__ dy,
                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                 const float* __restrict__ sdy, const float* __restrict__ sdyx, bf16* __restrict__ dx, int64_t npix, int C) {
  const int tpp = C / 8;
  const int ch0 = (threadIdx.x % tpp) * 8;
  const int ppb = BN_THREADS / tpp;
  if ((int)threadIdx.x >= ppb * tpp) return;
  const float inv_n = 1.0f / (float)npix;
  float mu[8], rs[8], gr[8], m1[8], m2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = mean[ch0 + k];
    rs[k] = rstd[ch0 + k];
    gr[k] = gamma[ch0 + k] * rs[k];
    m1[k] = sdy[ch0 + k] * inv_n;
    m2[k] = sdyx[ch0 + k] * inv_n;
  }
  for (int64_t p = (int64_t)blockIdx.x * ppb + threadIdx.x / tpp; p < npix; p += (int64_t)gridDim.x * ppb) {
    float fx[8], fy[8], fg[8], o[8];
    bn_unpack8(__ldg(reinterpret_cast<const uint4*>(x + p * C + ch0)), fx);
    bn_unpack8(__ldg(reinterpret_cast<const uint4*>(y + p * C + ch0)), fy);
    bn_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + p * C + ch0)), fg);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float g = fy[k] > 0.f ? fg[k] : 0.f;
      o[k] = gr[k] * (g - m1[k] - (fx[k] - mu[k]) * rs[k] * m2[k]);
    }
    *reinterpret_cast<uint4*>(dx + p * C + ch0) = bn_pack8(o);
  }
}

static int bn_grid(int64_t npix, int C) {
  const int ppb = BN_THREADS / (C / 8);
  int64_t blocks = (npix + ppb - 1) / ppb;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace psw

using namespace psw;

static int bn_check(const char* fn, int64_t npix, int C) {
  PSW_REQUIRE(npix > 0 && C > 0 && C % 8 == 0 && C <= BN_MAXC, PSW_ERR_UNSUPPORTED,
              "%s: needs 0 < C <= 256 and C %% 8 == 0 (pixels=%lld C=%d)", fn, (long long)npix, C);
  return 0;
}

extern "C" PSW_API int psw_bn_stats_fwd(const void* x, float* sum, float* sumsq, int64_t npix, int C, void* stream) {
  PSW_REQUIRE(x && sum && sumsq && aligned16(x), PSW_ERR_BAD_ARG, "psw_bn_stats_fwd: null or unaligned pointer");
  int rc = bn_check("psw_bn_stats_fwd", npix, C);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PSW_CUDA(cudaMemsetAsync(sum, 0, sizeof(float) * C, st));
  PSW_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * C, st));
  bn_stats_kernel<<<bn_grid(npix, C), BN_THREADS, 0, st>>>((const bf16*)x, sum, sumsq, npix, C);
  return launch_status("bn_stats_kernel");
}

extern "C" PSW_API int psw_bn_apply_relu_fwd(const void* x, void* y, const float* scale, const float* shift, int64_t npix, int C,
                                             void* stream) {
  PSW_REQUIRE(x && y && scale && shift && aligned16(x) && aligned16(y), PSW_ERR_BAD_ARG, "psw_bn_apply_relu_fwd: null or unaligned pointer");
  int rc = bn_check("psw_bn_apply_relu_fwd", npix, C);
  if (rc) return rc;
  bn_apply_relu_kernel<<<bn_grid(npix, C), BN_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, scale, shift, npix, C);
  return launch_status("bn_apply_relu_kernel");
}

extern "C" PSW_API int psw_bn_relu_bwd(const void* x, const void* y, const void* dy, const float* mean, const float* rstd,
                                       const float* gamma, void* dx, float* dgamma, float* dbeta, int64_t npix, int C,
                                       void* stream) {
  PSW_REQUIRE(x && y && dy && mean && rstd && gamma && dx && dgamma && dbeta, PSW_ERR_BAD_ARG, "psw_bn_relu_bwd: null pointer");
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(dy) && aligned16(dx), PSW_ERR_BAD_ARG, "psw_bn_relu_bwd: pointers must be 16-byte aligned");
  int rc = bn_check("psw_bn_relu_bwd", npix, C);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PSW_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * C, st));
  PSW_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * C, st));
  const int grid = bn_grid(npix, C);
  bn_bwd_reduce_kernel<<<grid, BN_THREADS, 0, st>>>((const bf16*)x, (const bf16*)y, (const bf16*)dy, mean, rstd, dbeta, dgamma, npix, C);
  rc = launch_status("bn_bwd_reduce_kernel");
  if (rc) return rc;
  bn_bwd_dx_kernel<<<grid, BN_THREADS, 0, st>>>((const bf16*)x, (const bf16*)y, (const bf16*)dy, mean, rstd, gamma, dbeta, dgamma,
                                               (bf16*)dx, npix, C);
  return launch_status("bn_bwd_dx_kernel");
}
