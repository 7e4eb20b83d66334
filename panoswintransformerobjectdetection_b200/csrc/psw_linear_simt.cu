// K2 (fp32 parity path): y = act(x . w^T + bias) (+ residual) on CUDA-core FMAs.
// Register-tiled 64x64x16 SGEMM, 256 threads, 4x4 outputs per thread; x [M,K] and w [N,K] are both
// K-contiguous (nn.Linear layout).  Exact fp32 FMA accumulation keeps the whole backbone within 1e-5 of
// the reference (BASELINE.json north_star, "fp32 mode"); the throughput path is psw_linear_tc.cu.
#include "psw_common.cuh"

namespace psw {

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

template <bool GELU>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ residual, float* __restrict__ y, int64_t M, int N, int K) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 64 rows x 16 k = 1024 elements, 4 per thread; k fastest for coalescing
  const int lk = tid & 15;
  const int lr = tid >> 4;   // 0..15, rows lr, lr+16, lr+32, lr+48
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int r = lr + 16 * p;
      int64_t gm = m0 + r;
      int gk = k0 + lk;
      As[lk][r] = (gm < M && gk < K) ? x[gm * K + gk] : 0.f;
      int gn = n0 + r;
      Bs[lk][r] = (gn < N && gk < K) ? w[(int64_t)gn * K + gk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j] + (bias ? bias[gn] : 0.f);
      if (GELU) v = gelu_erf(v);
      if (residual) v += residual[gm * N + gn];
      y[gm * N + gn] = v;
    }
  }
}

int linear_f32(const float* x, const float* w, const float* bias, const float* residual, float* y, int64_t M, int N,
               int K, int flags, cudaStream_t st) {
  dim3 grid((N + SG_BN - 1) / SG_BN, (unsigned)((M + SG_BM - 1) / SG_BM));
  // gridDim.y limit is 65535: fold large M onto x when needed
  if (grid.y > 65535u) {
    // process in slabs of 65535 row-tiles
    int64_t rows_per_slab = (int64_t)65535 * SG_BM;
    for (int64_t r0 = 0; r0 < M; r0 += rows_per_slab) {
      int64_t m = (M - r0 < rows_per_slab) ? (M - r0) : rows_per_slab;
      int rc = linear_f32(x + r0 * K, w, bias, residual ? residual + r0 * N : nullptr, y + r0 * N, m, N, K, flags, st);
      if (rc) return rc;
    }
    return 0;
  }
  if (flags & PSW_EPI_GELU)
    linear_f32_kernel<true><<<grid, 256, 0, st>>>(x, w, bias, residual, y, M, N, K);
  else
    linear_f32_kernel<false><<<grid, 256, 0, st>>>(x, w, bias, residual, y, M, N, K);
  return launch_status("linear_f32_kernel");
}

}  // namespace psw
