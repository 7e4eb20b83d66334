// Backward kernels of the bandwidth / GEMM ops of the PanoSwin block: LayerNorm and
// PatchMerging-LayerNorm gradients, GELU forward / backward for the training path (which keeps the pre-activation),
// the linear layer's input / weight / bias gradients, and a 2-D transpose.  The reference gets all of these from
// torch autograd (simple_panoswin_transformer.py has no backward code); formulas are the standard ones:
//   LayerNorm (:504, :534, :574, :768-772, :975-976)   xh = (x - mean) * rstd,  g = dy * gamma
//       dx = rstd * (g - mean(g) - xh * mean(g * xh));  dgamma = sum_rows dy * xh;  dbeta = sum_rows dy
//   Linear (:287, :309, :55-61, :575)                   dx = dy . W;  dW = dy^T . x;  db = sum_rows dy
//   GELU (nn.GELU, exact erf)                           d/dx = Phi(x) + x * phi(x)
// fp32 storage runs on CUDA-core FMAs (parity path, <= 1e-4 of torch autograd); bf16 storage sends dx through the
// tcgen05 GEMM (psw_linear_fwd on the transposed weight) and accumulates dW in fp32 from the bf16 operands.
#include "psw_common.cuh"

namespace psw {

// ---------------------------------------------------------------------------------------------------------------
// row address map: plain rows [rows, C], or the 2x2 PatchMerging gather (reference :563-573): merged row (b, i2, j2),
// column c = q * Cin + cc with quadrant q -> (dh, dw) = (q & 1, q >> 1); cells beyond an odd H / W are zero padding
// ---------------------------------------------------------------------------------------------------------------
struct RowMap {
  int merge;
  int C;                        // row width (4 * Cin when merge)
  int H, W, Cin, H2, W2;
};

__device__ __forceinline__ int64_t row_offset(const RowMap& m, int64_t r, int c) {
  if (!m.merge) return r * m.C + c;
  const int per = m.H2 * m.W2;
  const int b = (int)(r / per);
  const int rem = (int)(r - (int64_t)b * per);
  const int i2 = rem / m.W2, j2 = rem - i2 * m.W2;
  const int q = c / m.Cin, cc = c - q * m.Cin;
  const int h = 2 * i2 + (q & 1), w = 2 * j2 + (q >> 1);
  if (h >= m.H || w >= m.W) return -1;
  return (((int64_t)b * m.H + h) * m.W + w) * m.Cin + cc;
}

template <typename T> __device__ __forceinline__ void load4_or_zero(const T* base, int64_t off, float (&v)[4]) {
  if (off >= 0) load4(base + off, v);
  else v[0] = v[1] = v[2] = v[3] = 0.f;
}

// One warp per row, four passes over the row (it is re-read from L1 / L2): mean, centred variance, the two
// reductions of the gradient, dx.  stats[2r] = mean, stats[2r + 1] = rstd feed the parameter-gradient kernel.
// KP > 0 (C <= 128 KP): the parameter gradients ride along -- every lane keeps the partial sums of its 4 KP columns over
// the rows of its warp in registers, the eight warps of a CTA meet in shared memory, one atomic per column and CTA:
// dgamma / dbeta (zero on entry) need no second pass over x and dy.
template <typename TX, typename TDY, typename TDX, int KP>
__global__ void __launch_bounds__(256)
ln_bwd_rows_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ gamma,
                   TDX* __restrict__ dx, float* __restrict__ stats, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   RowMap m, int64_t rows, float eps) {
  constexpr int KQ = KP > 0 ? KP : 1;
  __shared__ float s_pg[KP > 0 ? 128 * KP : 1], s_pb[KP > 0 ? 128 * KP : 1];
  float pg[4 * KQ], pb[4 * KQ];
#pragma unroll
  for (int k = 0; k < 4 * KQ; ++k) { pg[k] = 0.f; pb[k] = 0.f; }
  if (KP > 0) {
    for (int c = threadIdx.x; c < 128 * KP; c += 256) { s_pg[c] = 0.f; s_pb[c] = 0.f; }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int C = m.C;
  const float inv_c = 1.0f / (float)C;
  for (int64_t r = warp0; r < rows; r += (int64_t)gridDim.x * 8) {
    float s = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float v[4];
      load4_or_zero(x, row_offset(m, r, c), v);
      s += (v[0] + v[1]) + (v[2] + v[3]);
    }
    const float mean = warp_sum(s) * inv_c;
    float ss = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float v[4];
      load4_or_zero(x, row_offset(m, r, c), v);
#pragma unroll
      for (int k = 0; k < 4; ++k) { const float d = v[k] - mean; ss = fmaf(d, d, ss); }
    }
    const float rstd = rsqrtf(warp_sum(ss) * inv_c + eps);
    float a = 0.f, b = 0.f;
    for (int c = lane * 4; c < C; c += 128) {
      float v[4], g[4], w[4];
      load4_or_zero(x, row_offset(m, r, c), v);
      load4(dy + r * C + c, g);
      load4(gamma + c, w);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gg = g[k] * w[k];
        a += gg;
        b = fmaf(gg, (v[k] - mean) * rstd, b);
      }
    }
    a = warp_sum(a) * inv_c;
    b = warp_sum(b) * inv_c;
    if (KP > 0) {
#pragma unroll
      for (int q = 0; q < KQ; ++q) {
        const int c = lane * 4 + 128 * q;
        if (c < C) {
          const int64_t off = row_offset(m, r, c);
          float v[4], g[4], w[4], o[4];
          load4_or_zero(x, off, v);
          load4(dy + r * C + c, g);
          load4(gamma + c, w);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float xh = (v[k] - mean) * rstd;
            o[k] = rstd * (g[k] * w[k] - a - xh * b);
            pg[4 * q + k] = fmaf(g[k], xh, pg[4 * q + k]);
            pb[4 * q + k] += g[k];
          }
          if (off >= 0) store4(dx + off, o);                // zero padding of an odd map: no gradient to store
        }
      }
    } else {
      for (int c = lane * 4; c < C; c += 128) {
        const int64_t off = row_offset(m, r, c);
        if (off < 0) continue;                              // zero padding of an odd map: no gradient to store
        float v[4], g[4], w[4], o[4];
        load4(x + off, v);
        load4(dy + r * C + c, g);
        load4(gamma + c, w);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = rstd * (g[k] * w[k] - a - (v[k] - mean) * rstd * b);
        store4(dx + off, o);
      }
    }
    if (lane == 0 && stats) { stats[2 * r] = mean; stats[2 * r + 1] = rstd; }
  }
  if (KP > 0) {
#pragma unroll
    for (int q = 0; q < KQ; ++q) {
      const int c = lane * 4 + 128 * q;
      if (c < C) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { atomicAdd(&s_pg[c + k], pg[4 * q + k]); atomicAdd(&s_pb[c + k], pb[4 * q + k]); }
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) { atomicAdd(dgamma + c, s_pg[c]); atomicAdd(dbeta + c, s_pb[c]); }
  }
}

// dgamma[c] += sum_r dy[r,c] * xh[r,c], dbeta[c] += sum_r dy[r,c]: a 32-column x ROWS-row slab per block, one atomic
// per column and block
constexpr int LNP_ROWS = 512;
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
ln_bwd_params_kernel(const TX* __restrict__ x, const TDY* __restrict__ dy, const float* __restrict__ stats,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, RowMap m, int64_t rows) {
  __shared__ float sg[8][33], sb[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t r0 = (int64_t)blockIdx.y * LNP_ROWS;
  const int64_t r1 = r0 + LNP_ROWS < rows ? r0 + LNP_ROWS : rows;
  float g = 0.f, b = 0.f;
  if (c < m.C) {
    for (int64_t r = r0 + ty; r < r1; r += 8) {
      const int64_t off = row_offset(m, r, c);
      const float xv = off >= 0 ? to_f32(x[off]) : 0.f;
      const float d = to_f32(dy[r * m.C + c]);
      g = fmaf(d, (xv - stats[2 * r]) * stats[2 * r + 1], g);
      b += d;
    }
  }
  sg[ty][tx] = g;
  sb[ty][tx] = b;
  __syncthreads();
  if (ty == 0 && c < m.C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { g += sg[k][tx]; b += sb[k][tx]; }
    atomicAdd(dgamma + c, g);
    atomicAdd(dbeta + c, b);
  }
}

template <typename TX, typename TDY, typename TDX>
static int ln_bwd_launch(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma, float* dbeta,
                         float* stats, const RowMap& m, int64_t rows, float eps, cudaStream_t st) {
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (dgamma && m.C <= 1024) {                              // parameter gradients inside the row pass
    PSW_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * m.C, st));
    PSW_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * m.C, st));
    const int64_t cap2 = (int64_t)num_sms() * 4;            // fewer, longer-lived CTAs: one atomic per column and CTA
    if (blocks > cap2) blocks = cap2;
    const int kp = (m.C + 127) / 128;
#define PSW_LN_BWD_KP(KP_) \
    ln_bwd_rows_kernel<TX, TDY, TDX, KP_><<<(unsigned)blocks, 256, 0, st>>>((const TX*)x, (const TDY*)dy, gamma, (TDX*)dx, stats, \
                                                                           dgamma, dbeta, m, rows, eps)
    if (kp <= 1) PSW_LN_BWD_KP(1);
    else if (kp <= 2) PSW_LN_BWD_KP(2);
    else if (kp <= 3) PSW_LN_BWD_KP(3);
    else if (kp <= 6) PSW_LN_BWD_KP(6);
    else PSW_LN_BWD_KP(8);
#undef PSW_LN_BWD_KP
    return launch_status("ln_bwd_rows_kernel");
  }
  ln_bwd_rows_kernel<TX, TDY, TDX, 0><<<(unsigned)blocks, 256, 0, st>>>((const TX*)x, (const TDY*)dy, gamma, (TDX*)dx, stats, nullptr,
                                                                         nullptr, m, rows, eps);
  int rc = launch_status("ln_bwd_rows_kernel");
  if (rc || !dgamma) return rc;
  PSW_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * m.C, st));
  PSW_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * m.C, st));
  dim3 grid((m.C + 31) / 32, (unsigned)((rows + LNP_ROWS - 1) / LNP_ROWS));
  ln_bwd_params_kernel<TX, TDY><<<grid, 256, 0, st>>>((const TX*)x, (const TDY*)dy, stats, dgamma, dbeta, m, rows);
  return launch_status("ln_bwd_params_kernel");
}

static int ln_bwd_dispatch(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma, float* dbeta,
                           float* stats, const RowMap& m, int64_t rows, float eps, int x_dtype, int dy_dtype,
                           cudaStream_t st) {
  // dx has the storage type of x (it is the gradient of that tensor)
  if (x_dtype == PSW_F32 && dy_dtype == PSW_F32) return ln_bwd_launch<float, float, float>(x, dy, gamma, dx, dgamma, dbeta, stats, m, rows, eps, st);
  if (x_dtype == PSW_F32 && dy_dtype == PSW_BF16) return ln_bwd_launch<float, bf16, float>(x, dy, gamma, dx, dgamma, dbeta, stats, m, rows, eps, st);
  if (x_dtype == PSW_BF16 && dy_dtype == PSW_BF16) return ln_bwd_launch<bf16, bf16, bf16>(x, dy, gamma, dx, dgamma, dbeta, stats, m, rows, eps, st);
  if (x_dtype == PSW_BF16 && dy_dtype == PSW_F32) return ln_bwd_launch<bf16, float, bf16>(x, dy, gamma, dx, dgamma, dbeta, stats, m, rows, eps, st);
  PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "layernorm backward: unknown dtype %d / %d", x_dtype, dy_dtype);
  return PSW_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------------------
// GELU (exact erf), training path: y = gelu(h) keeps h for the backward; dh = dy * (Phi(h) + h * phi(h))
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool BWD>
__global__ void gelu_kernel(const T* __restrict__ h, const T* __restrict__ dy, T* __restrict__ out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float v[4], g[4], o[4];
    load4(h + 4 * i, v);
    if (BWD) load4(dy + 4 * i, g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (BWD) {
        const float cdf = 0.5f * (1.0f + erff(v[k] * 0.70710678118654752440f));
        const float pdf = 0.39894228040143267794f * expf(-0.5f * v[k] * v[k]);
        o[k] = g[k] * (cdf + v[k] * pdf);
      } else {
        o[k] = gelu_erf(v[k]);
      }
    }
    store4(out + 4 * i, o);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// C[i][j] (+)= sum_l A(i, l) * B(j, l) on CUDA-core FMAs, fp32 accumulation, arbitrary element strides:
//   A(i, l) = A[i * sai + l * sal],  B(j, l) = B[j * sbj + l * sbl] (B is j-fast: sbj == 1).
// 64 x 64 x 16 tiles, 4 x 4 outputs per thread; the l range can be split over gridDim.z with fp32 atomics (weight
// gradients: the contraction runs over all rows of the batch while the output is a small [N, K] matrix).
//   dx[M,K] = dy[M,N] . W[N,K]      i = m, j = k, l = n : A = dy (l-fast), B = W (j-fast)
//   dW[N,K] = dy[M,N]^T . x[M,K]    i = n, j = k, l = m : A = dy (i-fast), B = x (j-fast), split over m
// ---------------------------------------------------------------------------------------------------------------
template <typename TA, typename TB, typename TC, bool A_IFAST>
__global__ void __launch_bounds__(256)
gemm_strided_kernel(const TA* __restrict__ A, const TB* __restrict__ B, TC* __restrict__ Cout, int64_t I, int J, int64_t L,
                    int64_t sai, int64_t sal, int64_t sbl, int ldc, int64_t l_per_split) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t i0 = (int64_t)blockIdx.x * 64;
  const int j0 = blockIdx.y * 64;
  const int64_t lbeg = (int64_t)blockIdx.z * l_per_split;
  const int64_t lend = lbeg + l_per_split < L ? lbeg + l_per_split : L;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int64_t l0 = lbeg; l0 < lend; l0 += 16) {
    if (A_IFAST) {
      const int ii = tid & 63, ll = tid >> 6;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int64_t l = l0 + ll + 4 * p, i = i0 + ii;
        As[ll + 4 * p][ii] = (i < I && l < lend) ? to_f32(A[i * sai + l * sal]) : 0.f;
      }
    } else {
      const int lk = tid & 15, lr = tid >> 4;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int64_t l = l0 + lk, i = i0 + lr + 16 * p;
        As[lk][lr + 16 * p] = (i < I && l < lend) ? to_f32(A[i * sai + l * sal]) : 0.f;
      }
    }
    {
      const int jj = tid & 63, ll = tid >> 6;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int64_t l = l0 + ll + 4 * p;
        const int j = j0 + jj;
        Bs[ll + 4 * p][jj] = (j < J && l < lend) ? to_f32(B[(int64_t)j + l * sbl]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = As[kk][ty * 4 + q];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = Bs[kk][tx * 4 + q];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int w = 0; w < 4; ++w) acc[q][w] = fmaf(a[q], b[w], acc[q][w]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t i = i0 + ty * 4 + q;
    if (i >= I) continue;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int j = j0 + tx * 4 + w;
      if (j >= J) continue;
      if constexpr (sizeof(TC) == 4) {
        if (gridDim.z > 1) atomicAdd(reinterpret_cast<float*>(Cout) + i * ldc + j, acc[q][w]);
        else Cout[i * ldc + j] = from_f32<TC>(acc[q][w]);
      } else {
        Cout[i * ldc + j] = from_f32<TC>(acc[q][w]);
      }
    }
  }
}

// db[n] = sum_m dy[m, n] (N % 4 == 0): a 128-column x rows_per_block slab per block, a lane owns four adjacent columns
// (one 8- or 16-byte load per row), the eight warps take every eighth row; one atomic per column and block
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ dy, float* __restrict__ db, int64_t M, int N, int rows_per_block) {
  __shared__ float sb[8][132];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < N) {
#pragma unroll 4
    for (int64_t r = r0 + w; r < r1; r += 8) {
      float v[4];
      load4(dy + r * N + c, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] += v[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) sb[w][lane * 4 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sb[k][threadIdx.x];
    atomicAdd(db + blockIdx.x * 128 + threadIdx.x, t);
  }
}

// dst[c][r] = src[r][c]
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t Cc) {
  __shared__ T tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < R && c < Cc) tile[ty + 8 * k][tx] = src[r * Cc + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < R && c < Cc) dst[c * R + r] = tile[tx][ty + 8 * k];
  }
}

template <typename T, typename TDX>
static int linear_bwd_simt(const T* x, const T* w, const T* dy, TDX* dx, float* dw, float* db, int64_t M, int N, int K,
                           cudaStream_t st) {
  if (dx) {                                                 // dx[M,K] = dy[M,N] . W[N,K]
    dim3 grid((unsigned)((M + 63) / 64), (K + 63) / 64, 1);
    gemm_strided_kernel<T, T, TDX, false><<<grid, 256, 0, st>>>(dy, w, dx, M, K, N, N, 1, K, K, N);
    int rc = launch_status("gemm_strided_kernel(dgrad)");
    if (rc) return rc;
  }
  if (dw) {                                                 // dW[N,K] = dy^T . x, the row dimension split over the grid
    const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
    int64_t splits = (4ll * num_sms() + tiles - 1) / tiles;
    const int64_t max_splits = (M + 255) / 256;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t per = ((M + splits - 1) / splits + 15) / 16 * 16;
    splits = (M + per - 1) / per;
    PSW_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    dim3 grid((N + 63) / 64, (K + 63) / 64, (unsigned)splits);
    gemm_strided_kernel<T, T, float, true><<<grid, 256, 0, st>>>(dy, x, dw, N, K, M, 1, N, K, K, per);
    int rc = launch_status("gemm_strided_kernel(wgrad)");
    if (rc) return rc;
  }
  if (db) {
    PSW_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
    PSW_REQUIRE(N % 4 == 0, PSW_ERR_UNSUPPORTED, "psw_linear_bwd: the bias gradient needs N %% 4 == 0 (N=%d)", N);
    const int col_groups = (N + 127) / 128;
    int64_t rpb = M * col_groups / (8ll * num_sms());       // ~8 blocks per SM in flight, 64 .. 2048 rows each
    rpb = rpb < 64 ? 64 : (rpb > 2048 ? 2048 : rpb);
    rpb = (rpb + 7) / 8 * 8;
    dim3 grid(col_groups, (unsigned)((M + rpb - 1) / rpb));
    colsum_kernel<T><<<grid, 256, 0, st>>>(dy, db, M, N, (int)rpb);
    return launch_status("colsum_kernel");
  }
  return 0;
}

int wgrad_tc(const bf16* dy, const bf16* x, float* dw, int64_t M, int N, int K, cudaStream_t st);   // psw_wgrad_tc.cu

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_layernorm_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma,
                                         float* dbeta, float* stats_ws, int64_t rows, int C, float eps, int x_dtype,
                                         int dy_dtype, void* stream) {
  PSW_REQUIRE(x && dy && gamma && dx && stats_ws, PSW_ERR_BAD_ARG, "psw_layernorm_bwd: null pointer");
  PSW_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), PSW_ERR_BAD_ARG, "psw_layernorm_bwd: dgamma and dbeta go together");
  PSW_REQUIRE(rows > 0 && C > 0 && C % 4 == 0, PSW_ERR_BAD_ARG, "psw_layernorm_bwd: rows=%lld C=%d (C %% 4 == 0)", (long long)rows, C);
  PSW_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma), PSW_ERR_BAD_ARG,
              "psw_layernorm_bwd: pointers must be 16-byte aligned");
  const RowMap m = {0, C, 0, 0, 0, 0, 0};
  return ln_bwd_dispatch(x, dy, gamma, dx, dgamma, dbeta, stats_ws, m, rows, eps, x_dtype, dy_dtype, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_patch_merge_ln_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma,
                                              float* dbeta, float* stats_ws, int B, int H, int W, int C, float eps,
                                              int x_dtype, int dy_dtype, void* stream) {
  PSW_REQUIRE(x && dy && gamma && dx && stats_ws, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_bwd: null pointer");
  PSW_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), PSW_ERR_BAD_ARG, "psw_patch_merge_ln_bwd: dgamma and dbeta go together");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_bwd: bad dims");
  PSW_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(gamma), PSW_ERR_BAD_ARG,
              "psw_patch_merge_ln_bwd: pointers must be 16-byte aligned");
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const RowMap m = {1, 4 * C, H, W, C, H2, W2};
  const int64_t rows = (int64_t)B * H2 * W2;
  // cells of the input that no merged row covers do not exist (every cell of an H x W map is covered); the gradient of
  // every real cell is written exactly once, so dx needs no zero fill
  return ln_bwd_dispatch(x, dy, gamma, dx, dgamma, dbeta, stats_ws, m, rows, eps, x_dtype, dy_dtype, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_gelu_fwd(const void* h, void* y, int64_t n, int dtype, void* stream) {
  PSW_REQUIRE(h && y && n > 0 && n % 4 == 0, PSW_ERR_BAD_ARG, "psw_gelu_fwd: null pointer or n %% 4 != 0");
  PSW_REQUIRE(aligned16(h) && aligned16(y), PSW_ERR_BAD_ARG, "psw_gelu_fwd: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32) gelu_kernel<float, false><<<(unsigned)blocks, 256, 0, st>>>((const float*)h, nullptr, (float*)y, n4);
  else if (dtype == PSW_BF16) gelu_kernel<bf16, false><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)h, nullptr, (bf16*)y, n4);
  else PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_gelu_fwd: unknown dtype %d", dtype);
  return launch_status("gelu_kernel");
}

extern "C" PSW_API int psw_gelu_bwd(const void* h, const void* dy, void* dh, int64_t n, int dtype, void* stream) {
  PSW_REQUIRE(h && dy && dh && n > 0 && n % 4 == 0, PSW_ERR_BAD_ARG, "psw_gelu_bwd: null pointer or n %% 4 != 0");
  PSW_REQUIRE(aligned16(h) && aligned16(dy) && aligned16(dh), PSW_ERR_BAD_ARG, "psw_gelu_bwd: pointers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + 255) / 256;
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32) gelu_kernel<float, true><<<(unsigned)blocks, 256, 0, st>>>((const float*)h, (const float*)dy, (float*)dh, n4);
  else if (dtype == PSW_BF16) gelu_kernel<bf16, true><<<(unsigned)blocks, 256, 0, st>>>((const bf16*)h, (const bf16*)dy, (bf16*)dh, n4);
  else PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_gelu_bwd: unknown dtype %d", dtype);
  return launch_status("gelu_kernel");
}

extern "C" PSW_API int psw_transpose(const void* src, void* dst, int64_t rows, int64_t cols, int dtype, void* stream) {
  PSW_REQUIRE(src && dst && rows > 0 && cols > 0, PSW_ERR_BAD_ARG, "psw_transpose: bad arguments");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
  PSW_REQUIRE(grid.y <= 65535u, PSW_ERR_UNSUPPORTED, "psw_transpose: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32) transpose_kernel<float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, rows, cols);
  else if (dtype == PSW_BF16) transpose_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)src, (bf16*)dst, rows, cols);
  else PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_transpose: unknown dtype %d", dtype);
  return launch_status("transpose_kernel");
}

extern "C" PSW_API int64_t psw_linear_bwd_workspace_bytes(int64_t M, int N, int K, int dtype) {
  (void)M;
  return dtype == PSW_BF16 ? (int64_t)N * K * 2 : 0;        // the transposed bf16 weight for the tcgen05 dgrad
}

extern "C" PSW_API int psw_linear_bwd(const void* x, const void* w, const void* dy, void* dx, float* dw, float* db,
                                      int64_t M, int N, int K, int dtype, int dx_dtype, void* workspace,
                                      int64_t workspace_bytes, void* stream) {
  PSW_REQUIRE(w && dy, PSW_ERR_BAD_ARG, "psw_linear_bwd: null pointer");
  PSW_REQUIRE(M > 0 && N > 0 && K > 0, PSW_ERR_BAD_ARG, "psw_linear_bwd: M=%lld N=%d K=%d", (long long)M, N, K);
  PSW_REQUIRE(!dw || x, PSW_ERR_BAD_ARG, "psw_linear_bwd: the weight gradient needs x");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32) {
    PSW_REQUIRE(dx_dtype == PSW_F32, PSW_ERR_BAD_ARG, "psw_linear_bwd: fp32 path writes an fp32 dx");
    return linear_bwd_simt<float, float>((const float*)x, (const float*)w, (const float*)dy, (float*)dx, dw, db, M, N, K, st);
  }
  PSW_REQUIRE(dtype == PSW_BF16, PSW_ERR_BAD_ARG, "psw_linear_bwd: unknown dtype %d", dtype);
  PSW_REQUIRE(dx_dtype == PSW_BF16 || dx_dtype == PSW_F32, PSW_ERR_BAD_ARG, "psw_linear_bwd: unknown dx_dtype %d", dx_dtype);
  void* dx_simt = dx;
  if (dx && N % 8 == 0 && K % 16 == 0 && workspace && workspace_bytes >= (int64_t)N * K * 2 && aligned16(workspace) &&
      aligned16(dy) && aligned16(dx)) {
    // dx = dy . W on the tcgen05 GEMM: psw_linear_fwd(x' = dy [M, N], w' = W^T [K, N]) -> [M, K]
    int rc = psw_transpose(w, workspace, N, K, PSW_BF16, stream);
    if (rc) return rc;
    rc = psw_linear_fwd(dy, workspace, nullptr, nullptr, dx, M, K, N, 0, PSW_BF16, dx_dtype, stream);
    if (rc) return rc;
    dx_simt = nullptr;
  }
  if (dw && N % 8 == 0 && K % 8 == 0 && aligned16(dy) && aligned16(x) && M >= 64) {
    // dW = dy^T . x on tcgen05 with MN-major operands straight from the row-major activations (psw_wgrad_tc.cu)
    PSW_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * K, st));
    int rc = wgrad_tc((const bf16*)dy, (const bf16*)x, dw, M, N, K, st);
    if (rc) return rc;
    dw = nullptr;
  }
  if (!dx_simt && !dw && !db) return 0;
  if (dx_dtype == PSW_F32)
    return linear_bwd_simt<bf16, float>((const bf16*)x, (const bf16*)w, (const bf16*)dy, (float*)dx_simt, dw, db, M, N, K, st);
  return linear_bwd_simt<bf16, bf16>((const bf16*)x, (const bf16*)w, (const bf16*)dy, (bf16*)dx_simt, dw, db, M, N, K, st);
}
