// Backward of the fused pano window attention (psw_window_attn_fwd / psw_window_attn_full_fwd), SURVEY.md §8 f-3.
// The reference gets it from autograd through roll / flip / cat / pad / window_partition / softmax / matmul
// (simple_panoswin_transformer.py:274-311, :376-409, :473-519); here one CTA per (window, head) recomputes the
// probabilities from q, k, v (nothing but qkv is kept from the forward) and produces, in one pass over HBM:
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - rowsum(dO o O)),  dq = scale * dS K,  dK = dS^T (scale * q)
//   d alpha[idx(i,j), head] += hav(i,j) * dS[i,j]     d beta[idx(i,j), head] += dS[i,j]      (_sphere_bias :241-260)
// dq / dk / dv go to the token's UN-shifted position (the shift, partition, reverse and crop are the same address
// arithmetic as in the forward, psw::source_token); padding cells -- zero tokens whose q / k / v equal the qkv bias
// (:486-491, :344-347) -- send their gradient to d qkv_bias instead.  The table gradients are reduced per CTA in
// shared memory and leave with one atomic per table entry; uv and the mask carry no gradient.
// CUDA-core kernel for fp32 (parity path) and bf16 storage, any window size / head_dim; bf16 with window 7 / head_dim 32
// takes the tensor-core kernel of psw_attn_bwd_mma.cu.
#include "psw_common.cuh"

namespace psw {

template <typename T>
__global__ void __launch_bounds__(128)
window_attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ alpha,
                       const float* __restrict__ beta, const float* __restrict__ qkv_bias, const float* __restrict__ uv,
                       const float* __restrict__ mask, T* __restrict__ dqkv, float* __restrict__ dalpha,
                       float* __restrict__ dbeta, float* __restrict__ dqkv_bias, WinGeom g, int C, int heads, float scale) {
  extern __shared__ float sm[];
  const int ws = g.ws;
  const int N = ws * ws;
  const int hd = C / heads;
  const int SP = N + 1;
  const int hp = hd + 1;                                  // row pitch of q / k / v / dO: lanes that walk over tokens at a
                                                          // fixed dim hit different banks (pitch hd = 32 is a 32-way conflict)
  const int tw = 2 * ws - 1, TAB = tw * tw;
  float* sq = sm;                                         // [N][hp]  q * scale
  float* sk = sq + N * hp;                                // [N][hp]
  float* sv = sk + N * hp;                                // [N][hp]
  float* sdo = sv + N * hp;                               // [N][hp]  dO (zero on padding cells: their rows are cropped)
  float* sP = sdo + N * hp;                               // [N][N+1] logits -> P -> dS
  float* su = sP + N * SP;                                // [N]
  float* sw = su + N;                                     // [N]
  float* sdelta = sw + N;                                 // [N] rowsum(dO o O)
  float* sta = sdelta + N;                                // [TAB] d alpha of this (window, head)
  float* stb = sta + TAB;                                 // [TAB] d beta
  int* ssrc = reinterpret_cast<int*>(stb + TAB);          // [N]

  const int tid = threadIdx.x;
  const int e = blockIdx.x % heads;
  const int win = blockIdx.x / heads;
  const int wpi = g.nWh * g.nWw;
  const int b = win / wpi;
  const int wi = win - b * wpi;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  const int64_t HW = (int64_t)g.H * g.W;

  for (int t = tid; t < N; t += blockDim.x) {
    const int r = t / ws, c = t - r * ws;
    const int s = source_token(g, wr * ws + r, wc * ws + c);
    ssrc[t] = s;
    float uu = 0.f, vv = 0.f;
    if (g.pano && s >= 0) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
    su[t] = uu;
    sw[t] = vv;
    sdelta[t] = 0.f;
  }
  for (int t = tid; t < 2 * TAB; t += blockDim.x) sta[t] = 0.f;          // sta and stb are contiguous
  __syncthreads();

  for (int idx = tid; idx < 4 * N * hd; idx += blockDim.x) {
    const int part = idx / (N * hd);
    const int rem = idx - part * N * hd;
    const int t = rem / hd, d = rem - t * hd;
    const int s = ssrc[t];
    float val;
    if (part < 3) {
      const int ch = part * C + e * hd + d;
      if (s >= 0) val = to_f32(qkv[((int64_t)b * HW + s) * (3 * C) + ch]);
      else        val = qkv_bias ? qkv_bias[ch] : 0.f;
      if (part == 0) val *= scale;
    } else {
      val = s >= 0 ? to_f32(dout[((int64_t)b * HW + s) * C + e * hd + d]) : 0.f;
    }
    sm[(part * N + t) * hp + d] = val;                      // sq, sk, sv, sdo are contiguous
  }
  __syncthreads();

  // logits (same evaluation order as the forward parity kernel)
  for (int p = tid; p < N * N; p += blockDim.x) {
    const int i = p / N, j = p - i * N;
    float dot = 0.f;
    for (int d = 0; d < hd; ++d) dot = fmaf(sq[i * hp + d], sk[j * hp + d], dot);
    const int ri = i / ws, ci = i - ri * ws, rj = j / ws, cj = j - rj * ws;
    const int idx = (ri - rj + ws - 1) * tw + (ci - cj + ws - 1);
    float bia = beta[idx * heads + e];
    if (g.pano) {
      const float sdv = sinf(0.5f * fabsf(sw[j] - sw[i]));
      const float sdu = sinf(0.5f * (su[j] - su[i]));
      const float a = sdv * sdv + (cosf(sw[j]) * cosf(sw[i])) * (sdu * sdu);
      bia = asinf(sqrtf(a)) * 2.0f * alpha[idx * heads + e] + bia;
    }
    float s = dot + bia;
    if (mask) s += mask[((int64_t)wi * N + i) * N + j];
    sP[i * SP + j] = s;
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarps) {
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, sP[i * SP + j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      const float ex = expf(sP[i * SP + j] - m);
      sP[i * SP + j] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) sP[i * SP + j] *= inv;
  }
  __syncthreads();

  // dV[j][d] = sum_i P[i][j] dO[i][d];  delta[i] = sum_d dO[i][d] * O[i][d] with O = P V
  for (int idx = tid; idx < N * hd; idx += blockDim.x) {
    const int t = idx / hd, d = idx - t * hd;
    float dv = 0.f, o = 0.f;
    for (int i = 0; i < N; ++i) dv = fmaf(sP[i * SP + t], sdo[i * hp + d], dv);
    for (int j = 0; j < N; ++j) o = fmaf(sP[t * SP + j], sv[j * hp + d], o);
    atomicAdd(&sdelta[t], o * sdo[t * hp + d]);
    const int s = ssrc[t];
    const int ch = 2 * C + e * hd + d;
    if (s >= 0) dqkv[((int64_t)b * HW + s) * (3 * C) + ch] = from_f32<T>(dv);
    else if (dqkv_bias) atomicAdd(dqkv_bias + ch, dv);
  }
  __syncthreads();

  // dS in place of P; table gradients into the per-CTA accumulators
  for (int p = tid; p < N * N; p += blockDim.x) {
    const int i = p / N, j = p - i * N;
    float dp = 0.f;
    for (int d = 0; d < hd; ++d) dp = fmaf(sdo[i * hp + d], sv[j * hp + d], dp);
    const float ds = sP[i * SP + j] * (dp - sdelta[i]);
    sP[i * SP + j] = ds;
    const int ri = i / ws, ci = i - ri * ws, rj = j / ws, cj = j - rj * ws;
    const int idx = (ri - rj + ws - 1) * tw + (ci - cj + ws - 1);
    atomicAdd(&stb[idx], ds);
    if (g.pano) {
      const float sdv = sinf(0.5f * fabsf(sw[j] - sw[i]));
      const float sdu = sinf(0.5f * (su[j] - su[i]));
      const float a = sdv * sdv + (cosf(sw[j]) * cosf(sw[i])) * (sdu * sdu);
      atomicAdd(&sta[idx], asinf(sqrtf(a)) * 2.0f * ds);
    }
  }
  __syncthreads();

  // dq[i][d] = scale * sum_j dS[i][j] k[j][d];  dk[j][d] = sum_i dS[i][j] (scale * q[i][d])
  for (int idx = tid; idx < N * hd; idx += blockDim.x) {
    const int t = idx / hd, d = idx - t * hd;
    float dq = 0.f, dk = 0.f;
    for (int j = 0; j < N; ++j) dq = fmaf(sP[t * SP + j], sk[j * hp + d], dq);
    for (int i = 0; i < N; ++i) dk = fmaf(sP[i * SP + t], sq[i * hp + d], dk);
    dq *= scale;
    const int s = ssrc[t];
    const int ch = e * hd + d;
    if (s >= 0) {
      T* row = dqkv + ((int64_t)b * HW + s) * (3 * C);
      row[ch] = from_f32<T>(dq);
      row[C + ch] = from_f32<T>(dk);
    } else if (dqkv_bias) {
      atomicAdd(dqkv_bias + ch, dq);
      atomicAdd(dqkv_bias + C + ch, dk);
    }
  }
  for (int t = tid; t < TAB; t += blockDim.x) {
    if (g.pano && dalpha) atomicAdd(dalpha + t * heads + e, sta[t]);
    if (dbeta) atomicAdd(dbeta + t * heads + e, stb[t]);
  }
}

template <typename T>
static int window_attn_bwd(const T* qkv, const T* dout, const float* alpha, const float* beta, const float* qkv_bias,
                           const float* uv, const float* mask, T* dqkv, float* dalpha, float* dbeta, float* dqkv_bias,
                           int B, int H, int W, int C, int heads, int window, int shift, int pano, float scale,
                           cudaStream_t st) {
  WinGeom g = make_geom(H, W, window, shift, pano);
  const int N = window * window, hd = C / heads, TAB = (2 * window - 1) * (2 * window - 1);
  const size_t smem = ((size_t)4 * N * (hd + 1) + (size_t)N * (N + 1) + 4 * (size_t)N + 2 * (size_t)TAB) * sizeof(float);
  PSW_REQUIRE(smem <= 220 * 1024, PSW_ERR_UNSUPPORTED, "psw_window_attn_bwd: window %d x head_dim %d needs %zu B smem", window, hd, smem);
  const int64_t blocks = (int64_t)B * g.nWh * g.nWw * heads;
  PSW_REQUIRE(blocks < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_bwd: too many windows");
  PSW_CUDA(cudaFuncSetAttribute(window_attn_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  window_attn_bwd_kernel<T><<<(unsigned)blocks, 128, smem, st>>>(qkv, dout, alpha, beta, qkv_bias, uv, mask, dqkv, dalpha, dbeta,
                                                               dqkv_bias, g, C, heads, scale);
  return launch_status("window_attn_bwd_kernel");
}

int window_attn_bwd_mma(const bf16* qkv, const bf16* dout, const float* alpha, const float* beta, const float* qkv_bias,
                        const float* uv, const float* mask, bf16* dqkv, float* dalpha, float* dbeta, float* dqkv_bias, int B,
                        int H, int W, int C, int heads, int shift, int pano, float scale, cudaStream_t st);

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_window_attn_bwd(const void* qkv, const void* dout, const float* alpha, const float* beta,
                                           const float* qkv_bias, const float* uv, const float* mask, void* dqkv,
                                           float* dalpha, float* dbeta, float* dqkv_bias, int B, int H, int W, int C,
                                           int heads, int window, int shift, int pano_mode, float scale, int dtype,
                                           void* stream) {
  PSW_REQUIRE(qkv && dout && alpha && beta && dqkv, PSW_ERR_BAD_ARG, "psw_window_attn_bwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && heads > 0 && window > 0 && C % heads == 0, PSW_ERR_BAD_ARG,
              "psw_window_attn_bwd: bad dims B=%d H=%d W=%d C=%d heads=%d window=%d", B, H, W, C, heads, window);
  PSW_REQUIRE(shift >= 0 && shift < window, PSW_ERR_BAD_ARG, "psw_window_attn_bwd: shift_size must be in [0, window)");
  PSW_REQUIRE(!pano_mode || uv, PSW_ERR_BAD_ARG, "psw_window_attn_bwd: pano mode needs the uv table");
  PSW_REQUIRE((qkv_bias == nullptr) == (dqkv_bias == nullptr), PSW_ERR_BAD_ARG,
              "psw_window_attn_bwd: qkv_bias and dqkv_bias go together (padding cells route their gradient to the bias)");
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_bwd: too many tokens");
  cudaStream_t st = (cudaStream_t)stream;
  // the table / bias gradients are accumulated with atomics: start from zero
  const size_t tab = sizeof(float) * (size_t)(2 * window - 1) * (2 * window - 1) * heads;
  if (dalpha) PSW_CUDA(cudaMemsetAsync(dalpha, 0, tab, st));
  if (dbeta) PSW_CUDA(cudaMemsetAsync(dbeta, 0, tab, st));
  if (dqkv_bias) PSW_CUDA(cudaMemsetAsync(dqkv_bias, 0, sizeof(float) * 3 * (size_t)C, st));
  if (dtype == PSW_F32)
    return window_attn_bwd<float>((const float*)qkv, (const float*)dout, alpha, beta, qkv_bias, uv, mask, (float*)dqkv, dalpha,
                                  dbeta, dqkv_bias, B, H, W, C, heads, window, shift, pano_mode, scale, st);
  PSW_REQUIRE(dtype == PSW_BF16, PSW_ERR_BAD_ARG, "psw_window_attn_bwd: unknown dtype %d", dtype);
  if (window == 7 && C / heads == 32 && aligned16(qkv) && aligned16(dout) && aligned16(dqkv))   // tensor-core kernel (psw_attn_bwd_mma.cu)
    return window_attn_bwd_mma((const bf16*)qkv, (const bf16*)dout, alpha, beta, qkv_bias, uv, mask, (bf16*)dqkv, dalpha, dbeta,
                               dqkv_bias, B, H, W, C, heads, shift, pano_mode, scale, st);
  return window_attn_bwd<bf16>((const bf16*)qkv, (const bf16*)dout, alpha, beta, qkv_bias, uv, mask, (bf16*)dqkv, dalpha, dbeta,
                               dqkv_bias, B, H, W, C, heads, window, shift, pano_mode, scale, st);
}
