// K3 (fp32 parity path, any window / head_dim): fused pano-shift + window partition + attention +
// reverse on CUDA cores.  One CTA per (window, head); the window's q/k/v and the NxN logits stay in
// shared memory, so HBM sees each qkv element once and each output element once — the shifted,
// padded and partitioned copies the reference materialises (simple_panoswin_transformer.py:508-519)
// never exist.  Math order follows BasicWindowAttention.forward (:290-308) and haversine22
// (lzx/models/great_circle.py:82-86) in fp32.  The throughput path is psw_attn_tc.cu.
#include "psw_common.cuh"

namespace psw {

template <typename T>
__global__ void __launch_bounds__(128)
window_attn_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, const float* __restrict__ alpha,
                        const float* __restrict__ beta, const float* __restrict__ qkv_bias,
                        const float* __restrict__ uv, const float* __restrict__ mask, WinGeom g, int C, int heads,
                        float scale) {
  extern __shared__ float sm[];
  const int ws = g.ws;
  const int N = ws * ws;
  const int hd = C / heads;
  const int SP = N + 1;                                   // padded row pitch of the logits
  const int hp = hd + 1;                                  // padded row pitch of q / k / v (bank-conflict-free token walks)
  float* sq = sm;                                         // [N][hp]  (already * scale)
  float* sk = sq + N * hp;                                // [N][hp]
  float* sv = sk + N * hp;                                // [N][hp]
  float* sS = sv + N * hp;                                // [N][N+1]
  float* su = sS + N * SP;                                // [N]
  float* sw = su + N;                                     // [N] (v coordinate)
  int* ssrc = reinterpret_cast<int*>(sw + N);             // [N]

  const int tid = threadIdx.x;
  const int e = blockIdx.x % heads;
  const int win = blockIdx.x / heads;
  const int wpi = g.nWh * g.nWw;
  const int b = win / wpi;
  const int wi = win - b * wpi;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  const int64_t HW = (int64_t)g.H * g.W;

  for (int t = tid; t < N; t += blockDim.x) {
    int r = t / ws, c = t - r * ws;
    int s = source_token(g, wr * ws + r, wc * ws + c);
    ssrc[t] = s;
    float uu = 0.f, vv = 0.f;
    if (g.pano && s >= 0) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
    su[t] = uu;
    sw[t] = vv;
  }
  __syncthreads();

  for (int idx = tid; idx < 3 * N * hd; idx += blockDim.x) {
    int part = idx / (N * hd);
    int rem = idx - part * N * hd;
    int t = rem / hd, d = rem - t * hd;
    int ch = part * C + e * hd + d;
    int s = ssrc[t];
    float val;
    if (s >= 0) val = to_f32(qkv[((int64_t)b * HW + s) * (3 * C) + ch]);
    else        val = qkv_bias ? qkv_bias[ch] : 0.f;       // zero (padding) token: qkv = bias
    if (part == 0) val *= scale;
    sm[(part * N + t) * hp + d] = val;                      // sq, sk, sv are contiguous
  }
  __syncthreads();

  const int tw = 2 * ws - 1;
  for (int p = tid; p < N * N; p += blockDim.x) {
    int i = p / N, j = p - i * N;
    float dot = 0.f;
    for (int d = 0; d < hd; ++d) dot = fmaf(sq[i * hp + d], sk[j * hp + d], dot);
    int ri = i / ws, ci = i - ri * ws, rj = j / ws, cj = j - rj * ws;
    int idx = (ri - rj + ws - 1) * tw + (ci - cj + ws - 1);
    float bia = beta[idx * heads + e];
    if (g.pano) {
      float sdv = sinf(0.5f * fabsf(sw[j] - sw[i]));
      float sdu = sinf(0.5f * (su[j] - su[i]));
      float a = sdv * sdv + (cosf(sw[j]) * cosf(sw[i])) * (sdu * sdu);
      float hav = asinf(sqrtf(a)) * 2.0f;
      bia = hav * alpha[idx * heads + e] + bia;
    }
    float s = dot + bia;
    if (mask) s += mask[((int64_t)wi * N + i) * N + j];
    sS[i * SP + j] = s;
  }
  __syncthreads();

  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp; i < N; i += nwarps) {
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, sS[i * SP + j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < N; j += 32) {
      float ex = expf(sS[i * SP + j] - m);
      sS[i * SP + j] = ex;
      sum += ex;
    }
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    for (int j = lane; j < N; j += 32) sS[i * SP + j] *= inv;
  }
  __syncthreads();

  for (int idx = tid; idx < N * hd; idx += blockDim.x) {
    int t = idx / hd, d = idx - t * hd;
    int s = ssrc[t];
    if (s < 0) continue;                                   // padded cells are cropped (:516)
    float acc = 0.f;
    for (int j = 0; j < N; ++j) acc = fmaf(sS[t * SP + j], sv[j * hp + d], acc);
    out[((int64_t)b * HW + s) * C + e * hd + d] = from_f32<T>(acc);
  }
}

template <typename T>
int window_attn_simt(const T* qkv, T* out, const float* alpha, const float* beta, const float* qkv_bias,
                     const float* uv, const float* mask, int B, int H, int W, int C, int heads, int window, int shift,
                     int pano, float scale, cudaStream_t st) {
  WinGeom g = make_geom(H, W, window, shift, pano);
  int N = window * window, hd = C / heads;
  size_t smem = ((size_t)3 * N * (hd + 1) + (size_t)N * (N + 1) + 3 * (size_t)N) * sizeof(float);
  PSW_REQUIRE(smem <= 220 * 1024, PSW_ERR_UNSUPPORTED, "window attention (simt): window %d x head_dim %d needs %zu B smem",
              window, hd, smem);
  int64_t blocks = (int64_t)B * g.nWh * g.nWw * heads;
  PSW_REQUIRE(blocks < (1ll << 31), PSW_ERR_UNSUPPORTED, "window attention (simt): too many windows");
  PSW_CUDA(cudaFuncSetAttribute(window_attn_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  window_attn_simt_kernel<T><<<(unsigned)blocks, 128, smem, st>>>(qkv, out, alpha, beta, qkv_bias, uv, mask, g, C, heads, scale);
  return launch_status("window_attn_simt_kernel");
}

template int window_attn_simt<float>(const float*, float*, const float*, const float*, const float*, const float*,
                                     const float*, int, int, int, int, int, int, int, int, float, cudaStream_t);
template int window_attn_simt<bf16>(const bf16*, bf16*, const float*, const float*, const float*, const float*,
                                    const float*, int, int, int, int, int, int, int, int, float, cudaStream_t);

}  // namespace psw
