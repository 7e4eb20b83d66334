// Backward of the fused pano window attention for bf16 activations, window 7 / head_dim 32 (every shipped PanoSwin
// config), on warp-level tensor cores (mma.sync m16n8k16, fp32 accumulation) -- SURVEY.md §8 f-3.  Same mathematics and
// the same outputs as the CUDA-core kernel of psw_attn_bwd.cu (which stays the fp32 gradient-parity path and the route
// for other window sizes / head dims); reference: autograd through simple_panoswin_transformer.py:274-311, :376-409.
//
// One CTA of four warps owns one (window position, head) and walks over the images of the batch, so everything that
// depends on the geometry and the block's tables only is set up once per CTA: the token map, the additive logit terms
// hav(i, j) * alpha[idx] + beta[idx] (+ planar shift mask), the great-circle distances needed for d alpha -- and the
// gradient of the bias tables is accumulated in registers over the images and leaves with one shared-memory reduction
// and one global atomic per table entry per CTA.  Per image (49 tokens padded to 64, warp w owns query rows 16w..16w+15):
//   S = Q K^T, dP = dO V^T              A / B fragments with ldmatrix from 80-byte-pitch rows (conflict-free)
//   P = softmax(S scale + bias)         on the accumulator fragments (row max / sum: two quad shuffles), exp2 as forward
//   dS = P o (dP - rowsum(P o dP))      rowsum(P o dP) = rowsum(dO o O): O is never formed
//   dQ = scale dS K                     A = dS straight from the registers (accumulator -> A-fragment re-packing)
//   dK = scale dS^T Q, dV = P^T dO      A = transposed ldmatrix of the bf16 P / dS tiles in shared memory
// dq / dk / dv go to the token's UN-shifted position with 16-byte stores; padding cells (zero tokens whose q / k / v
// equal the qkv bias) send their gradient to d qkv_bias.
#include "psw_common.cuh"

namespace psw {

constexpr int BW_N = 49;           // tokens per window
constexpr int BW_RP = 40;          // q / k / v / dO row pitch in bf16 (32 + 8: the eight rows of an ldmatrix hit different banks)
constexpr int BW_PP = 72;          // P / dS row pitch in bf16 (64 + 8)
constexpr int BW_TP = 50;          // pitch of the fp32 bias / distance tables
constexpr int BW_TAB = 169;        // (2 * 7 - 1)^2 relative positions
constexpr int BW_THREADS = 128;
constexpr float BW_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float bw_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

__global__ void __launch_bounds__(BW_THREADS, 3)
window_attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, const float* __restrict__ alpha,
                           const float* __restrict__ beta, const float* __restrict__ qkv_bias, const float* __restrict__ uv,
                           const float* __restrict__ mask, bf16* __restrict__ dqkv, float* __restrict__ dalpha,
                           float* __restrict__ dbeta, float* __restrict__ dqkv_bias, WinGeom g, int B, int C, int heads,
                           float scale) {
  constexpr int N = BW_N, WS = 7, TW = 13;
  extern __shared__ __align__(16) uint8_t bw_smem[];
  bf16* sq = reinterpret_cast<bf16*>(bw_smem);             // [64][BW_RP] each: q, k, v, dO (later dq, dk, dv staging)
  bf16* sk = sq + 64 * BW_RP;
  bf16* sv = sk + 64 * BW_RP;
  bf16* sdo = sv + 64 * BW_RP;
  bf16* sP = sdo + 64 * BW_RP;                             // [64][BW_PP]
  bf16* sdS = sP + 64 * BW_PP;
  float* sbias = reinterpret_cast<float*>(sdS + 64 * BW_PP);   // [49][BW_TP] additive logit terms x log2 e
  float* shav = sbias + N * BW_TP;                         // [49][BW_TP] great-circle distances
  float* sta = shav + N * BW_TP;                           // [169] d alpha of this (window, head)
  float* stb = sta + BW_TAB;                               // [169] d beta
  float* su = stb + BW_TAB;                                // [49] u, [49] v of the window's tokens
  float* svv = su + N;
  int* ssrc = reinterpret_cast<int*>(svv + N);             // [64] source token or -1
  uint4* spad = reinterpret_cast<uint4*>(ssrc + 64);       // [3][4]: bf16 qkv bias of this head (padding cells)
  float* sqb = reinterpret_cast<float*>(spad + 12);        // [3][32] d qkv_bias of this head through the padding cells

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int e = blockIdx.x % heads;
  const int wi = blockIdx.x / heads;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  const int64_t HW = (int64_t)g.H * g.W;
  const int C3 = 3 * C;

  // ---------------------------------------------------------------- once per CTA
  for (int i = tid; i < 4 * 64 * BW_RP / 8; i += BW_THREADS) reinterpret_cast<uint4*>(sq)[i] = make_uint4(0, 0, 0, 0);
  if (tid < 96) sqb[tid] = 0.f;
  float* sal = reinterpret_cast<float*>(sdS);              // [169] alpha, [169] beta of this head (scratch: dS is not in use yet)
  float* sbe = sal + BW_TAB;
  for (int i = tid; i < BW_TAB; i += BW_THREADS) { sal[i] = alpha[i * heads + e]; sbe[i] = beta[i * heads + e]; }
  if (tid < 64) {
    int s = -1;
    float uu = 0.f, vv = 0.f;
    if (tid < N) {
      const int r = tid / WS, c = tid - r * WS;
      s = source_token(g, wr * WS + r, wc * WS + c);
      if (g.pano && s >= 0) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
      su[tid] = uu;
      svv[tid] = vv;
    }
    ssrc[tid] = s;
  } else if (tid < 76) {
    const int j = tid - 64;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (qkv_bias) {
      const float* bsrc = qkv_bias + (j >> 2) * C + e * 32 + (j & 3) * 8;
      r = make_uint4(pack_bf16x2(bsrc[0], bsrc[1]), pack_bf16x2(bsrc[2], bsrc[3]), pack_bf16x2(bsrc[4], bsrc[5]),
                     pack_bf16x2(bsrc[6], bsrc[7]));
    }
    spad[j] = r;
  }
  __syncthreads();
  // great-circle distances: symmetric, so every pair is evaluated once (fp32 evaluation order of the forward kernels,
  // great_circle.py:82-86; cos v per token in su's neighbour array)
  if (g.pano) {
    float* scos = reinterpret_cast<float*>(sP);            // [49] scratch: the P tile is not in use yet
    if (tid < N) scos[tid] = cosf(svv[tid]);
    __syncthreads();
    for (int p = tid; p < 25 * (N + 1); p += BW_THREADS) {   // upper triangle, rows r and N - 1 - r folded into one of N + 1
      const int r = p / (N + 1), c = p - r * (N + 1);
      int i, j;
      if (c < N - r) { i = r; j = r + c; } else { i = N - 1 - r; j = i + (c - (N - r)); }
      const float sdv = sinf(0.5f * fabsf(svv[j] - svv[i]));
      const float sdu = sinf(0.5f * (su[j] - su[i]));
      const float a = sdv * sdv + (scos[j] * scos[i]) * (sdu * sdu);
      const float d = asinf(sqrtf(fminf(a, 1.0f))) * 2.0f;
      shav[i * BW_TP + j] = d;
      shav[j * BW_TP + i] = d;
    }
    __syncthreads();
  }
  for (int p = tid; p < N * N; p += BW_THREADS) {
    const int i = p / N, j = p - i * N;
    const int ri = i / WS, ci = i - ri * WS, rj = j / WS, cj = j - rj * WS;
    const int idx = (ri - rj + WS - 1) * TW + (ci - cj + WS - 1);
    const float d = g.pano ? shav[i * BW_TP + j] : 0.f;
    float b = fmaf(d, sal[idx], sbe[idx]);
    if (mask) b += mask[((int64_t)wi * N + i) * N + j];
    sbias[i * BW_TP + j] = b * BW_LOG2E;
  }
  __syncthreads();

  const int gq = lane >> 2, tq = lane & 3;                 // fragment coordinates: rows gq, gq + 8; columns 2 tq, 2 tq + 1
  const int i0 = 16 * warp + gq, i1 = i0 + 8;              // my two query rows (S, dP, dQ) / key rows (dK, dV)
  const int lm = lane >> 3, lr = lane & 7;                 // ldmatrix: this lane addresses row lr of matrix lm
  const float scale_l2 = scale * BW_LOG2E;
  float tab[8][4];                                         // sum over the images of dS at my fragment positions
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) tab[nt][c] = 0.f;

  // images of this CTA: the batch is split over gridDim.y when (windows x heads) alone does not fill the GPU
  const int b_per = (B + (int)gridDim.y - 1) / (int)gridDim.y;
  const int b_begin = (int)blockIdx.y * b_per;
  const int b_end = b_begin + b_per < B ? b_begin + b_per : B;
  for (int b = b_begin; b < b_end; ++b) {
    // ---- q / k / v / dO rows of this window, head and image -> shared memory (16-byte chunks; padding cells: bias / 0).
    //      All of a thread's loads are issued before the first store, so their latencies overlap.
    {
      constexpr int ITEMS = (4 * N * 4 + BW_THREADS - 1) / BW_THREADS;     // 7
      uint4 v[ITEMS];
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = tid + it * BW_THREADS;
        v[it] = make_uint4(0, 0, 0, 0);
        if (idx < 4 * N * 4) {
          const int part = idx / (N * 4);
          const int rem = idx - part * (N * 4);
          const int tok = rem >> 2, ch = rem & 3;
          const int s = ssrc[tok];
          if (s >= 0) {
            const bf16* src = part < 3 ? qkv + ((int64_t)b * HW + s) * C3 + part * C + e * 32 + ch * 8
                                       : dout + ((int64_t)b * HW + s) * C + e * 32 + ch * 8;
            v[it] = __ldg(reinterpret_cast<const uint4*>(src));
          } else if (part < 3) {
            v[it] = spad[part * 4 + ch];
          }
        }
      }
#pragma unroll
      for (int it = 0; it < ITEMS; ++it) {
        const int idx = tid + it * BW_THREADS;
        if (idx < 4 * N * 4) {
          const int part = idx / (N * 4);
          const int rem = idx - part * (N * 4);
          *reinterpret_cast<uint4*>(sq + (part * 64 + (rem >> 2)) * BW_RP + (rem & 3) * 8) = v[it];
        }
      }
    }
    __syncthreads();

    // ---- S = Q K^T and dP = dO V^T for my 16 rows
    float sacc[8][4], dacc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int c = 0; c < 4; ++c) { sacc[nt][c] = 0.f; dacc[nt][c] = 0.f; }
    {
      uint32_t aq[2][4], ad[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int off = (16 * warp + lr + (lm & 1) * 8) * BW_RP + 16 * ks + (lm >> 1) * 8;
        ldsm_x4(aq[ks], sq + off);
        ldsm_x4(ad[ks], sdo + off);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bk[4], bv[4];
        const int off = (8 * nt + lr) * BW_RP + lm * 8;    // matrices: k 0..7, 8..15, 16..23, 24..31 of key rows 8 nt ..
        ldsm_x4(bk, sk + off);
        ldsm_x4(bv, sv + off);
        mma_bf16(sacc[nt], aq[0], bk[0], bk[1]);
        mma_bf16(sacc[nt], aq[1], bk[2], bk[3]);
        mma_bf16(dacc[nt], ad[0], bv[0], bv[1]);
        mma_bf16(dacc[nt], ad[1], bv[2], bv[3]);
      }
    }
    // ---- P = softmax(S scale + bias) on the fragments; rows / columns beyond 49 do not exist
    const bool v0 = i0 < N, v1 = i1 < N;
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = 8 * nt + 2 * tq + c;
        const bool vj = j < N;
        const float t0 = (v0 && vj) ? fmaf(sacc[nt][c], scale_l2, sbias[i0 * BW_TP + j]) : -INFINITY;
        const float t1 = (v1 && vj) ? fmaf(sacc[nt][2 + c], scale_l2, sbias[i1 * BW_TP + j]) : -INFINITY;
        sacc[nt][c] = t0;
        sacc[nt][2 + c] = t1;
        m0 = fmaxf(m0, t0);
        m1 = fmaxf(m1, t1);
      }
    }
    m0 = quad_max(m0);
    m1 = quad_max(m1);
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float p0 = v0 ? bw_exp2(sacc[nt][c] - m0) : 0.f;          // exp2(-inf) = 0 for the missing columns
        const float p1 = v1 ? bw_exp2(sacc[nt][2 + c] - m1) : 0.f;
        sacc[nt][c] = p0;
        sacc[nt][2 + c] = p1;
        l0 += p0;
        l1 += p1;
      }
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float r0 = v0 ? 1.0f / l0 : 0.f, r1 = v1 ? 1.0f / l1 : 0.f;
    float e0 = 0.f, e1 = 0.f;                              // rowsum(P o dP)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        sacc[nt][c] *= r0;
        sacc[nt][2 + c] *= r1;
        e0 = fmaf(sacc[nt][c], dacc[nt][c], e0);
        e1 = fmaf(sacc[nt][2 + c], dacc[nt][2 + c], e1);
      }
    }
    e0 = quad_sum(e0);
    e1 = quad_sum(e1);
    // ---- dS = P o (dP - rowsum); bf16 copies of P and dS for the transposed products; table-gradient accumulation
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      dacc[nt][0] = sacc[nt][0] * (dacc[nt][0] - e0);
      dacc[nt][1] = sacc[nt][1] * (dacc[nt][1] - e0);
      dacc[nt][2] = sacc[nt][2] * (dacc[nt][2] - e1);
      dacc[nt][3] = sacc[nt][3] * (dacc[nt][3] - e1);
#pragma unroll
      for (int c = 0; c < 4; ++c) tab[nt][c] += dacc[nt][c];
      const int j = 8 * nt + 2 * tq;
      *reinterpret_cast<uint32_t*>(sP + i0 * BW_PP + j) = pack_bf16x2(sacc[nt][0], sacc[nt][1]);
      *reinterpret_cast<uint32_t*>(sP + i1 * BW_PP + j) = pack_bf16x2(sacc[nt][2], sacc[nt][3]);
      *reinterpret_cast<uint32_t*>(sdS + i0 * BW_PP + j) = pack_bf16x2(dacc[nt][0], dacc[nt][1]);
      *reinterpret_cast<uint32_t*>(sdS + i1 * BW_PP + j) = pack_bf16x2(dacc[nt][2], dacc[nt][3]);
    }
    // ---- dQ = scale dS K for my 16 query rows: A = dS from the registers, B = K as [key][dim] (transposed ldmatrix)
    float qacc[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int c = 0; c < 4; ++c) qacc[dt][c] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(dacc[2 * kk][0], dacc[2 * kk][1]);
      a[1] = pack_bf16x2(dacc[2 * kk][2], dacc[2 * kk][3]);
      a[2] = pack_bf16x2(dacc[2 * kk + 1][0], dacc[2 * kk + 1][1]);
      a[3] = pack_bf16x2(dacc[2 * kk + 1][2], dacc[2 * kk + 1][3]);
#pragma unroll
      for (int d2 = 0; d2 < 2; ++d2) {
        uint32_t bb[4];                                    // matrices: (keys 0..7, dims 0..7), (keys 8..15, dims 0..7), (0..7, 8..15), (8..15, 8..15)
        ldsm_x4_t(bb, sk + (16 * kk + lr + (lm & 1) * 8) * BW_RP + 16 * d2 + (lm >> 1) * 8);
        mma_bf16(qacc[2 * d2], a, bb[0], bb[1]);
        mma_bf16(qacc[2 * d2 + 1], a, bb[2], bb[3]);
      }
    }
    __syncthreads();                                       // P and dS tiles complete
    // ---- dK = scale dS^T Q and dV = P^T dO for my 16 key rows: contraction over all 64 query rows
    float kacc[4][4], vacc[4][4];
#pragma unroll
    for (int dt = 0; dt < 4; ++dt)
#pragma unroll
      for (int c = 0; c < 4; ++c) { kacc[dt][c] = 0.f; vacc[dt][c] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ads[4], ap[4];                              // A^T: matrices (m 0..7, k 0..7), (m 8..15, k 0..7), (m 0..7, k 8..15), (m 8..15, k 8..15)
      const int offa = (16 * kk + lr + (lm >> 1) * 8) * BW_PP + 16 * warp + (lm & 1) * 8;
      ldsm_x4_t(ads, sdS + offa);
      ldsm_x4_t(ap, sP + offa);
#pragma unroll
      for (int d2 = 0; d2 < 2; ++d2) {
        uint32_t bq[4], bo[4];
        const int offb = (16 * kk + lr + (lm & 1) * 8) * BW_RP + 16 * d2 + (lm >> 1) * 8;
        ldsm_x4_t(bq, sq + offb);
        ldsm_x4_t(bo, sdo + offb);
        mma_bf16(kacc[2 * d2], ads, bq[0], bq[1]);
        mma_bf16(kacc[2 * d2 + 1], ads, bq[2], bq[3]);
        mma_bf16(vacc[2 * d2], ap, bo[0], bo[1]);
        mma_bf16(vacc[2 * d2 + 1], ap, bo[2], bo[3]);
      }
    }
    __syncthreads();                                       // every read of q / k / v / dO is done: reuse them as staging
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      const int d = 8 * dt + 2 * tq;
      *reinterpret_cast<uint32_t*>(sq + i0 * BW_RP + d) = pack_bf16x2(qacc[dt][0] * scale, qacc[dt][1] * scale);
      *reinterpret_cast<uint32_t*>(sq + i1 * BW_RP + d) = pack_bf16x2(qacc[dt][2] * scale, qacc[dt][3] * scale);
      *reinterpret_cast<uint32_t*>(sk + i0 * BW_RP + d) = pack_bf16x2(kacc[dt][0] * scale, kacc[dt][1] * scale);
      *reinterpret_cast<uint32_t*>(sk + i1 * BW_RP + d) = pack_bf16x2(kacc[dt][2] * scale, kacc[dt][3] * scale);
      *reinterpret_cast<uint32_t*>(sv + i0 * BW_RP + d) = pack_bf16x2(vacc[dt][0], vacc[dt][1]);
      *reinterpret_cast<uint32_t*>(sv + i1 * BW_RP + d) = pack_bf16x2(vacc[dt][2], vacc[dt][3]);
    }
    __syncthreads();
    // ---- dq / dk / dv rows to the tokens' un-shifted positions (padding cells -> d qkv_bias)
    for (int idx = tid; idx < 3 * N * 4; idx += BW_THREADS) {
      const int part = idx / (N * 4);
      const int rem = idx - part * (N * 4);
      const int tok = rem >> 2, ch = rem & 3;
      const int s = ssrc[tok];
      const uint4 v = *reinterpret_cast<const uint4*>(sq + (part * 64 + tok) * BW_RP + ch * 8);
      if (s >= 0) {
        *reinterpret_cast<uint4*>(dqkv + ((int64_t)b * HW + s) * C3 + part * C + e * 32 + ch * 8) = v;
      } else if (dqkv_bias) {
        const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w4[k]);
          atomicAdd(sqb + part * 32 + ch * 8 + 2 * k, __low2float(h));
          atomicAdd(sqb + part * 32 + ch * 8 + 2 * k + 1, __high2float(h));
        }
      }
    }
    __syncthreads();                                       // before the next image overwrites the tiles
  }

  // ---------------------------------------------------------------- table gradients of this (window, head)
  // d beta[idx] = sum of dS over the (i, j) pairs at relative position idx, d alpha[idx] the same weighted with the
  // great-circle distance: the accumulated dS goes to shared memory (the bias table is dead by now) and one thread per
  // table entry walks over its <= 49 pairs -- no shared-memory atomics.
  float* sds = sbias;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = (c & 2) ? i1 : i0;
      const int j = 8 * nt + 2 * tq + (c & 1);
      if (i < N && j < N) sds[i * BW_TP + j] = tab[nt][c];
    }
  }
  __syncthreads();
  for (int t = tid; t < BW_TAB; t += BW_THREADS) {
    const int dr = t / TW - (WS - 1), dc = t - (t / TW) * TW - (WS - 1);     // idx = (ri - rj + 6) * 13 + (ci - cj + 6)
    float sb = 0.f, sa = 0.f;
    for (int ri = (dr > 0 ? dr : 0); ri < (dr < 0 ? WS + dr : WS); ++ri) {
      for (int ci = (dc > 0 ? dc : 0); ci < (dc < 0 ? WS + dc : WS); ++ci) {
        const int i = ri * WS + ci, j = (ri - dr) * WS + (ci - dc);
        const float d = sds[i * BW_TP + j];
        sb += d;
        sa = fmaf(shav[i * BW_TP + j], d, sa);
      }
    }
    if (g.pano && dalpha) atomicAdd(dalpha + t * heads + e, sa);
    if (dbeta) atomicAdd(dbeta + t * heads + e, sb);
  }
  if (dqkv_bias && tid < 96 && sqb[tid] != 0.f) atomicAdd(dqkv_bias + (tid >> 5) * C + e * 32 + (tid & 31), sqb[tid]);
}

// bf16, window 7, head_dim 32; the table / bias gradients must be zero on entry (the caller clears them)
int window_attn_bwd_mma(const bf16* qkv, const bf16* dout, const float* alpha, const float* beta, const float* qkv_bias,
                        const float* uv, const float* mask, bf16* dqkv, float* dalpha, float* dbeta, float* dqkv_bias, int B,
                        int H, int W, int C, int heads, int shift, int pano, float scale, cudaStream_t st) {
  WinGeom g = make_geom(H, W, 7, shift, pano);
  const size_t smem = (size_t)4 * 64 * BW_RP * 2 + (size_t)2 * 64 * BW_PP * 2 + (size_t)2 * BW_N * BW_TP * 4 + 2 * BW_TAB * 4 +
                      2 * BW_N * 4 + 64 * 4 + 12 * 16 + 96 * 4;
  const int64_t blocks = (int64_t)g.nWh * g.nWw * heads;
  PSW_REQUIRE(blocks < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_bwd: too many windows");
  // a CTA walks over images serially (its per-window setup is amortised); split the batch until ~3 waves of CTAs exist
  int splits = (int)((3ll * 3 * num_sms() + blocks - 1) / blocks);
  if (splits > B) splits = B;
  if (splits < 1) splits = 1;
  splits = (B + ((B + splits - 1) / splits) - 1) / ((B + splits - 1) / splits);
  PSW_CUDA(cudaFuncSetAttribute(window_attn_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  window_attn_bwd_mma_kernel<<<dim3((unsigned)blocks, (unsigned)splits), BW_THREADS, smem, st>>>(qkv, dout, alpha, beta, qkv_bias, uv, mask, dqkv, dalpha,
                                                                          dbeta, dqkv_bias, g, B, C, heads, scale);
  return launch_status("window_attn_bwd_mma_kernel");
}

}  // namespace psw
