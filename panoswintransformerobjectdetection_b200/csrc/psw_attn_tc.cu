// K3 (throughput path): fused pano-shift + window-partition + multi-head attention + window-reverse +
// un-shift for bf16 activations on sm_100a.  Replaces WindowTransition / pad_x / window_partition /
// BasicWindowAttention core / window_reverse of the reference
// (simple_panoswin_transformer.py:376-409, :486-491, :64-92, :290-308) in one pass over HBM.
//
// Work unit = one (window, head): 49 tokens x head_dim 32.  Two consecutive units form a "pair" that
// shares one 128-row tensor-core tile:
//   S[128x128] = [Q_u0;Q_u1] . [K_u0;K_u1]^T      tcgen05.mma M=128 N=128 K=32, only the two diagonal
//                                                 64x64 blocks are used (tensor pipe is far from the bound)
//   O[128x64]  = P[128x64 keys] . [V_u0 | V_u1]   tcgen05.mma M=128 N=64 K=64, P read from TMEM (or smem),
//                                                 V consumed MN-major straight from its [key][dim] rows;
//                                                 rows of unit u use output columns [32u, 32u+32)
// Thread r of the CTA owns row r of the tile (TMEM lane r): it reads its 49 logits with tcgen05.ld,
// adds the great-circle bias d(i,j)*alpha[idx]+beta[idx] (d from shared memory, computed once per window
// from the fp32 uv table), does the softmax in registers (exp2, fp32), writes un-normalised bf16 P back,
// and finally scales its O row by 1/sum and stores it with 128-bit stores to the token's UN-shifted
// position.  q/k/v rows are gathered by cp.async (16 B) directly from the un-shifted [B,H,W,3C] qkv tensor
// into the 64B-swizzled UMMA layout: the pano shift with longitude wrap-around, the odd-W zero column,
// the window padding (padding tokens = qkv bias) and the partition are pure address arithmetic
// (psw::source_token).  The next pair is prefetched while the current one is computed; 2-3 CTAs per SM
// overlap each other's MMA / softmax / store phases.
//
// Algorithmic HBM bytes per unit: 49 * 32 * 2 B * 4 (q, k, v read + o written) = 12,544 B.
#include "psw_common.cuh"

namespace psw {

constexpr int AT_THREADS = 128;
constexpr int AT_PART_BYTES = 128 * 64;            // 128 rows x 64 B (32 bf16), SWIZZLE_64B
constexpr int AT_BUF_BYTES = 3 * AT_PART_BYTES;    // q, k, v
constexpr int AT_HAV_PITCH = 52;                   // floats per distance-matrix row (16 B aligned rows)
constexpr int AT_TMEM_COLS = 128;
constexpr int AT_P_COL = 0;                        // P (bf16x2 packed): TMEM columns [0, 32)
constexpr int AT_O_COL = 32;                       // O (fp32): TMEM columns [32, 96)
constexpr float LOG2E = 1.4426950408889634f;
#ifndef PSW_ATTN_TC_DEFAULT_VARIANT
#define PSW_ATTN_TC_DEFAULT_VARIANT 0              // 0: P stays in TMEM (A-from-TMEM MMA), 1: P through smem
#endif

__device__ __forceinline__ void tmem_ld_x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  const bf16* qkv;
  bf16* out;
  const float* alpha;
  const float* beta;
  const float* qkv_bias;
  const float* uv;
  const float* mask;
  WinGeom g;
  int B, C, heads;
  float scale;
  int64_t total_units;
};

template <int WS, bool P_IN_TMEM>
__global__ void __launch_bounds__(AT_THREADS)
window_attn_tc_kernel(const AttnParams p) {
  constexpr int N = WS * WS;                     // tokens per window (<= 64)
  constexpr int TW = 2 * WS - 1;                 // relative-position table width
  constexpr int TAB = TW * TW;
  static_assert(N <= 64, "window too large for the 64-row unit tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bufs = smem;                                                  // [2][3][128 x 64 B]
  uint8_t* psm = bufs + 2 * AT_BUF_BYTES;                                // [128 x 128 B] P (smem variant only)
  float* hav = reinterpret_cast<float*>(psm + (P_IN_TMEM ? 0 : 128 * 128));  // [2][N][52]
  float2* tab = reinterpret_cast<float2*>(hav + 2 * N * AT_HAV_PITCH);   // [2][TAB] (alpha, beta) * log2(e)
  int* tok_src = reinterpret_cast<int*>(tab + 2 * TAB);                  // [4][64]
  float* tok_u = reinterpret_cast<float*>(tok_src + 4 * 64);             // [4][64]
  float* tok_v = tok_u + 4 * 64;                                         // [4][64]
  float* tok_cv = tok_v + 4 * 64;                                        // [4][64] cos(v)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tok_cv + 4 * 64);         // [2]: S ready, O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  bf16* bias_bf = reinterpret_cast<bf16*>(tmem_slot + 4);                // [3C] (16 B aligned)

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int C = p.C, heads = p.heads;
  const WinGeom g = p.g;
  const int wpi = g.nWh * g.nWw;
  const int64_t HW = (int64_t)g.H * g.W;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < (2 * AT_BUF_BYTES + (P_IN_TMEM ? 0 : 128 * 128)) / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(bufs)[i] = make_uint4(0, 0, 0, 0);          // padding rows must stay finite
  for (int i = tid; i < 2 * N * AT_HAV_PITCH; i += AT_THREADS) hav[i] = 0.f;   // planar mode: d == 0
  for (int i = tid; i < 3 * C; i += AT_THREADS) bias_bf[i] = __float2bfloat16_rn(p.qkv_bias ? p.qkv_bias[i] : 0.f);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int64_t n_pairs = (p.total_units + 1) >> 1;
  const int64_t pair_begin = n_pairs * blockIdx.x / gridDim.x;
  const int64_t pair_end = n_pairs * (blockIdx.x + 1) / gridDim.x;

  int tok_ready_win = -1;          // windows <= this have their token maps in tok_*[win & 3]
  int hav_win0 = -1, hav_win1 = -1;  // window whose distance matrix sits in hav slot 0 / 1

  // token map + coordinates of window `win` (all threads call; threads < N work)
  auto prep_tokens = [&](int win) {
    if (tid < N) {
      const int wi = win % wpi;
      const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
      const int r = tid / WS, c = tid - r * WS;
      const int s = source_token(g, wr * WS + r, wc * WS + c);
      float uu = 0.f, vv = 0.f;
      if (g.pano && s >= 0) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p.uv) + s);
        uu = t.x; vv = t.y;
      }
      const int o = (win & 3) * 64 + tid;
      tok_src[o] = s;
      tok_u[o] = uu;
      tok_v[o] = vv;
      tok_cv[o] = cosf(vv);
    }
  };
  // great-circle distance matrix of window `win` (haversine22, lzx/models/great_circle.py:82-86)
  auto compute_hav = [&](int win) {
    const int ts = (win & 3) * 64;
    float* h = hav + (win & 1) * N * AT_HAV_PITCH;
    for (int q = tid; q < N * N; q += AT_THREADS) {
      const int i = q / N, j = q - i * N;
      const float sdv = __sinf(0.5f * (tok_v[ts + j] - tok_v[ts + i]));
      const float sdu = __sinf(0.5f * (tok_u[ts + j] - tok_u[ts + i]));
      float a = sdv * sdv + tok_cv[ts + j] * tok_cv[ts + i] * (sdu * sdu);
      a = fminf(fmaxf(a, 0.f), 1.f);
      h[i * AT_HAV_PITCH + j] = 2.0f * asinf(sqrtf(a));
    }
  };
  // gather q/k/v rows of pair `pr` into buffer `buf` (cp.async 16 B; padding tokens take the qkv bias)
  auto issue_loads = [&](int64_t pr, int buf) {
    uint8_t* base = bufs + buf * AT_BUF_BYTES;
    for (int id = tid; id < 2 * 3 * N * 4; id += AT_THREADS) {
      const int chunk = id & 3;
      const int item = id >> 2;
      const int unit = item / (3 * N);
      const int rem = item - unit * 3 * N;
      const int part = rem / N;
      const int t = rem - part * N;
      const int64_t u = 2 * pr + unit;
      if (u >= p.total_units) continue;
      const int win = (int)(u / heads);
      const int e = (int)(u - (int64_t)win * heads);
      const int row = unit * 64 + t;
      uint8_t* dst = base + part * AT_PART_BYTES + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
      const int s = tok_src[(win & 3) * 64 + t];
      const int ch = part * C + e * 32 + chunk * 8;
      if (s >= 0) {
        const int b = win / wpi;
        cp_async16(dst, p.qkv + ((int64_t)b * HW + s) * (3 * C) + ch);
      } else {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(bias_bf + ch);
      }
    }
  };
  auto prep_pair_tokens = [&](int64_t pr) {
    const int64_t u0 = 2 * pr;
    int64_t u1 = u0 + 1;
    if (u1 >= p.total_units) u1 = u0;
    const int w0 = (int)(u0 / heads), w1 = (int)(u1 / heads);
    for (int w = (tok_ready_win + 1 > w0 ? tok_ready_win + 1 : w0); w <= w1; ++w) prep_tokens(w);
    if (w1 > tok_ready_win) tok_ready_win = w1;
  };

  if (pair_begin < pair_end) {
    prep_pair_tokens(pair_begin);
    __syncthreads();
    issue_loads(pair_begin, 0);
  }
  cp_async_commit();

  int it = 0;
  for (int64_t pr = pair_begin; pr < pair_end; ++pr, ++it) {
    const int buf = it & 1;
    const uint32_t par = (uint32_t)(it & 1);
    __syncthreads();         // every thread is done with the previous pair's token maps / tables / TMEM rows
    // ---- A. prefetch the next pair (its buffer was released by the PV commit of the previous iteration)
    if (pr + 1 < pair_end) prep_pair_tokens(pr + 1);
    __syncthreads();
    if (pr + 1 < pair_end) issue_loads(pr + 1, buf ^ 1);
    cp_async_commit();

    // ---- B. per-window distance matrix + per-head tables of this pair
    const int64_t u0 = 2 * pr;
    const bool valid1 = (u0 + 1) < p.total_units;
    const int win0 = (int)(u0 / heads);
    const int e0 = (int)(u0 - (int64_t)win0 * heads);
    const int win1 = valid1 ? (int)((u0 + 1) / heads) : win0;
    const int e1 = valid1 ? (int)((u0 + 1) - (int64_t)win1 * heads) : e0;
    if (g.pano) {
      if ((win0 & 1) ? (hav_win1 != win0) : (hav_win0 != win0)) {
        compute_hav(win0);
        if (win0 & 1) hav_win1 = win0; else hav_win0 = win0;
      }
      if (win1 != win0 && ((win1 & 1) ? (hav_win1 != win1) : (hav_win0 != win1))) {
        compute_hav(win1);
        if (win1 & 1) hav_win1 = win1; else hav_win0 = win1;
      }
    }
    for (int i = tid; i < 2 * TAB; i += AT_THREADS) {
      const int unit = i / TAB;
      const int k = i - unit * TAB;
      const int e = unit ? e1 : e0;
      tab[i] = make_float2(__ldg(p.alpha + k * heads + e) * LOG2E, __ldg(p.beta + k * heads + e) * LOG2E);
    }

    // ---- C. this pair's q/k/v have landed
    cp_async_wait<1>();
    fence_async_shared();
    __syncthreads();

    // ---- D. S = Q . K^T (both units at once, block diagonal)
    const uint32_t sq = smem_u32(bufs + buf * AT_BUF_BYTES);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      const uint64_t dq = umma_smem_desc(sq, 16, 512, UMMA_SWIZZLE_64B);
      const uint64_t dk = umma_smem_desc(sq + AT_PART_BYTES, 16, 512, UMMA_SWIZZLE_64B);
      umma_ss(tmem_base, dq, dk, idesc, 0);
      umma_ss(tmem_base, dq + 2, dk + 2, idesc, 1);          // head_dim 16..31: +32 B inside the swizzle row
      umma_commit(&bars[0]);
    }
    mbar_wait(&bars[0], par);
    tc_fence_after();

    // ---- E. bias + softmax on my row
    const int unit = tid >> 6;                               // warp-uniform
    const int i = tid & 63;
    const bool row_valid = (i < N) && (unit == 0 || valid1);
    const int ic = i < N ? i : 0;
    const int my_win = unit ? win1 : win0;
    float sum = 1.f;
    {
      uint32_t sr[N];
      const uint32_t s_addr = tmem_base + lane_base + (uint32_t)(unit * 64);
      {
        uint32_t t32[32];
        tmem_ld_x32(s_addr, t32);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) sr[k] = t32[k];
      }
      if constexpr (N > 32) {
        uint32_t t16[16];
        tmem_ld_x16(s_addr + 32, t16);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 16 && 32 + k < N; ++k) sr[32 + k] = t16[k];
      }
      if constexpr (N > 48) {
#pragma unroll
        for (int k = 48; k < N; ++k) {
          uint32_t t1;
          tmem_ld_x1(s_addr + (uint32_t)k, t1);
          tmem_ld_wait();
          sr[k] = t1;
        }
      }
      const float* hrow = hav + (my_win & 1) * N * AT_HAV_PITCH + ic * AT_HAV_PITCH;
      const int ri = ic / WS, ci = ic - ri * WS;
      const float2* trow = tab + unit * TAB + (ri + WS - 1) * TW + (ci + WS - 1);
      const float* mrow = p.mask ? p.mask + ((int64_t)(my_win % wpi) * N + ic) * N : nullptr;
      const float sc = p.scale * LOG2E;
      float t[N];
      float mx = -INFINITY;
      const float4* h4 = reinterpret_cast<const float4*>(hrow);   // rows are 16 B aligned (pitch 52 floats)
#pragma unroll
      for (int jc = 0; jc < (N + 3) / 4; ++jc) {
        const float4 hq = h4[jc];
        const float hv[4] = {hq.x, hq.y, hq.z, hq.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * jc + q;
          if (j < N) {
            const float2 ab = trow[-((j / WS) * TW + (j % WS))];
            float bia = fmaf(hv[q], ab.x, ab.y);
            if (mrow) bia = fmaf(__ldg(mrow + j), LOG2E, bia);
            t[j] = fmaf(__uint_as_float(sr[j]), sc, bia);
            mx = fmaxf(mx, t[j]);
          }
        }
      }
      sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float p0 = 0.f, p1 = 0.f;
        if (2 * k < N) { p0 = fast_exp2(t[2 * k] - mx); sum += p0; }
        if (2 * k + 1 < N) { p1 = fast_exp2(t[2 * k + 1] - mx); sum += p1; }
        pk[k] = row_valid ? pack_bf16x2(p0, p1) : 0u;
      }
      if constexpr (P_IN_TMEM) {
        tmem_st_x32(tmem_base + lane_base + AT_P_COL, pk);
        tmem_st_wait();
      } else {
        uint8_t* prow = psm + tid * 128;                     // K-major, SWIZZLE_128B
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(prow + ((c ^ (tid & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
        fence_async_shared();
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- F. O = P . [V_u0 | V_u1]
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);     // B (V) is MN-major: [key][dim] rows
      const uint32_t sv = sq + 2 * AT_PART_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {                              // 16 keys per MMA = two 8-key groups of 512 B
        const uint64_t dv = umma_smem_desc(sv + k * 1024, 4096, 512, UMMA_SWIZZLE_64B);
        if constexpr (P_IN_TMEM) {
          umma_ts(tmem_base + AT_O_COL, tmem_base + AT_P_COL + k * 8, dv, idesc, k > 0);
        } else {
          const uint64_t dp = umma_smem_desc(smem_u32(psm) + k * 32, 16, 1024, UMMA_SWIZZLE_128B);
          umma_ss(tmem_base + AT_O_COL, dp, dv, idesc, k > 0);
        }
      }
      umma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], par);
    tc_fence_after();

    // ---- G. normalise and store my output row at the token's un-shifted position
    {
      uint32_t orow[32];
      tmem_ld_x32(tmem_base + lane_base + AT_O_COL + (uint32_t)(unit * 32), orow);
      tmem_ld_wait();
      const int s = tok_src[(my_win & 3) * 64 + ic];
      if (row_valid && s >= 0) {
        const float inv = 1.0f / sum;
        const int b = my_win / wpi;
        const int e = unit ? e1 : e0;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)b * HW + s) * C + e * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(orow[8 * c + 0]) * inv, __uint_as_float(orow[8 * c + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(orow[8 * c + 2]) * inv, __uint_as_float(orow[8 * c + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(orow[8 * c + 4]) * inv, __uint_as_float(orow[8 * c + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(orow[8 * c + 6]) * inv, __uint_as_float(orow[8 * c + 7]) * inv);
          dst[c] = v;
        }
      }
    }
    tc_fence_before();       // the next iteration's barrier orders these TMEM reads before the next S MMA
  }

  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AT_TMEM_COLS) : "memory");
  }
}

static size_t attn_tc_smem_bytes(int ws, int C, bool p_in_tmem) {
  const int N = ws * ws, TAB = (2 * ws - 1) * (2 * ws - 1);
  size_t b = 1024;                                   // alignment slack
  b += 2 * AT_BUF_BYTES;
  b += p_in_tmem ? 0 : 128 * 128;
  b += (size_t)2 * N * AT_HAV_PITCH * 4;
  b += (size_t)2 * TAB * 8;
  b += 4 * 64 * 4 * 4;
  b += 2 * 8 + 16;
  b += (size_t)3 * C * 2;
  return b;
}

template <int WS, bool P_IN_TMEM>
static int launch_attn_tc(const AttnParams& p, int ctas_per_sm, cudaStream_t st) {
  const size_t smem = attn_tc_smem_bytes(WS, p.C, P_IN_TMEM);
  auto kern = window_attn_tc_kernel<WS, P_IN_TMEM>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_pairs = (p.total_units + 1) / 2;
  int64_t grid = (int64_t)num_sms() * ctas_per_sm;
  if (grid > n_pairs) grid = n_pairs;
  kern<<<(unsigned)grid, AT_THREADS, smem, st>>>(p);
  return launch_status("window_attn_tc_kernel");
}

// variant: 0 = P through TMEM (tcgen05.mma A-from-TMEM), 1 = P through shared memory
int window_attn_tc_variant(const bf16* qkv, bf16* out, const float* alpha, const float* beta, const float* qkv_bias,
                           const float* uv, const float* mask, int B, int H, int W, int C, int heads, int window,
                           int shift, int pano, float scale, int variant, cudaStream_t st) {
  AttnParams p;
  p.qkv = qkv; p.out = out; p.alpha = alpha; p.beta = beta; p.qkv_bias = qkv_bias; p.uv = uv; p.mask = mask;
  p.g = make_geom(H, W, window, shift, pano);
  p.B = B; p.C = C; p.heads = heads; p.scale = scale;
  p.total_units = (int64_t)B * p.g.nWh * p.g.nWw * heads;
  PSW_REQUIRE(window == 7, PSW_ERR_UNSUPPORTED,
              "psw_window_attn_fwd(bf16): the tcgen05 kernel is instantiated for window 7 (every shipped PanoSwin config); got %d",
              window);
  if (variant == 0) return launch_attn_tc<7, true>(p, 2, st);
  return launch_attn_tc<7, false>(p, 2, st);
}

int window_attn_tc(const bf16* qkv, bf16* out, const float* alpha, const float* beta, const float* qkv_bias,
                   const float* uv, const float* mask, int B, int H, int W, int C, int heads, int window, int shift,
                   int pano, float scale, cudaStream_t st) {
  return window_attn_tc_variant(qkv, out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads, window, shift, pano, scale,
                                PSW_ATTN_TC_DEFAULT_VARIANT, st);
}

}  // namespace psw

extern "C" PSW_API int psw_window_attn_fwd_tc_variant(const void* qkv, void* out, const float* alpha, const float* beta,
                                              const float* qkv_bias, const float* uv, const float* mask, int B, int H,
                                              int W, int C, int heads, int window, int shift, int pano_mode,
                                              float scale, int variant, void* stream) {
  using namespace psw;
  PSW_REQUIRE(qkv && out && alpha && beta, PSW_ERR_BAD_ARG, "psw_window_attn_fwd_tc_variant: null pointer");
  PSW_REQUIRE(C / heads == 32 && C % heads == 0, PSW_ERR_UNSUPPORTED, "head_dim must be 32");
  return window_attn_tc_variant((const bf16*)qkv, (bf16*)out, alpha, beta, qkv_bias, uv, mask, B, H, W, C, heads, window,
                                shift, pano_mode, scale, variant, (cudaStream_t)stream);
}
