// K3 (throughput path): fused pano-shift + window-partition + multi-head attention + window-reverse +
// un-shift for bf16 activations on sm_100a.  Replaces WindowTransition / pad_x / window_partition /
// BasicWindowAttention core / window_reverse of the reference
// (simple_panoswin_transformer.py:376-409, :486-491, :64-92, :290-308) in one pass over HBM.
//
// A "unit" is one (window, head): 49 tokens x head_dim 32.  The tensor-core tile has 128 rows = two units:
//   S[128x128] = [Q_u0;Q_u1] . [K_u0;K_u1]^T      tcgen05.mma M=128 N=128 K=32; only the two diagonal 64x64
//                                                 blocks are used (the tensor pipe is far from the bound)
//   O[128x64]  = P[128x64 keys] . [V_u0 | V_u1]   tcgen05.mma M=128 N=64 K=64, P read from TMEM (bf16),
//                                                 V consumed MN-major straight from its [key][dim] rows;
//                                                 rows of unit u use output columns [32u, 32u+32)
// Thread r owns tile row r (TMEM lane r): tcgen05.ld of its 49 logits, t = S * (scale * log2 e) + bias (every
// additive term -- great-circle bias, relative-position bias, planar shift mask -- comes precomputed and already
// multiplied by log2 e from psw_window_bias_full), row maximum, exp2, un-normalised bf16 P back to TMEM, finally
// the O row * 1/sum stored with 128-bit stores at the token's UN-shifted position.  q/k/v rows are gathered with
// cp.async (16 B) straight from the un-shifted [B,H,W,3C] qkv tensor into the 64B-swizzled UMMA layout: pano shift
// with longitude wrap-around, the odd-W zero column, window padding (padding tokens = qkv bias) and partition are
// address arithmetic (psw::source_token).  q/k/v are double-buffered per CTA: the whole next step is prefetched
// while the S MMA of the current one runs; 4 CTAs per SM (TMEM 4 x 128 columns) overlap each other's phases.
//
// Two schedules:
//   window_attn_bi_kernel    (batches >= 4) the two units of a tile are the SAME window position and head of TWO
//                            images; the token map and the thread's bias row are fetched once per unit of work and
//                            reused for a chunk of image pairs
//   window_attn_pair_kernel  (batches < 4)  two adjacent windows of one image per tile, one head per item
//
// Algorithmic HBM bytes per unit: 49 * 32 * 2 B * 4 (q, k, v read + o written) = 12,544 B.
#include <cuda_fp16.h>

#include "psw_common.cuh"

namespace psw {

#ifdef PSW_DIAGNOSTICS
constexpr bool kDiag = true;
#else
constexpr bool kDiag = false;
#endif

constexpr int AT_THREADS = 128;
constexpr int AT_PART_BYTES = 128 * 64;            // 128 rows x 64 B (32 bf16), SWIZZLE_64B
constexpr int AT_BUF_BYTES = 3 * AT_PART_BYTES;    // q, k, v
constexpr int AT_TMEM_COLS = 128;
constexpr int AT_P_COL = 0;                        // P (bf16 pairs): TMEM columns [0, 32)
constexpr int AT_O_COL = 32;                       // O (fp32): TMEM columns [32, 96)
constexpr int AT_CTAS_PER_SM = 4;                  // 128 registers per thread (no spills), 4 x 128 TMEM columns
constexpr int AT_FULL_CHUNKS = 13;                 // float4 chunks per bias row: 49 logits padded to 52
constexpr float LOG2E = 1.4426950408889634f;

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
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// bias + softmax numerator of one row: sr = the 49 raw logits of TMEM, bias = the row's additive terms (x log2 e),
// scale_l2 = scale * log2 e.  Writes the un-normalised probabilities as bf16 pairs (key columns 49..63 zero) and, for
// the fp32 schedule that needs it, returns their fp32 sum.  Rows that do not exist (tile rows 49..63 of a unit, the
// second unit of an odd tail) are NOT zeroed: a row of P only feeds the same row of O, which is never stored.
template <int N>
__device__ __forceinline__ float softmax_row(const uint32_t (&sr)[N], const float* bias, float scale_l2, uint32_t (&pk)[32]) {
  float t[N];
#pragma unroll
  for (int j = 0; j < N; ++j) t[j] = fmaf(__uint_as_float(sr[j]), scale_l2, bias[j]);
  float mx = t[N - 1];
#pragma unroll
  for (int j = 0; j + 1 < N; j += 2) mx = fmax3(mx, t[j], t[j + 1]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    uint32_t w = 0u;
    if (2 * k < N) {
      // (packed ex2.approx.bf16x2 / f16x2 were tried: sm_100a issues one MUFU.EX2 per element either way, so they
      // save nothing, and fp16 probabilities against bf16 values are not a legal kind::f16 operand pair)
      const float p0 = fast_exp2(t[2 * k] - mx);
      const float p1 = (2 * k + 1 < N) ? fast_exp2(t[2 * k + 1] - mx) : 0.f;
      sum += p0 + p1;
      w = pack_bf16x2(p0, p1);
    }
    pk[k] = w;
  }
  return sum;
}

struct AttnParams {
  const bf16* qkv;
  bf16* out;
  const float4* bias_full; // [windows per image][heads][AT_FULL_CHUNKS][64 rows] x 4 fp32 (x log2 e) from psw_window_bias_full
  const float* qkv_bias;
  WinGeom g;
  int B, C, heads;
  int hc;                 // pair kernel: heads per work item
  int n_windows;          // B * windows per image
  int n_items;            // pair kernel: work items
  int nslots;             // batch-innermost kernel: ranges per head (grid = nslots * heads)
  int steps_per_head;     // batch-innermost kernel: windows per image * image pairs
  float scale_l2;         // scale * log2(e)
  long long* dbg;         // diagnostics build: per-phase cycle totals of CTA 0 (nullptr otherwise)
  int mode;               // diagnostics build: 0 normal, 1 memory skeleton, 2 no bias loads, 3 no q/k/v loads
  int variant;            // diagnostics build: bit 5 = per-CTA {SM, start, end} -> dbg[8 + 3 * CTA], phase cycles -> dbg[8 + 3 * 1024 + 8 * CTA]
};

// One pipeline step of the window-pair kernel = one head of one window pair.
struct Step {
  int item;               // work item index (for range checks)
  int wp;                 // window pair: windows 2*wp, 2*wp + 1
  int e;                  // head
  int el;                 // head index inside the item, 0 .. hc-1
  int n;                  // running step count of this CTA (parity selects the buffers)
};

template <int WS>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS_PER_SM)
window_attn_pair_kernel(const AttnParams p) {
  constexpr int N = WS * WS;                     // tokens per window (<= 64)
  static_assert(N == 49, "TMEM row load below is written for 49 logits");

  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bufs = smem;                                                  // [2 stages][3][128 x 64 B]
  int* src = reinterpret_cast<int*>(bufs + 2 * AT_BUF_BYTES);            // [3 slots][2 units][64]: this, next, next-next pair
  uint64_t* bars = reinterpret_cast<uint64_t*>(src + 3 * 2 * 64);        // [2]: S ready, O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int C = p.C, heads = p.heads, C3 = 3 * p.C;
  const WinGeom g = p.g;
  const int wpi = g.nWh * g.nWw;
  const int64_t HW = (int64_t)g.H * g.W;
  const int n_hc = heads / p.hc;

  // ---------------------------------------------------------------- one-time setup
  for (int i = tid; i < 2 * AT_BUF_BYTES / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(bufs)[i] = make_uint4(0, 0, 0, 0);          // padding rows must stay finite
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<AT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  // 8 bias values (one 16-byte chunk of a padding token's q / k / v row) as bf16; padding cells are rare
  auto bias_chunk = [&](int ch) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (p.qkv_bias) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.qkv_bias + ch));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.qkv_bias + ch) + 1);
      r = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    return r;
  };

  const int item_begin = (int)((int64_t)p.n_items * blockIdx.x / gridDim.x);
  const int item_end = (int)((int64_t)p.n_items * (blockIdx.x + 1) / gridDim.x);

  // this thread's tile row
  const int unit = tid >> 6;                               // warp-uniform
  const int ti = tid & 63;
  const int ic = ti < N ? ti : 0;
  const int ri = ic / WS, ci = ic - ri * WS;
  // loader role of this thread: 16-byte chunk c of tokens lt0 and lt0 + 32
  const int lc = tid & 3;
  const int lt0 = tid >> 2;

  auto first_step = [&](int item) {
    Step s;
    s.item = item;
    s.wp = item / n_hc;
    s.e = (item - s.wp * n_hc) * p.hc;
    s.el = 0;
    s.n = 0;
    return s;
  };
  auto next_step = [&](Step s) {
    ++s.n;
    if (++s.el < p.hc) { ++s.e; return s; }
    s.el = 0;
    ++s.item;
    if (s.e + 1 < heads) { ++s.e; } else { s.e = 0; ++s.wp; }
    return s;
  };
  // token maps of both windows of pair `wp` -> src[wp % 3] (thread = (unit, token)); entry = global token index
  // b*H*W + h*W + w of the cell's source token, or -1 for a padding cell
  auto prep_src = [&](int wp) {
    const int w = 2 * wp + unit;
    int s = -1;
    if (ti < N && w < p.n_windows) {
      const int b = w / wpi;
      const int wi = w - b * wpi;
      const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
      const int t = source_token(g, wr * WS + ri, wc * WS + ci);
      if (t >= 0) s = b * (int)HW + t;                     // global token index (< 2^31, checked by the host)
    }
    src[(wp % 3) * 128 + tid] = s;
  };
  // per-thread constants of the loader role: rows lt0, lt0+32 of unit 0 and unit 1, chunk lc
  int ld_row[4], ld_dst[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = lt0 + 32 * (k & 1);
    ld_row[k] = t < N ? (k >> 1) * 64 + t : -1;
    const int row = (k >> 1) * 64 + t;
    ld_dst[k] = row * 64 + ((lc ^ ((row >> 1) & 3)) << 4);
  }
  // gather all q/k/v rows of step `st` into stage (st.n & 1)
  auto issue_loads = [&](const Step& st) {
    uint8_t* base = bufs + (st.n & 1) * AT_BUF_BYTES;
    const int* smap = src + (st.wp % 3) * 128;
    const bf16* gq = p.qkv + st.e * 32 + lc * 8;
    const int bch = st.e * 32 + lc * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (ld_row[k] >= 0) {
        const int s = smap[ld_row[k]];
        uint8_t* dst = base + ld_dst[k];
        if (s >= 0) {
          const bf16* grow = gq + (int64_t)s * C3;
          cp_async16(dst, grow);
          cp_async16(dst + AT_PART_BYTES, grow + C);
          cp_async16(dst + 2 * AT_PART_BYTES, grow + 2 * C);
        } else if (2 * st.wp + (k >> 1) < p.n_windows) {   // padding cell of a real window: q/k/v = bias
          *reinterpret_cast<uint4*>(dst) = bias_chunk(bch);
          *reinterpret_cast<uint4*>(dst + AT_PART_BYTES) = bias_chunk(bch + C);
          *reinterpret_cast<uint4*>(dst + 2 * AT_PART_BYTES) = bias_chunk(bch + 2 * C);
        }
      }
    }
  };

  Step cur = first_step(item_begin);
  int src_wp = -1;                                         // highest window pair whose token maps are in smem
  if (item_begin < item_end) {
    prep_src(cur.wp);
    src_wp = cur.wp;
    const Step n1 = next_step(cur);
    if (n1.item < item_end && n1.wp > src_wp) { prep_src(n1.wp); src_wp = n1.wp; }
    __syncthreads();
    issue_loads(cur);
  }
  cp_async_commit();

  // my row of the precomputed bias of a step, requested one step ahead (right after the previous softmax has
  // consumed the registers) so the L2 latency hides behind the P.V MMA, the store and the next S MMA
  float bfull[4 * AT_FULL_CHUNKS];
  auto load_bias = [&](const Step& st) {
    const int w = 2 * st.wp + unit;
    if (ti < N && w < p.n_windows && !(kDiag && p.mode == 2)) {
      const float4* brow = p.bias_full + ((size_t)(w % wpi) * heads + st.e) * (AT_FULL_CHUNKS * 64) + ti;
#pragma unroll
      for (int k = 0; k < AT_FULL_CHUNKS; ++k) {
        const float4 b4 = __ldg(brow + k * 64);
        bfull[4 * k] = b4.x; bfull[4 * k + 1] = b4.y; bfull[4 * k + 2] = b4.z; bfull[4 * k + 3] = b4.w;
      }
    }
  };
#pragma unroll
  for (int k = 0; k < 4 * AT_FULL_CHUNKS; ++k) bfull[k] = 0.f;
  if (item_begin < item_end) load_bias(cur);

  uint32_t par = 0;
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = kDiag && p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  while (cur.item < item_end) {
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (prof) c0 = clock64();
    const Step nxt = next_step(cur);
    const bool has_next = nxt.item < item_end;
    const int my_w = 2 * cur.wp + unit;
    const bool row_valid = (ti < N) && (my_w < p.n_windows);
    // ---- 1. this step's q/k/v (requested one step ago) have landed; one barrier orders them, the previous step's
    //         TMEM reads and the token maps prepared during the previous step
    cp_async_wait<0>();
    fence_async_shared();
    __syncthreads();

    if (prof) c1 = clock64();
    if (kDiag && p.mode == 1) {
      // memory skeleton: prefetch as usual, then copy my row's q chunk to the output position (64 B per row)
      if (has_next) issue_loads(nxt);
      cp_async_commit();
      if (has_next) {
        const Step nn = next_step(nxt);
        if (nn.item < item_end && nn.wp > src_wp) { prep_src(nn.wp); src_wp = nn.wp; }
      }
      const int s = src[(cur.wp % 3) * 128 + unit * 64 + ic];
      if (row_valid && s >= 0) {
        const uint8_t* qrow = bufs + (cur.n & 1) * AT_BUF_BYTES + tid * 64;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (int64_t)s * C + cur.e * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = *reinterpret_cast<const uint4*>(qrow + 16 * c);
      }
      __syncthreads();
      cur = nxt;
      continue;
    }
    // ---- 2. S = Q . K^T (both units at once, block diagonal)
    const uint32_t sq = smem_u32(bufs + (cur.n & 1) * AT_BUF_BYTES);
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      const uint64_t dq = umma_smem_desc(sq, 16, 512, UMMA_SWIZZLE_64B);
      const uint64_t dk = umma_smem_desc(sq + AT_PART_BYTES, 16, 512, UMMA_SWIZZLE_64B);
      umma_ss(tmem_base, dq, dk, idesc, 0);
      umma_ss(tmem_base, dq + 2, dk + 2, idesc, 1);        // head_dim 16..31: +32 B inside the swizzle row
      umma_commit(&bars[0]);
    }
    // while the MMA runs: prefetch the WHOLE next step into the other stage (released by the previous step's PV
    // MMA, which every thread has waited for)
    if (has_next) issue_loads(nxt);
    cp_async_commit();
    mbar_wait(&bars[0], par);
    tc_fence_after();

    if (prof) c2 = clock64();
    // ---- 3. bias + softmax on my row
    float sum;
    {
      uint32_t sr[N];
      const uint32_t s_addr = tmem_base + lane_base + (uint32_t)(unit * 64);
      {
        uint32_t t32[32], t16[16], t1;
        tmem_ld_x32(s_addr, t32);
        tmem_ld_x16(s_addr + 32, t16);
        tmem_ld_x1(s_addr + 48, t1);
        tmem_ld_wait();                                    // one wait for the three loads
#pragma unroll
        for (int k = 0; k < 32; ++k) sr[k] = t32[k];
#pragma unroll
        for (int k = 0; k < 16; ++k) sr[32 + k] = t16[k];
        sr[48] = t1;
      }
      uint32_t pk[32];
      sum = softmax_row<N>(sr, bfull, p.scale_l2, pk);   // fp32 sums: this schedule has no sum MMA
      tmem_st_x32(tmem_base + lane_base + AT_P_COL, pk);
      tmem_st_wait();
    }
    if (has_next) load_bias(nxt);                          // the bias registers are free again
    tc_fence_before();
    __syncthreads();

    if (prof) c3 = clock64();
    // ---- 4. O = P . [V_u0 | V_u1]
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);     // B (V) is MN-major: [key][dim] rows
      const uint32_t sv = sq + 2 * AT_PART_BYTES;
#pragma unroll
      for (int k = 0; k < 4; ++k) {                              // 16 keys per MMA = two 8-key groups of 512 B
        const uint64_t dv = umma_smem_desc(sv + k * 1024, 4096, 512, UMMA_SWIZZLE_64B);
        umma_ts(tmem_base + AT_O_COL, tmem_base + AT_P_COL + k * 8, dv, idesc, k > 0);
      }
      umma_commit(&bars[1]);
    }
    // while the MMA runs: token maps of the step after next (visible to its loader after the next barriers)
    if (has_next) {
      const Step nn = next_step(nxt);
      if (nn.item < item_end && nn.wp > src_wp) { prep_src(nn.wp); src_wp = nn.wp; }
    }
    mbar_wait(&bars[1], par);
    tc_fence_after();

    if (prof) c4 = clock64();
    // ---- 5. normalise and store my output row at the token's un-shifted position
    {
      uint32_t orow[32];
      tmem_ld_x32(tmem_base + lane_base + AT_O_COL + (uint32_t)(unit * 32), orow);
      tmem_ld_wait();
      const int s = src[(cur.wp % 3) * 128 + unit * 64 + ic];
      if (row_valid && s >= 0) {
        const float inv = 1.0f / sum;
        uint4* dst = reinterpret_cast<uint4*>(p.out + (int64_t)s * C + cur.e * 32);
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
    if (prof) {
      const long long c5 = clock64();
      ph[0] += c1 - c0; ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += c5 - c4; ph[5] += 1;
    }
    par ^= 1;
    cur = nxt;
  }
  if (prof)
    for (int k = 0; k < 6; ++k) p.dbg[k] = ph[k];

  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------
// Batch-innermost schedule (production path for batches of 4 images or more).
// The 128-row tile holds the SAME window position and head of TWO images.  A CTA owns ONE head for its whole life and a
// contiguous range of the (window position, image pair) list of that head, image pair fastest: everything that
// depends on the geometry or the head only -- the token map of the window, the thread's 49-entry bias row, the
// padding cells' q / k / v -- is set up once per window position and reused for all image pairs, so a step is nothing
// but the q/k/v gather, the two MMAs, the softmax and the store.  CTA c has head c % heads and range slot c / heads:
// the CTAs of one slot walk through the same windows and images at the same pace, one head each, so the head slices
// that share a 128-byte line of the qkv rows meet in L2.
//
// The two images of a tile share the 64 key columns of S: S = [Q_img0; Q_img1] . K_img^T is issued once per image with
// N = 64 and the OTHER image's 64 output lanes disabled (tcgen05.mma disable-output-lane), so both land in TMEM columns
// [0, 64), O gets its own 64 columns, and the S MMA of step n+1 is issued as soon as the P.V MMA of step n has
// consumed P -- before O(n) is drained and stored.  q / k / v travel with cp.async whose completion is tracked by an
// mbarrier per stage (cp.async.mbarrier.arrive.noinc): no thread ever waits for its own copies and there is no proxy
// fence on the step's critical path -- only the MMA-issuing thread waits for the stage, right before the S MMA that
// reads it.  Per step and CTA:
//   (a) wait S(n)
//   (b) softmax(n): tcgen05.ld S, + bias, max, exp2, row sum (registers), bf16 P -> TMEM;          barrier
//   (c) one thread: P.V MMA(n)        (d) wait P.V(n): P and the stage's q / k / v are free
//   (e) one thread: wait for the q / k / v of step n+1 (requested a whole step ago), S MMA(n+1)
//   (f) q / k / v rows of step n+2 -> the stage just released; token map of the window of step n+3
//   (g) O(n) * 1/sum -> global, while S(n+1) runs on the tensor core
// Padding cells (zero tokens whose q / k / v equal the qkv bias) are the same cells for every step of a window and
// keep their place in a stage buffer, so they are written (generic stores + one proxy fence, before that step's
// copies are requested) only by the first two steps of a window, one per stage.
// ---------------------------------------------------------------------------------------------------
struct BiIt {
  int wi;                 // window position inside an image
  int bp;                 // image pair: images 2*bp, 2*bp + 1
  int ic;                 // running window count of this CTA (selects the per-window shared-memory slot)
  int k;                  // running step count of this CTA (parity selects the stage)
};

// Shared-memory map of the batch-innermost kernel (offsets from the 1024-byte aligned base): two stages of
// [q | k | v] x [128 rows (img*64 + token) x 64 B] SWIZZLE_64B, then the per-window data.
constexpr int BI_MISC = 2 * AT_BUF_BYTES;          // token maps, row offsets, padding rows, barriers, TMEM slot
constexpr int BI_SLOTS = 4;                        // per-window slots: the windows of steps n .. n+3 are alive at once
constexpr int BI_S_COL = 0;                        // S (fp32, 64 columns shared by both images) then P (bf16 pairs, 32 columns)
constexpr int BI_O_COL = 64;                       // O (fp32): image h in columns [64 + 32h, 96 + 32h)

// D[tmem] (+)= A[smem] * B[smem] with the output lanes whose mask bit is set left untouched (disable-output-lane)
__device__ __forceinline__ void umma_ss_masked(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3) : "memory");
}

template <int WS>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS_PER_SM)
window_attn_bi_kernel(const AttnParams p) {
  constexpr int N = WS * WS;
  static_assert(N == 49, "TMEM row load below is written for 49 logits");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  int* src = reinterpret_cast<int*>(smem + BI_MISC);                     // [BI_SLOTS][64]: token index in the image or -1
  int* soff = src + BI_SLOTS * 64;                                       // [BI_SLOTS][64]: token index * 3C (element offset of the qkv row)
  uint4* padrow = reinterpret_cast<uint4*>(soff + BI_SLOTS * 64);        // [3][4]: bf16 qkv bias of this CTA's head
  uint64_t* bars = reinterpret_cast<uint64_t*>(padrow + 12);             // S ready, O ready, stage 0 full, stage 1 full
  uint64_t* full = bars + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int C = p.C, heads = p.heads, C3 = 3 * p.C;
  const WinGeom g = p.g;
  const int64_t HW = (int64_t)g.H * g.W;
  const int BP = (p.B + 1) / 2;                            // steps per window position

  for (int i = tid; i < BI_MISC / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);          // rows 49..63 of a unit: must stay finite
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&full[0], AT_THREADS);
    mbar_init(&full[1], AT_THREADS);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<AT_TMEM_COLS>(tmem_slot);

  // this CTA's head and its range [l0, l1) of the head's (window position, image pair) list
  const int e = blockIdx.x % heads;
  const int range_slot = blockIdx.x / heads;
  const int l0 = (int)((int64_t)p.steps_per_head * range_slot / p.nslots);
  const int nsteps = (int)((int64_t)p.steps_per_head * (range_slot + 1) / p.nslots) - l0;
  if (tid >= 64 && tid < 76) {                             // padding cells: [q|k|v][4 x 16 B] of the head's bf16 bias
    const int j = tid - 64;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (p.qkv_bias) {
      const float* bsrc = p.qkv_bias + (j >> 2) * C + e * 32 + (j & 3) * 8;
      const float4 a = __ldg(reinterpret_cast<const float4*>(bsrc));
      const float4 b = __ldg(reinterpret_cast<const float4*>(bsrc) + 1);
      r = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    padrow[j] = r;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int img = tid >> 6;                                // warp-uniform: which image of the pair my row belongs to
  const int ti = tid & 63;
  const int ic = ti < N ? ti : 0;
  const int ri = ic / WS, ci = ic - ri * WS;
  const int lc = tid & 3;
  const int lt0 = tid >> 2;

  auto advance = [&](BiIt& it) {
    ++it.k;
    if (++it.bp == BP) { it.bp = 0; ++it.wi; ++it.ic; }
  };
  // token map of a window position -> shared-memory slot (ic % BI_SLOTS), threads 0..63
  auto prep_item = [&](const BiIt& it) {
    if (tid < 64) {
      int t = -1;
      if (tid < N) {
        const int wr = it.wi / g.nWw, wc = it.wi - wr * g.nWw;
        t = source_token(g, wr * WS + ri, wc * WS + ci);
      }
      const int slot = it.ic & (BI_SLOTS - 1);
      src[slot * 64 + tid] = t;
      soff[slot * 64 + tid] = t * C3;
    }
  };
  // loader role of this thread: 16-byte chunk lc of tokens lt0 and lt0 + 32 of both images; row = img*64 + token of
  // 64-byte rows (SWIZZLE_64B: chunk ^ ((row >> 1) & 3)), the same offset in the q, k and v parts
  const bool ld_ok1 = lt0 + 32 < N;                        // token lt0 (< 32) always exists; lt0 + 32 only below 49
  uint32_t ld_dst[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int row = (k >> 1) * 64 + lt0 + 32 * (k & 1);
    ld_dst[k] = (uint32_t)(row * 64 + ((lc ^ ((row >> 1) & 3)) << 4));
  }
  const int64_t img_stride = HW * C3;                      // elements per image of the qkv tensor
  const bool no_loads = kDiag && p.mode == 3;
  auto issue_loads = [&](const BiIt& it) {                  // all q / k / v rows of a step into stage (k & 1)
    const int stage = it.k & 1;
    const int* so = soff + (it.ic & (BI_SLOTS - 1)) * 64;
    const int o0 = so[lt0];
    const int o1 = ld_ok1 ? so[lt0 + 32] : 0;
    const bf16* g0 = p.qkv + (int64_t)(2 * it.bp) * img_stride + e * 32 + lc * 8;
    const bool img1 = 2 * it.bp + 1 < p.B;
    uint8_t* base = smem + stage * AT_BUF_BYTES;
    if (it.bp < 2 || it.k < 2) {                           // first visit of this window to this stage: its padding cells
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((k & 1) && !ld_ok1) continue;
        if (((k & 1) ? o1 : o0) < 0) {
          uint8_t* dst = base + ld_dst[k];
          *reinterpret_cast<uint4*>(dst) = padrow[lc];
          *reinterpret_cast<uint4*>(dst + AT_PART_BYTES) = padrow[4 + lc];
          *reinterpret_cast<uint4*>(dst + 2 * AT_PART_BYTES) = padrow[8 + lc];
        }
      }
      fence_async_shared();                                // before this step's copies are requested: nothing young in flight
    }
    if (!no_loads) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((k & 1) && !ld_ok1) continue;
        if ((k >> 1) && !img1) continue;
        const int o = (k & 1) ? o1 : o0;
        if (o >= 0) {
          uint8_t* dst = base + ld_dst[k];
          const bf16* grow = g0 + ((k >> 1) ? img_stride : 0) + o;
          cp_async16(dst, grow);
          cp_async16(dst + AT_PART_BYTES, grow + C);
          cp_async16(dst + 2 * AT_PART_BYTES, grow + 2 * C);
        }
      }
    }
    cp_async_mbar_arrive_noinc(&full[stage]);              // arrives when this thread's copies have landed
  };
  // my row of the window's precomputed bias: [window][head][chunk][row] float4
  float bias[4 * AT_FULL_CHUNKS];
  auto load_bias = [&](const BiIt& it) {
    if (ti < N && !(kDiag && p.mode == 2)) {
      const float4* brow = p.bias_full + ((size_t)it.wi * heads + e) * (AT_FULL_CHUNKS * 64) + ti;
#pragma unroll
      for (int k = 0; k < AT_FULL_CHUNKS; ++k) {
        const float4 b4 = __ldg(brow + k * 64);
        bias[4 * k] = b4.x; bias[4 * k + 1] = b4.y; bias[4 * k + 2] = b4.z; bias[4 * k + 3] = b4.w;
      }
    }
  };
#pragma unroll
  for (int k = 0; k < 4 * AT_FULL_CHUNKS; ++k) bias[k] = 0.f;

  // one thread: wait for the stage, then S(img) = [Q_img0; Q_img1] . K_img^T with the other image's lanes masked
  auto issue_s_mma = [&](int k) {
    const int stage = k & 1;
    mbar_wait(&full[stage], (uint32_t)(k >> 1) & 1u);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    const uint32_t sq = smem_u32(smem + stage * AT_BUF_BYTES);
    const uint64_t da = umma_smem_desc(sq, 16, 512, UMMA_SWIZZLE_64B);
    const uint64_t db0 = umma_smem_desc(sq + AT_PART_BYTES, 16, 512, UMMA_SWIZZLE_64B);
    const uint64_t db1 = umma_smem_desc(sq + AT_PART_BYTES + 64 * 64, 16, 512, UMMA_SWIZZLE_64B);
    umma_ss_masked(tmem_base + BI_S_COL, da, db0, idesc, 0, 0u, 0u, ~0u, ~0u);          // image 0: lanes 0..63
    umma_ss_masked(tmem_base + BI_S_COL, da + 2, db0 + 2, idesc, 1, 0u, 0u, ~0u, ~0u);  // head_dim 16..31: +32 B in the swizzle row
    umma_ss_masked(tmem_base + BI_S_COL, da, db1, idesc, 0, ~0u, ~0u, 0u, 0u);          // image 1: lanes 64..127
    umma_ss_masked(tmem_base + BI_S_COL, da + 2, db1 + 2, idesc, 1, ~0u, ~0u, 0u, 0u);
    umma_commit(&bars[0]);
  };

  long long t_start = 0, c_start = 0, c_loop = 0;
  if (kDiag && (p.variant & 32) && tid == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start)); c_start = clock64(); }
  // software pipeline over steps: cur = n (computed now), nxt = n+1, nn = n+2, n3 = n+3
  BiIt cur, nxt, nn, n3;
  cur.wi = l0 / BP; cur.bp = l0 - cur.wi * BP; cur.ic = 0; cur.k = 0;
  nxt = cur; advance(nxt);
  nn = nxt; advance(nn);
  n3 = nn; advance(n3);
  if (nsteps > 0) {
    prep_item(cur);
    if (nxt.k < nsteps && nxt.ic != cur.ic) prep_item(nxt);
    if (nn.k < nsteps && nn.ic != nxt.ic) prep_item(nn);
    __syncthreads();
    issue_loads(cur);
    if (nxt.k < nsteps) issue_loads(nxt);
    load_bias(cur);
    if (tid == 0) issue_s_mma(0);
  }

  uint32_t par = 0;
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = kDiag && p.dbg != nullptr && (blockIdx.x == 0 || (p.variant & 32)) && tid == 0;
  if (prof) c_loop = clock64();
  while (cur.k < nsteps) {
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (prof) c0 = clock64();
    const bool has_next = nxt.k < nsteps;
    const int my_b = 2 * cur.bp + img;
    const int ts = src[(cur.ic & (BI_SLOTS - 1)) * 64 + ic];   // my token
    // ---- (a) S(n) is ready
    mbar_wait(&bars[0], par);
    tc_fence_after();
    if (prof) c1 = clock64();
    // ---- (b) bias + softmax on my row
    float sum;
    {
      uint32_t sr[N];
      const uint32_t s_addr = tmem_base + lane_base + BI_S_COL;
      {
        uint32_t t32[32], t16[16], t1;
        tmem_ld_x32(s_addr, t32);
        tmem_ld_x16(s_addr + 32, t16);
        tmem_ld_x1(s_addr + 48, t1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) sr[k] = t32[k];
#pragma unroll
        for (int k = 0; k < 16; ++k) sr[32 + k] = t16[k];
        sr[48] = t1;
      }
      uint32_t pk[32];
      sum = softmax_row<N>(sr, bias, p.scale_l2, pk);
      tmem_st_x32(tmem_base + lane_base + BI_S_COL, pk);
      tmem_st_wait();
    }
    if (has_next && nxt.ic != cur.ic) load_bias(nxt);      // new window: the bias registers are free again
    tc_fence_before();
    __syncthreads();
    if (prof) c2 = clock64();
    // ---- (c) O = P . [V_img0 | V_img1]
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);       // B (V) is MN-major: [key][dim] rows
      const uint32_t sv = smem_u32(smem + (cur.k & 1) * AT_BUF_BYTES + 2 * AT_PART_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                                // 16 keys per MMA = two 8-key groups of 512 B
        const uint64_t dvd = umma_smem_desc(sv + k * 1024, 4096, 512, UMMA_SWIZZLE_64B);
        umma_ts(tmem_base + BI_O_COL, tmem_base + BI_S_COL + k * 8, dvd, idesc, k > 0);
      }
      umma_commit(&bars[1]);
    }
    // ---- (d) P.V(n) done: P and this stage's q / k / v are free
    mbar_wait(&bars[1], par);
    tc_fence_after();
    if (prof) c3 = clock64();
    // ---- (e) S(n+1) as soon as its q / k / v (requested a whole step ago) have landed
    if (tid == 0 && has_next) issue_s_mma(nxt.k);
    // ---- (f) q / k / v rows of step n+2 into the stage just released; token map of the window of step n+3
    if (nn.k < nsteps) issue_loads(nn);
    if (n3.k < nsteps && n3.ic != nn.ic) prep_item(n3);
    if (prof) c4 = clock64();
    // ---- (g) normalise and store my output row at the token's un-shifted position
    {
      uint32_t orow[32];
      tmem_ld_x32(tmem_base + lane_base + BI_O_COL + (uint32_t)(img * 32), orow);
      tmem_ld_wait();
      if (ti < N && my_b < p.B && ts >= 0) {
        const float inv = 1.0f / sum;
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((int64_t)my_b * HW + ts) * C + e * 32);
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
    tc_fence_before();       // the next barrier orders these TMEM reads before the next P.V MMA overwrites O
    if (prof) {
      const long long c5 = clock64();
      ph[0] += c1 - c0; ph[1] += c2 - c1; ph[2] += c3 - c2; ph[3] += c4 - c3; ph[4] += c5 - c4; ph[5] += 1;
    }
    par ^= 1;
    cur = nxt;
    nxt = nn;
    nn = n3;
    advance(n3);
  }
  if (prof && blockIdx.x == 0)
    for (int k = 0; k < 6; ++k) p.dbg[k] = ph[k];
  if (kDiag && (p.variant & 32) && tid == 0 && p.dbg != nullptr) {
    for (int k = 0; k < 6; ++k) p.dbg[8 + 3 * 1024 + 8 * blockIdx.x + k] = ph[k];
    p.dbg[8 + 3 * 1024 + 8 * blockIdx.x + 6] = clock64() - c_start;
    p.dbg[8 + 3 * 1024 + 8 * blockIdx.x + 7] = clock64() - c_loop;
    long long t_end;
    uint32_t smid;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.dbg[8 + 3 * blockIdx.x] = smid;
    p.dbg[9 + 3 * blockIdx.x] = t_start;
    p.dbg[10 + 3 * blockIdx.x] = t_end;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<AT_TMEM_COLS>(tmem_base);
  }
}

static size_t attn_tc_smem_bytes() {
  size_t b = 1024;                                   // alignment slack
  b += 2 * AT_BUF_BYTES;
  b += 3072 + 512;                                   // token maps + row offsets, padding rows, barriers, TMEM slot
  // keep the CTA count per SM at AT_CTAS_PER_SM (register budget 65536 / (3 * 128) = 170 per thread)
  const size_t floor_bytes = (size_t)(233472 / (AT_CTAS_PER_SM + 1)) - 1024 + 16;
  return b < floor_bytes ? floor_bytes : b;
}

// ---------------------------------------------------------------------------------------------------
// Full additive bias of every (window of one image, head): bias[i][j] = hav(uv_i, uv_j) * alpha[idx(i,j)][head] +
// beta[idx(i,j)][head] (+ mask[window][i][j] in planar mode), fp32, MULTIPLIED BY log2(e) (the kernels evaluate
// softmax with exp2), laid out for the kernels' row-per-lane reads: [window][head][chunk k = j / 4][row i (64)][j % 4].
// Depends on the geometry and on the block's alpha / beta, not on the batch: built once per block and resolution,
// read (L2-resident) by every image.  Same formulas and fp32 evaluation order as the parity kernel
// (lzx/models/great_circle.py:82-86, reference :241-272).
// ---------------------------------------------------------------------------------------------------
__global__ void bias_full_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ uv,
                                 const float* __restrict__ mask, float* __restrict__ table, WinGeom g, int heads) {
  // One CTA per (window, head group): the great-circle distances depend on the window only, so they are evaluated once
  // (symmetric: 1,225 pairs for a 7 x 7 window) and every head of the group re-uses them -- a CTA per (window, head)
  // spent its time in sinf / asinf, 3x (stage 0) to 24x (stage 3) redundantly.
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1;
  __shared__ float su[64], sv[64], scv[64];
  __shared__ float shav[64 * 65];                               // [i][j], pitch 65
  __shared__ uint16_t sidx[AT_FULL_CHUNKS * 64 * 4];            // table index (ri - rj + ws - 1) * tw + (ci - cj + ws - 1), 0xffff outside
  const int wi = blockIdx.x;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    const int r = t / ws, c = t - r * ws;
    const int s = source_token(g, wr * ws + r, wc * ws + c);
    float uu = 0.f, vv = 0.f;
    if (s >= 0 && uv != nullptr) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
    su[t] = uu; sv[t] = vv; scv[t] = cosf(vv);
  }
  for (int q = threadIdx.x; q < AT_FULL_CHUNKS * 64 * 4; q += blockDim.x) {
    const int k = q >> 8, i = (q >> 2) & 63, j = 4 * k + (q & 3);
    uint16_t idx = 0xffffu;
    if (i < N && j < N) {
      const int ri = i / ws, ci = i - ri * ws, rj = j / ws, cj = j - rj * ws;
      idx = (uint16_t)((ri - rj + ws - 1) * tw + (ci - cj + ws - 1));
    }
    sidx[q] = idx;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < N * N; p += blockDim.x) {
    const int i = p / N, j = p - i * N;
    if (j < i) continue;                                        // upper triangle; mirrored below
    float d = 0.f;
    if (uv != nullptr) {
      const float sdv = sinf(0.5f * fabsf(sv[j] - sv[i]));
      const float sdu = sinf(0.5f * (su[j] - su[i]));
      const float a = sdv * sdv + (scv[j] * scv[i]) * (sdu * sdu);
      d = asinf(sqrtf(fminf(a, 1.0f))) * 2.0f;
    }
    shav[i * 65 + j] = d;
    shav[j * 65 + i] = d;
  }
  __syncthreads();
  for (int e = blockIdx.y; e < heads; e += gridDim.y) {
    float* out = table + ((size_t)wi * heads + e) * (AT_FULL_CHUNKS * 64 * 4);
    for (int q = threadIdx.x; q < AT_FULL_CHUNKS * 64 * 4; q += blockDim.x) {
      const int idx = sidx[q];
      float b = 0.f;
      if (idx != 0xffff) {
        const int k = q >> 8, i = (q >> 2) & 63, j = 4 * k + (q & 3);
        b = fmaf(shav[i * 65 + j], alpha[idx * heads + e], beta[idx * heads + e]);
        if (mask != nullptr) b += mask[((size_t)wi * N + i) * N + j];
        b *= LOG2E;
      }
      out[q] = b;
    }
  }
}

int window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask, void* table, int H, int W,
                     int heads, int window, int shift, int pano, cudaStream_t st) {
  PSW_REQUIRE(window * window <= 4 * AT_FULL_CHUNKS && window * window <= 64, PSW_ERR_UNSUPPORTED,
              "psw_window_bias_full: window %d too large", window);
  WinGeom g = make_geom(H, W, window, shift, pano);
  const int windows = g.nWh * g.nWw;
  int groups = (4 * num_sms() + windows - 1) / windows;         // head groups per window: enough CTAs for ~4 per SM
  if (groups > heads) groups = heads;
  if (groups < 1) groups = 1;
  bias_full_kernel<<<dim3(windows, groups), 256, 0, st>>>(alpha, beta, pano ? uv : nullptr, mask, (float*)table, g, heads);
  return launch_status("bias_full_kernel");
}

template <typename K>
static int launch_attn(K kern, const AttnParams& p, int grid, size_t smem, cudaStream_t st, const char* name) {
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  kern<<<grid, AT_THREADS, smem, st>>>(p);
  return launch_status(name);
}

// `variant` (diagnostics build only; 0 in production): 15 forces the window-pair schedule.
int window_attn_tc(const bf16* qkv, bf16* out, const float* qkv_bias, const void* bias_full, int B, int H, int W, int C,
                   int heads, int window, int shift, int pano, float scale, long long* dbg, int mode, int variant,
                   cudaStream_t st) {
  PSW_REQUIRE(window == 7 && C / heads == 32, PSW_ERR_UNSUPPORTED,
              "psw_window_attn_full_fwd: the tcgen05 kernel is instantiated for window 7 and head_dim 32 (every shipped "
              "PanoSwin config); got window %d head_dim %d -- use psw_window_attn_fwd (generic route)", window, C / heads);
  PSW_REQUIRE(bias_full != nullptr, PSW_ERR_BAD_ARG, "psw_window_attn_full_fwd: needs the bias table of psw_window_bias_full");
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_full_fwd: too many tokens");
  AttnParams p;
  p.qkv = qkv; p.out = out; p.qkv_bias = qkv_bias;
  p.bias_full = (const float4*)bias_full;
  p.g = make_geom(H, W, window, shift, pano);
  p.B = B; p.C = C; p.heads = heads; p.scale_l2 = scale * LOG2E;
  p.dbg = kDiag ? dbg : nullptr;
  p.mode = kDiag ? mode : 0;
  p.variant = kDiag ? variant : 0;
  p.n_windows = B * p.g.nWh * p.g.nWw;
  p.hc = 1; p.n_items = 0; p.nslots = 1; p.steps_per_head = 0;
  const int forced = kDiag ? (variant & 15) : 0;
  const size_t smem = attn_tc_smem_bytes();
  const int ctas = num_sms() * AT_CTAS_PER_SM;
  const bool batch_inner = B >= 4 && forced != 15 && p.mode != 1;
  if (!batch_inner) {
    // window-pair schedule, one head per work item: sharing a window pair's token maps between the heads of a wider
    // item saves little (measured: equal at stage 0), while single-head items balance the small launches of the late
    // stages better (items are dealt out in contiguous ranges; -10% at stage 3)
    p.n_items = ((p.n_windows + 1) / 2) * heads;
    return launch_attn(window_attn_pair_kernel<7>, p, ctas < p.n_items ? ctas : p.n_items, smem, st, "window_attn_pair_kernel");
  }
  // every resident CTA gets one head and an equal share of that head's (window position, image pair) list
  const int64_t steps = (int64_t)p.g.nWh * p.g.nWw * ((B + 1) / 2);
  PSW_REQUIRE(steps < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_window_attn_full_fwd: too many windows");
  p.steps_per_head = (int)steps;
  p.nslots = ctas / heads;
  if (p.nslots < 1) p.nslots = 1;
  if (p.nslots > steps) p.nslots = (int)steps;
  return launch_attn(window_attn_bi_kernel<7>, p, p.nslots * heads, smem, st, "window_attn_bi_kernel");
}

}  // namespace psw
