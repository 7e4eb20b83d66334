// K3 (throughput path): fused pano-shift + window-partition + multi-head attention + window-reverse +
// un-shift for bf16 activations on sm_100a.  Replaces WindowTransition / pad_x / window_partition /
// BasicWindowAttention core / window_reverse of the reference
// (simple_panoswin_transformer.py:376-409, :486-491, :64-92, :290-308) in one pass over HBM.
//
// Work decomposition.  A "unit" is one (window, head): 49 tokens x head_dim 32.  The tensor-core tile has
// 128 rows = two units: the SAME head of two consecutive windows (w0 = 2k, w1 = 2k+1).  A work item is a
// window pair x a chunk of HC heads, so everything that depends only on the windows (token maps, the
// great-circle distance rows) is fetched once per item and reused by its HC steps:
//   S[128x128] = [Q_u0;Q_u1] . [K_u0;K_u1]^T      tcgen05.mma M=128 N=128 K=32; only the two diagonal 64x64
//                                                 blocks are used (the tensor pipe is far from the bound)
//   O[128x64]  = P[128x64 keys] . [V_u0 | V_u1]   tcgen05.mma M=128 N=64 K=64, P read from TMEM (bf16),
//                                                 V consumed MN-major straight from its [key][dim] rows;
//                                                 rows of unit u use output columns [32u, 32u+32)
// Thread r owns tile row r (TMEM lane r): tcgen05.ld of its 49 logits, + d(i,j)*alpha[idx]+beta[idx]
// (d from the fp16 distance table staged in shared memory, tables per head in shared memory), softmax in
// registers (exp2, fp32), un-normalised bf16 P back to TMEM, finally O row * 1/sum stored with 128-bit
// stores at the token's UN-shifted position.  q/k/v rows are gathered with cp.async (16 B) straight from the
// un-shifted [B,H,W,3C] qkv tensor into the 64B-swizzled UMMA layout: pano shift with longitude wrap-around,
// the odd-W zero column, window padding (padding tokens = qkv bias) and partition are address arithmetic
// (psw::source_token).  q/k/v
// buffers are double-buffered per CTA: the whole next step is prefetched while the S MMA of the current one runs,
// the row's bias is computed in the same window; 4 CTAs per SM (TMEM 4 x 128 columns) overlap each other's
// MMA / softmax / store phases.
//
// Algorithmic HBM bytes per unit: 49 * 32 * 2 B * 4 (q, k, v read + o written) = 12,544 B.
#include <cuda_fp16.h>

#include "psw_common.cuh"

namespace psw {

int attn_debug_hc();          // diagnostics: force the heads-per-item choice (0 = heuristic)


constexpr int AT_THREADS = 128;
constexpr int AT_PART_BYTES = 128 * 64;            // 128 rows x 64 B (32 bf16), SWIZZLE_64B
constexpr int AT_BUF_BYTES = 3 * AT_PART_BYTES;    // q, k, v
constexpr int AT_HAV_PITCH = 56;                   // halfs per distance-table row (112 B = 7 x 16 B)
constexpr int AT_TMEM_COLS = 128;
constexpr int AT_P_COL = 0;                        // P (bf16x2 packed): TMEM columns [0, 32)
constexpr int AT_O_COL = 32;                       // O (fp32): TMEM columns [32, 96)
constexpr int AT_SUM_COL = 96;                     // row sums of P (fp32, batch-innermost kernel): TMEM columns [96, 112)
constexpr int AT_CTAS_PER_SM = 4;                 // 128 registers per thread (no spills), 4 x 128 TMEM columns
constexpr int AT_TAB_PITCH = 39;                   // half2 (alpha, beta) words per table row: 13 used; 39 = 7 mod 32 makes
                                                   // the row-per-lane lookups bank-conflict-free
constexpr int AT_TAB_WORDS = 508;                  // 13 * 39 = 507 words per head, padded to a multiple of 4 (16 B)
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
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Great-circle distance table: hav[win][i][j] = haversine22(uv_i, uv_j) for the 49 tokens of every window
// of ONE image (it depends on the window position only, not on batch / head / weights), fp16, row pitch 56.
// Built once per (H, W, shift) by the host and kept L2-resident; padding tokens sit at uv = (0, 0)
// (reference :486-491, :344-347).  Formula and evaluation order: lzx/models/great_circle.py:82-86, fp32.
// ---------------------------------------------------------------------------------------------------
__global__ void hav_table_kernel(const float* __restrict__ uv, __half* __restrict__ table, WinGeom g) {
  const int ws = g.ws, N = ws * ws;
  __shared__ float su[64], sv[64], scv[64];
  const int wi = blockIdx.x;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    const int r = t / ws, c = t - r * ws;
    const int s = source_token(g, wr * ws + r, wc * ws + c);
    float uu = 0.f, vv = 0.f;
    if (s >= 0) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
    su[t] = uu; sv[t] = vv; scv[t] = cosf(vv);
  }
  __syncthreads();
  __half* out = table + (size_t)wi * N * AT_HAV_PITCH;
  for (int q = threadIdx.x; q < N * AT_HAV_PITCH; q += blockDim.x) {
    const int i = q / AT_HAV_PITCH, j = q - i * AT_HAV_PITCH;
    float d = 0.f;
    if (j < N) {
      const float sdv = sinf(0.5f * fabsf(sv[j] - sv[i]));
      const float sdu = sinf(0.5f * (su[j] - su[i]));
      const float a = sdv * sdv + (scv[j] * scv[i]) * (sdu * sdu);
      d = asinf(sqrtf(fminf(a, 1.0f))) * 2.0f;
    }
    out[q] = __float2half_rn(d);
  }
}

struct AttnParams {
  const bf16* qkv;
  bf16* out;
  const float* alpha;
  const float* beta;
  const __half2* tables;  // [heads][AT_TAB_WORDS] packed (alpha, beta) from psw_window_bias_tables
  const float4* bias_full; // FULL kernels: [windows per image][heads][AT_FULL_CHUNKS][64 rows] x 4 fp32 from psw_window_bias_full
  const float* qkv_bias;
  const __half* hav;      // [wpi][N][56] or nullptr (planar mode: d == 0)
  const float* mask;
  WinGeom g;
  int B, C, heads;
  int hc;                 // heads per work item (divides heads)
  int n_windows;          // B * windows per image
  int n_items;            // ceil(n_windows / 2) * (heads / hc)
  float scale;
  long long* dbg;         // diagnostics: per-phase cycle totals of CTA 0 (nullptr in production)
  int mode;               // diagnostics: 0 = normal, 1 = memory skeleton (same gathers and stores, no MMA / softmax)
};

// One pipeline step = one head of one window pair.
struct Step {
  int item;               // work item index (for range checks)
  int wp;                 // window pair: windows 2*wp, 2*wp + 1
  int e;                  // head
  int el;                 // head index inside the item, 0 .. hc-1
  int n;                  // running step count of this CTA (parity selects the buffers)
};

// FULL: the whole additive bias of a (window, head) -- great-circle term, relative-position term and, in planar mode,
// the shift mask -- comes precomputed from p.bias_full (psw_window_bias_full): 13 coalesced 16-byte loads per row
// instead of ~300 instructions of table lookups per row and step.
template <int WS, bool HAS_MASK, bool FULL>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS_PER_SM)
window_attn_tc_kernel(const AttnParams p) {
  constexpr int N = WS * WS;                     // tokens per window (<= 64)
  constexpr int TW = 2 * WS - 1;                 // relative-position table width (13)
  constexpr int TAB = TW * TW;
  constexpr int TP = AT_TAB_PITCH;               // smem row pitch of the table: bank-conflict-free for row-per-lane reads
  constexpr int TABS = AT_TAB_WORDS;             // half2 words per table slot
  static_assert(N <= 64, "window too large for the 64-row unit tile");
  static_assert(TP >= TW, "table pitch too small");

  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bufs = smem;                                                  // [2 stages][3][128 x 64 B]
  __half2* tab = reinterpret_cast<__half2*>(bufs + 2 * AT_BUF_BYTES);    // [2 stages][TABS] packed (alpha, beta)
  int* src = reinterpret_cast<int*>(tab + 2 * TABS);                     // [3 slots][2 units][64]: this, next, next-next pair
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
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
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
  // token maps of both windows of pair `wp` -> src[wp & 3] (thread = (unit, token)); entry = global token index
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
  // gather all q/k/v rows and the per-head tables of step `st` into stage (st.n & 1)
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
    // per-head tables: one contiguous 2032-byte block, 127 x 16 B
    if (!FULL && tid < TABS / 4) cp_async16(tab + (st.n & 1) * TABS + 4 * tid, p.tables + (size_t)st.e * TABS + 4 * tid);
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

  uint4 hreg[AT_HAV_PITCH / 8];                            // my distance-table row (fp16), kept across the item's heads
#pragma unroll
  for (int k = 0; k < AT_HAV_PITCH / 8; ++k) hreg[k] = make_uint4(0, 0, 0, 0);
  int hav_wp = -1;

  // FULL: my row of the precomputed bias of a step, requested one step ahead (right after the previous softmax has
  // consumed the registers) so the L2 latency hides behind the P.V MMA, the store and the next S MMA
  float bfull[FULL ? 4 * AT_FULL_CHUNKS : 1];
  auto load_bias = [&](const Step& st) {
    if constexpr (FULL) {
      const int w = 2 * st.wp + unit;
      if (ti < N && w < p.n_windows && p.mode != 2) {        // mode 2 (diagnostics): bias loads skipped
        const float4* brow = p.bias_full + ((size_t)(w % wpi) * heads + st.e) * (AT_FULL_CHUNKS * 64) + ti;
#pragma unroll
        for (int k = 0; k < AT_FULL_CHUNKS; ++k) {
          const float4 b4 = __ldg(brow + k * 64);
          bfull[4 * k] = b4.x; bfull[4 * k + 1] = b4.y; bfull[4 * k + 2] = b4.z; bfull[4 * k + 3] = b4.w;
        }
      }
    }
  };
  if constexpr (FULL) {
#pragma unroll
    for (int k = 0; k < 4 * AT_FULL_CHUNKS; ++k) bfull[k] = 0.f;
    if (item_begin < item_end) load_bias(cur);
  }

  uint32_t par = 0;
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  while (cur.item < item_end) {
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (prof) c0 = clock64();
    const Step nxt = next_step(cur);
    const bool has_next = nxt.item < item_end;
    const int my_w = 2 * cur.wp + unit;
    const bool row_valid = (ti < N) && (my_w < p.n_windows);
    // ---- 1. this step's q/k/v + tables (requested one step ago) have landed; one barrier orders them, the
    //         previous step's TMEM reads and the token maps prepared during the previous step
    cp_async_wait<0>();
    fence_async_shared();
    __syncthreads();

    if (prof) c1 = clock64();
    if (p.mode == 1) {
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
    // MMA, which every thread has waited for) and fetch my distance row when the window pair changed
    if (has_next) issue_loads(nxt);
    cp_async_commit();
    if (!FULL && p.hav != nullptr && hav_wp != cur.wp) {
      hav_wp = cur.wp;
      if (row_valid) {
        const uint4* grow = reinterpret_cast<const uint4*>(p.hav + ((size_t)(my_w % wpi) * N + ti) * AT_HAV_PITCH);
#pragma unroll
        for (int k = 0; k < AT_HAV_PITCH / 8; ++k) hreg[k] = __ldg(grow + k);
      }
    }
    // still inside the MMA window: my row's bias d(i,j) * alpha[idx] + beta[idx] (+ mask) — independent of S.
    // Table entries of one key row are fetched as a batch of 7 before they are consumed (LDS latency overlaps).
    float bia[FULL ? 1 : N];
    if constexpr (FULL) {
      bia[0] = 0.f;
    } else {
      const __half2* trow = tab + (cur.n & 1) * TABS + (ri + WS - 1) * TP + (ci + WS - 1);
      const float* mrow = nullptr;
      if constexpr (HAS_MASK) mrow = p.mask + ((int64_t)((my_w < p.n_windows ? my_w : 0) % wpi) * N + ic) * N;
      const uint32_t* hw = reinterpret_cast<const uint32_t*>(hreg);
#pragma unroll
      for (int rj = 0; rj < WS; ++rj) {
        __half2 ab[WS];
#pragma unroll
        for (int cj = 0; cj < WS; ++cj) ab[cj] = trow[-(rj * TP + cj)];
#pragma unroll
        for (int cj = 0; cj < WS; ++cj) {
          const int j = rj * WS + cj;
          const __half2 hh = *reinterpret_cast<const __half2*>(&hw[j >> 1]);
          const float hv = (j & 1) ? __high2float(hh) : __low2float(hh);
          bia[j] = fmaf(hv, __low2float(ab[cj]), __high2float(ab[cj]));
          if constexpr (HAS_MASK) bia[j] += __ldg(mrow + j);
        }
      }
    }
    mbar_wait(&bars[0], par);
    tc_fence_after();

    if (prof) c2 = clock64();
    // ---- 3. bias + softmax on my row
    float sum = 1.f;
    {
      uint32_t sr[N];
      const uint32_t s_addr = tmem_base + lane_base + (uint32_t)(unit * 64);
      {
        static_assert(N == 49, "TMEM row load below is written for 49 logits");
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
      float t[N];
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        t[j] = fmaf(__uint_as_float(sr[j]), p.scale, FULL ? bfull[j] : bia[j]);
        mx = fmaxf(mx, t[j]);
      }
      const float mneg = -mx * LOG2E;
      sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float p0 = 0.f, p1 = 0.f;
        if (2 * k < N) { p0 = fast_exp2(fmaf(t[2 * k], LOG2E, mneg)); sum += p0; }
        if (2 * k + 1 < N) { p1 = fast_exp2(fmaf(t[2 * k + 1], LOG2E, mneg)); sum += p1; }
        pk[k] = row_valid ? pack_bf16x2(p0, p1) : 0u;
      }
      tmem_st_x32(tmem_base + lane_base + AT_P_COL, pk);
      tmem_st_wait();
    }
    if (FULL && has_next) load_bias(nxt);                  // the bias registers are free again
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AT_TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// Batch-innermost variant (production path for batches of 4 images or more, FULL bias only).
// The 128-row tile holds the SAME window position and head of TWO images; a work item is (window position, head)
// and its steps walk over the image pairs.  Everything that depends on the geometry or the head only -- the token
// map of the window and the thread's 49-entry bias row -- is fetched once per item and kept (shared memory /
// registers) for all of its steps, so a step is nothing but the q/k/v gather, the two MMAs, the softmax and the
// store.  An item is cut into units of p.hc image pairs (chosen by the host for balance); units are numbered (window, chunk, head) with the
// head fastest and CTA c runs units c, c + grid, c + 2 grid, ...: at any moment the resident CTAs work on ALL heads of
// the same windows and images, so the head slices sharing a 128-byte line of the qkv rows meet in L2 (a contiguous
// range per CTA separates them by ~40 us and doubles the DRAM reads -- measured).
// ---------------------------------------------------------------------------------------------------
struct BiStep {
  int u;                  // unit index: (window position * chunks + chunk) * heads + head
  int item;               // window position * heads + head
  int bp;                 // image pair: images 2*bp, 2*bp + 1
  int bp_end;             // end of the unit's image-pair range
  int n;                  // running step count of this CTA (parity selects the q/k/v stage)
  int ic;                 // running unit count of this CTA (parity selects the token-map slot)
};

// GATHER: the q/k/v rows are fetched by TMA (cp.async.bulk.tensor tile::gather4: four 64-byte token rows per
// instruction, written with the 64-byte swizzle the MMA descriptors expect; tools/probes/tma_gather4_probe.cu) instead
// of per-thread cp.async: 78 instructions per step, no address arithmetic per 16-byte chunk, no proxy fence, and only
// the MMA-issuing thread waits for the data.  Needs the qkv tensor to carry one extra row (index B*H*W) holding the
// bf16 qkv bias, which padding cells gather (psw_window_attn_full_fwd: qkv_rows == B*H*W + 1).
__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}

template <int WS, bool GATHER>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS_PER_SM)
window_attn_bi_kernel(const AttnParams p, const __grid_constant__ CUtensorMap map_qkv) {
  constexpr int N = WS * WS;
  static_assert(N == 49, "TMEM row load below is written for 49 logits");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* bufs = smem;                                                  // [2 stages][3][128 x 64 B]
  int* src = reinterpret_cast<int*>(bufs + 2 * AT_BUF_BYTES);            // [2 item slots][64]: token index in the image or -1
  uint64_t* bars = reinterpret_cast<uint64_t*>(src + 2 * 64);            // [2]: S ready, O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  uint8_t* ones = bufs + 2 * AT_BUF_BYTES + 1024;                        // 1 KB of bf16 1.0: B operand of the row-sum MMA
  uint64_t* full = reinterpret_cast<uint64_t*>(tmem_slot + 2);           // [2] GATHER: q/k/v of a stage have landed
  uint4* padrow = reinterpret_cast<uint4*>(bufs + 2 * AT_BUF_BYTES + 2048); // [2 item slots][3][4]: bf16 qkv bias of the unit's head

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int C = p.C, heads = p.heads, C3 = 3 * p.C;
  const WinGeom g = p.g;
  const int64_t HW = (int64_t)g.H * g.W;
  const int BP = (p.B + 1) / 2;                            // steps per item

  for (int i = tid; i < 2 * AT_BUF_BYTES / 16; i += AT_THREADS)
    reinterpret_cast<uint4*>(bufs)[i] = make_uint4(0, 0, 0, 0);          // padding rows must stay finite
  if (tid < 64) reinterpret_cast<uint4*>(ones)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    if (GATHER) tma_prefetch_desc(&map_qkv);
    mbar_fence_init();
  }
  if (GATHER) fence_async_shared();                        // zero-filled stages / ones tile vs. the async proxy
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(AT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  auto bias_chunk = [&](int ch) {
    uint4 r = make_uint4(0, 0, 0, 0);
    if (p.qkv_bias) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p.qkv_bias + ch));
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.qkv_bias + ch) + 1);
      r = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    }
    return r;
  };

  const int n_units = p.n_items;                           // (window positions) x chunks x heads
  const int CH = p.hc;                                     // image pairs per unit
  const int NCH = (BP + CH - 1) / CH;

  const int unit = tid >> 6;                               // warp-uniform: which image of the pair
  const int ti = tid & 63;
  const int ic = ti < N ? ti : 0;
  const int ri = ic / WS, ci = ic - ri * WS;
  const int lc = tid & 3;
  const int lt0 = tid >> 2;

  auto decode = [&](BiStep& st) {                          // unit index -> item and image-pair range
    const int e = st.u % heads;
    const int wc = st.u / heads;
    const int chunk = wc % NCH;
    st.item = (wc / NCH) * heads + e;
    st.bp = chunk * CH;
    st.bp_end = st.bp + CH < BP ? st.bp + CH : BP;
  };
  auto first_step = [&](int u) {
    BiStep st;
    st.u = u; st.n = 0; st.ic = 0;
    decode(st);
    return st;
  };
  auto next_step = [&](BiStep st) {
    ++st.n;
    if (++st.bp == st.bp_end) {
      st.u += gridDim.x;
      ++st.ic;
      if (st.u < n_units) decode(st);
    }
    return st;
  };
  // token map of the item's window position -> src[ic & 1] (threads 0..63)
  auto prep_item = [&](const BiStep& st) {
    if (tid < 64) {
      int t = -1;
      if (tid < N) {
        const int wi = st.item / heads;
        const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
        t = source_token(g, wr * WS + ri, wc * WS + ci);
      }
      src[(st.ic & 1) * 64 + tid] = t;
    } else if (tid < 76) {                                 // the unit's head slice of the bf16 qkv bias: [q|k|v][4 x 16 B]
      const int j = tid - 64;
      padrow[(st.ic & 1) * 12 + j] = bias_chunk((j >> 2) * C + (st.item % heads) * 32 + (j & 3) * 8);
    }
  };
  int ld_t[2], ld_dst[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = lt0 + 32 * (k & 1);
    if (k < 2) ld_t[k] = t < N ? t : -1;
    const int row = (k >> 1) * 64 + t;
    ld_dst[k] = row * 64 + ((lc ^ ((row >> 1) & 3)) << 4);
  }
  // GATHER loader role: thread j < 78 fetches token slots 4*grp .. 4*grp+3 of image `gu` for part `gpart` (q / k / v)
  const int gpart = tid / 26, grem = tid - gpart * 26;
  const int gu = grem / 13, ggrp = grem - gu * 13;
  const int bias_row = p.B * (int)HW;                      // the extra row of the qkv tensor: bf16 qkv bias
  auto issue_loads = [&](const BiStep& st) {
    if constexpr (GATHER) {
      uint64_t* fb = &full[st.n & 1];
      if (tid == 0) mbar_expect_tx(fb, 78 * 256);
      if (tid < 78) {
        const int4 t4 = *reinterpret_cast<const int4*>(src + (st.ic & 1) * 64 + 4 * ggrp);   // slots >= 49 hold -1
        const int b = 2 * st.bp + gu;
        const int base_row = b * (int)HW;
        const bool bv = b < p.B;
        const int r0 = (bv && t4.x >= 0) ? base_row + t4.x : bias_row;
        const int r1 = (bv && t4.y >= 0) ? base_row + t4.y : bias_row;
        const int r2 = (bv && t4.z >= 0) ? base_row + t4.z : bias_row;
        const int r3 = (bv && t4.w >= 0) ? base_row + t4.w : bias_row;
        uint8_t* dst = bufs + (st.n & 1) * AT_BUF_BYTES + gpart * AT_PART_BYTES + (gu * 64 + 4 * ggrp) * 64;
        tma_gather4(dst, &map_qkv, fb, gpart * C + (st.item % heads) * 32, r0, r1, r2, r3);
      }
      return;
    }
    uint8_t* base = bufs + (st.n & 1) * AT_BUF_BYTES;
    const int* smap = src + (st.ic & 1) * 64;
    const int e = st.item % heads;
    const bf16* gq = p.qkv + e * 32 + lc * 8;
    const int bch = e * 32 + lc * 8;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int b = 2 * st.bp + (k >> 1);
      if (ld_t[k & 1] >= 0 && b < p.B) {
        const int t = smap[ld_t[k & 1]];
        uint8_t* dst = base + ld_dst[k];
        if (t >= 0) {
          const bf16* grow = gq + ((int64_t)b * HW + t) * C3;
          cp_async16(dst, grow);
          cp_async16(dst + AT_PART_BYTES, grow + C);
          cp_async16(dst + 2 * AT_PART_BYTES, grow + 2 * C);
        } else if (p.mode == 4) {                          // diagnostics: padding cells straight from the fp32 bias
          *reinterpret_cast<uint4*>(dst) = bias_chunk(bch);
          *reinterpret_cast<uint4*>(dst + AT_PART_BYTES) = bias_chunk(bch + C);
          *reinterpret_cast<uint4*>(dst + 2 * AT_PART_BYTES) = bias_chunk(bch + 2 * C);
        } else {                                           // padding cell: q/k/v = bias (staged per unit by prep_item)
          const uint4* pr = padrow + (st.ic & 1) * 12 + lc;
          *reinterpret_cast<uint4*>(dst) = pr[0];
          *reinterpret_cast<uint4*>(dst + AT_PART_BYTES) = pr[4];
          *reinterpret_cast<uint4*>(dst + 2 * AT_PART_BYTES) = pr[8];
        }
      }
    }
  };
  // my row of the item's precomputed bias: [window][head][chunk][row] float4
  float bias[4 * AT_FULL_CHUNKS];
  auto load_bias = [&](const BiStep& st) {
    if (ti < N && p.mode != 2) {
      const float4* brow = p.bias_full + (size_t)st.item * (AT_FULL_CHUNKS * 64) + ti;
#pragma unroll
      for (int k = 0; k < AT_FULL_CHUNKS; ++k) {
        const float4 b4 = __ldg(brow + k * 64);
        bias[4 * k] = b4.x; bias[4 * k + 1] = b4.y; bias[4 * k + 2] = b4.z; bias[4 * k + 3] = b4.w;
      }
    }
  };
#pragma unroll
  for (int k = 0; k < 4 * AT_FULL_CHUNKS; ++k) bias[k] = 0.f;

  BiStep cur = first_step(blockIdx.x);
  if (cur.u < n_units) {
    prep_item(cur);
    const BiStep n1 = next_step(cur);
    if (n1.u < n_units && n1.ic != cur.ic) prep_item(n1);
    __syncthreads();
    issue_loads(cur);
    load_bias(cur);
  }
  cp_async_commit();

  uint32_t par = 0;
  long long ph[6] = {0, 0, 0, 0, 0, 0};
  const bool prof = p.dbg != nullptr && blockIdx.x == 0 && tid == 0;
  while (cur.u < n_units) {
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (prof) c0 = clock64();
    const BiStep nxt = next_step(cur);
    const bool has_next = nxt.u < n_units;
    const int my_b = 2 * cur.bp + unit;
    const bool row_valid = (ti < N) && (my_b < p.B);
    const int e = cur.item % heads;
    // ---- 1. this step's q/k/v (requested one step ago) have landed
    if constexpr (!GATHER) {
      cp_async_wait<0>();
      fence_async_shared();
    }
    __syncthreads();
    if (prof) c1 = clock64();
    // ---- 2. S = Q . K^T (both images at once, block diagonal)
    const uint32_t sq = smem_u32(bufs + (cur.n & 1) * AT_BUF_BYTES);
    if (tid == 0) {
      if constexpr (GATHER) mbar_wait(&full[cur.n & 1], (uint32_t)(cur.n >> 1) & 1);
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 128, 0, 0);
      const uint64_t dq = umma_smem_desc(sq, 16, 512, UMMA_SWIZZLE_64B);
      const uint64_t dk = umma_smem_desc(sq + AT_PART_BYTES, 16, 512, UMMA_SWIZZLE_64B);
      umma_ss(tmem_base, dq, dk, idesc, 0);
      umma_ss(tmem_base, dq + 2, dk + 2, idesc, 1);
      umma_commit(&bars[0]);
    }
    if (has_next && p.mode != 3) issue_loads(nxt);         // whole next step into the other stage (mode 3: diagnostics, no loads)
    cp_async_commit();
    mbar_wait(&bars[0], par);
    tc_fence_after();
    if (prof) c2 = clock64();
    // ---- 3. bias + softmax on my row
    {
      uint32_t sr[N];
      const uint32_t s_addr = tmem_base + lane_base + (uint32_t)(unit * 64);
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
      float t[N];
#pragma unroll
      for (int j = 0; j < N; ++j) t[j] = fmaf(__uint_as_float(sr[j]), p.scale, bias[j]);
      float mx = t[N - 1];
#pragma unroll
      for (int j = 0; j + 1 < N; j += 2) mx = fmax3(mx, t[j], t[j + 1]);
      const float mneg = -mx * LOG2E;
      uint32_t pk[32];                                     // the row sum comes from the tensor core (P . ones)
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        float p0 = 0.f, p1 = 0.f;
        if (2 * k < N) p0 = fast_exp2(fmaf(t[2 * k], LOG2E, mneg));
        if (2 * k + 1 < N) p1 = fast_exp2(fmaf(t[2 * k + 1], LOG2E, mneg));
        pk[k] = row_valid ? pack_bf16x2(p0, p1) : 0u;
      }
      tmem_st_x32(tmem_base + lane_base + AT_P_COL, pk);
      tmem_st_wait();
    }
    if (has_next && nxt.ic != cur.ic) load_bias(nxt);      // new (window, head): the bias registers are free again
    tc_fence_before();
    __syncthreads();
    if (prof) c3 = clock64();
    // ---- 4. O = P . [V_img0 | V_img1]
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, 64, 0, 1);
      const uint32_t idesc1 = umma_idesc_bf16(128, 16, 0, 1);    // row sums: P . ones[keys][16]
      const uint32_t sv = sq + 2 * AT_PART_BYTES;
      const uint64_t d1 = umma_smem_desc(smem_u32(ones), 4096, 512, UMMA_SWIZZLE_64B);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t dv = umma_smem_desc(sv + k * 1024, 4096, 512, UMMA_SWIZZLE_64B);
        umma_ts(tmem_base + AT_O_COL, tmem_base + AT_P_COL + k * 8, dv, idesc, k > 0);
        umma_ts(tmem_base + AT_SUM_COL, tmem_base + AT_P_COL + k * 8, d1, idesc1, k > 0);
      }
      umma_commit(&bars[1]);
    }
    // while the MMA runs: token map of the item after next's first step (its loader runs after the next barriers)
    if (has_next) {
      const BiStep nn = next_step(nxt);
      if (nn.u < n_units && nn.ic != nxt.ic) prep_item(nn);
    }
    mbar_wait(&bars[1], par);
    tc_fence_after();
    if (prof) c4 = clock64();
    // ---- 5. normalise and store my output row at the token's un-shifted position
    {
      uint32_t orow[32], osum;
      tmem_ld_x32(tmem_base + lane_base + AT_O_COL + (uint32_t)(unit * 32), orow);
      tmem_ld_x1(tmem_base + lane_base + AT_SUM_COL, osum);
      tmem_ld_wait();
      const int ts = src[(cur.ic & 1) * 64 + ic];
      if (row_valid && ts >= 0) {
        const float inv = 1.0f / __uint_as_float(osum);
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
    tc_fence_before();
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(AT_TMEM_COLS) : "memory");
  }
}

static size_t attn_tc_smem_bytes(int ws, int C) {
  const int TW = 2 * ws - 1;
  size_t b = 1024;                                   // alignment slack
  b += 2 * AT_BUF_BYTES;
  b += (size_t)2 * AT_TAB_WORDS * 4;
  (void)TW;
  b += 3 * 2 * 64 * 4;
  b += 2 * 8 + 16;
  (void)C;
  // keep the CTA count per SM at AT_CTAS_PER_SM (register budget 65536 / (3 * 128) = 170 per thread)
  const size_t floor_bytes = (size_t)(233472 / (AT_CTAS_PER_SM + 1)) - 1024 + 16;
  return b < floor_bytes ? floor_bytes : b;
}

// (alpha, beta)[idx][head] fp32 -> tables[head][AT_TAB_WORDS] half2 (alpha, beta): entry (r, c) of the (2w-1)^2
// table sits at word r * AT_TAB_PITCH + c — exactly the shared-memory image the attention kernel copies per head
__global__ void bias_tables_kernel(const float* __restrict__ alpha, const float* __restrict__ beta,
                                   __half2* __restrict__ out, int heads, int tw) {
  const int n = heads * AT_TAB_WORDS;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int e = i / AT_TAB_WORDS;
    const int o = i - e * AT_TAB_WORDS;
    const int r = o / AT_TAB_PITCH, c = o - r * AT_TAB_PITCH;
    float a = 0.f, b = 0.f;
    if (r < tw && c < tw) { a = alpha[(r * tw + c) * heads + e]; b = beta[(r * tw + c) * heads + e]; }
    out[i] = __floats2half2_rn(a, b);
  }
}

// ---------------------------------------------------------------------------------------------------
// Full additive bias of every (window of one image, head): bias[i][j] = hav(uv_i, uv_j) * alpha[idx(i,j)][head] +
// beta[idx(i,j)][head] (+ mask[window][i][j] in planar mode), fp32, laid out for the kernel's row-per-lane reads:
// [window][head][chunk k = j / 4][row i (64)][j % 4].  Depends on the geometry and on the block's alpha / beta, not
// on the batch: built once per block and resolution, read (L2-resident) by every image.  Same formulas and fp32
// evaluation order as the parity kernel (lzx/models/great_circle.py:82-86, reference :241-272).
// ---------------------------------------------------------------------------------------------------
__global__ void bias_full_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ uv,
                                 const float* __restrict__ mask, float* __restrict__ table, WinGeom g, int heads) {
  const int ws = g.ws, N = ws * ws, tw = 2 * ws - 1;
  __shared__ float su[64], sv[64], scv[64];
  const int wi = blockIdx.x, e = blockIdx.y;
  const int wr = wi / g.nWw, wc = wi - wr * g.nWw;
  for (int t = threadIdx.x; t < N; t += blockDim.x) {
    const int r = t / ws, c = t - r * ws;
    const int s = source_token(g, wr * ws + r, wc * ws + c);
    float uu = 0.f, vv = 0.f;
    if (s >= 0 && uv != nullptr) { uu = uv[2 * s]; vv = uv[2 * s + 1]; }
    su[t] = uu; sv[t] = vv; scv[t] = cosf(vv);
  }
  __syncthreads();
  float* out = table + ((size_t)wi * heads + e) * (AT_FULL_CHUNKS * 64 * 4);
  for (int q = threadIdx.x; q < AT_FULL_CHUNKS * 64 * 4; q += blockDim.x) {
    const int k = q >> 8, i = (q >> 2) & 63, j = 4 * k + (q & 3);
    float b = 0.f;
    if (i < N && j < N) {
      float d = 0.f;
      if (uv != nullptr) {
        const float sdv = sinf(0.5f * fabsf(sv[j] - sv[i]));
        const float sdu = sinf(0.5f * (su[j] - su[i]));
        const float a = sdv * sdv + (scv[j] * scv[i]) * (sdu * sdu);
        d = asinf(sqrtf(fminf(a, 1.0f))) * 2.0f;
      }
      const int ri = i / ws, ci = i - ri * ws, rj = j / ws, cj = j - rj * ws;
      const int idx = (ri - rj + ws - 1) * tw + (ci - cj + ws - 1);
      b = fmaf(d, alpha[idx * heads + e], beta[idx * heads + e]);
      if (mask != nullptr) b += mask[((size_t)wi * N + i) * N + j];
    }
    out[q] = b;
  }
}

int window_bias_full(const float* alpha, const float* beta, const float* uv, const float* mask, void* table, int H, int W,
                     int heads, int window, int shift, int pano, cudaStream_t st) {
  PSW_REQUIRE(window * window <= 4 * AT_FULL_CHUNKS && window * window <= 64, PSW_ERR_UNSUPPORTED,
              "psw_window_bias_full: window %d too large", window);
  WinGeom g = make_geom(H, W, window, shift, pano);
  bias_full_kernel<<<dim3(g.nWh * g.nWw, heads), 256, 0, st>>>(alpha, beta, pano ? uv : nullptr, mask, (float*)table, g, heads);
  return launch_status("bias_full_kernel");
}

int window_bias_tables(const float* alpha, const float* beta, void* tables, int heads, int window, cudaStream_t st) {
  const int tw = 2 * window - 1;
  PSW_REQUIRE(tw * AT_TAB_PITCH <= AT_TAB_WORDS, PSW_ERR_UNSUPPORTED, "psw_window_bias_tables: window %d too large", window);
  const int n = heads * AT_TAB_WORDS;
  bias_tables_kernel<<<(n + 255) / 256, 256, 0, st>>>(alpha, beta, (__half2*)tables, heads, tw);
  return launch_status("bias_tables_kernel");
}

int window_attn_tc(const bf16* qkv, bf16* out, const float* alpha, const float* beta, const void* tables,
                   const float* qkv_bias, const void* hav_table, const float* mask, const void* bias_full,
                   bool bias_row_present, int B, int H, int W, int C, int heads, int window, int shift, int pano, float scale,
                   long long* dbg, int mode, cudaStream_t st) {
  PSW_REQUIRE(window == 7, PSW_ERR_UNSUPPORTED,
              "psw_window_attn_fwd(bf16): the tcgen05 kernel is instantiated for window 7 (every shipped PanoSwin config); got %d",
              window);
  PSW_REQUIRE(bias_full != nullptr || tables != nullptr, PSW_ERR_BAD_ARG,
              "psw_window_attn_fwd(bf16): needs the per-head bias tables (psw_window_bias_tables)");
  PSW_REQUIRE(bias_full != nullptr || !pano || hav_table, PSW_ERR_BAD_ARG,
              "psw_window_attn_fwd(bf16): pano mode needs the great-circle table (psw_window_hav_table)");
  AttnParams p;
  p.qkv = qkv; p.out = out; p.alpha = alpha; p.beta = beta; p.qkv_bias = qkv_bias;
  p.tables = (const __half2*)tables;
  p.bias_full = (const float4*)bias_full;
  p.hav = pano ? (const __half*)hav_table : nullptr;
  p.mask = mask;
  p.g = make_geom(H, W, window, shift, pano);
  p.B = B; p.C = C; p.heads = heads; p.scale = scale; p.dbg = dbg; p.mode = mode;
  p.n_windows = B * p.g.nWh * p.g.nWw;
  // heads per work item: one.  Sharing a window pair's token maps / distance rows between the heads of a wider
  // item saves little (measured: equal at stage 0), while single-head items balance the small launches of the late
  // stages better (items are dealt out in contiguous ranges: a CTA runs ceil(items / CTAs) of them; -10% at stage 3).
  p.hc = 1;
  if (attn_debug_hc() > 0 && heads % attn_debug_hc() == 0) p.hc = attn_debug_hc();
  p.n_items = ((p.n_windows + 1) / 2) * (heads / p.hc);
  const size_t smem = attn_tc_smem_bytes(window, C);
  PSW_REQUIRE(smem <= 220 * 1024, PSW_ERR_UNSUPPORTED, "psw_window_attn_fwd(bf16): C=%d too large", C);
  PSW_REQUIRE((int64_t)B * H * W < (1ll << 31) / 1, PSW_ERR_UNSUPPORTED, "psw_window_attn_fwd(bf16): too many tokens");
  const bool batch_inner = bias_full != nullptr && B >= 4 && attn_debug_hc() != 15 && mode != 1;
  if (batch_inner) {
    // image pairs per unit: as many as possible (the bias row and the token map are fetched once per unit) while the
    // units still spread evenly over the resident CTAs (CTA c runs units c, c + grid, ...)
    const int bp_total = (B + 1) / 2, ctas = num_sms() * AT_CTAS_PER_SM;
    const int wh = p.g.nWh * p.g.nWw * heads;
    double best = -1.0;
    p.hc = 1;
    for (int ch = 4; ch >= 1; ch >>= 1) {                 // measured: 8 is not better than 4 where both balance
      const int64_t units = (int64_t)wh * ((bp_total + ch - 1) / ch);
      const double steps_max = (double)((units + ctas - 1) / ctas) * ch;             // steps of the busiest CTA (upper bound)
      const double eff = (double)wh * bp_total / ctas / steps_max;
      const int forced = attn_debug_hc();
      if (forced ? ch == forced : (best < 0.9 && eff > best + 1e-9)) { best = eff; p.hc = ch; }
    }
    p.n_items = wh * ((bp_total + p.hc - 1) / p.hc);                                // units: (window, chunk, head)
  }
  if (batch_inner) {
    const bool gather = bias_row_present && attn_debug_hc() != 14;
    CUtensorMap map_qkv;
    if (gather) {
      const uint64_t dims[2] = {(uint64_t)(3 * C), (uint64_t)((int64_t)B * H * W + 1)};
      const uint64_t strides[1] = {(uint64_t)(3 * C) * 2};
      const uint32_t box[2] = {32, 1};
      int rc = make_tensor_map_nd(&map_qkv, qkv, 2, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    } else {
      memset(&map_qkv, 0, sizeof(map_qkv));
    }
    auto kern = gather ? window_attn_bi_kernel<7, true> : window_attn_bi_kernel<7, false>;
    PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int grid = num_sms() * AT_CTAS_PER_SM;
    if (grid > p.n_items) grid = p.n_items;
    kern<<<grid, AT_THREADS, smem, st>>>(p, map_qkv);
    return launch_status("window_attn_bi_kernel");
  }
  auto kern = bias_full ? window_attn_tc_kernel<7, false, true>
                        : (mask ? window_attn_tc_kernel<7, true, false> : window_attn_tc_kernel<7, false, false>);
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int grid = num_sms() * AT_CTAS_PER_SM;
  if (grid > p.n_items) grid = p.n_items;
  kern<<<grid, AT_THREADS, smem, st>>>(p);
  return launch_status("window_attn_tc_kernel");
}

static int g_attn_hc = 0;
int attn_debug_hc() { return g_attn_hc; }
void attn_debug_set_hc(int hc) { g_attn_hc = hc; }

int window_hav_table(const float* uv, void* table, int H, int W, int window, int shift, cudaStream_t st) {
  WinGeom g = make_geom(H, W, window, shift, 1);
  hav_table_kernel<<<g.nWh * g.nWw, 128, 0, st>>>(uv, (__half*)table, g);
  return launch_status("hav_table_kernel");
}

}  // namespace psw
