// K2 (throughput path): y = act(x . w^T + bias) (+ residual) with bf16 operands on the 5th-generation
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged by TMA.
//
// Persistent, warp-specialised CTA (one per SM, 320 or 576 threads):
//   warp 0      TMA producer : [128 x 64] x-tile and [BLOCK_N x 64] w-tile per stage, SWIZZLE_128B, mbarrier ring
//   warp 1      MMA issuer   : one thread issues 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per stage; commits free
//                              the smem slot and, per tile, publish the accumulator
//   warps 2..   epilogue     : two groups of 4 or 8 warps.  The accumulator is double-buffered in TMEM (2 x 256
//                              columns); group g drains buffer g, i.e. every second tile of the CTA, so one group's
//                              math overlaps the other group's stores and the MMAs of the tile after.  Per 64-byte
//                              wide chunk: tcgen05.ld 32 rows x CW columns -> +bias (shared-memory copy) -> [GELU]
//                              -> [+residual, TMA-loaded two chunks ahead into a per-warp ring] -> transpose through
//                              a swizzled smem tile -> 16-byte global stores in which four lanes cover one 64-byte
//                              row segment (whole sectors; measured faster than TMA stores of such small boxes:
//                              no proxy fence, no ~300-cycle UTMASTG issue on the critical path).
// CTA-pair mode (CG = 2, used for deep-K shapes): a cluster of two CTAs owns a 256-row tile; each CTA stages its
// 128 x-rows and HALF of the w-tile rows, the leader issues tcgen05.mma.cta_group::2 (M=256) which reads both CTAs'
// shared memory and writes both CTAs' TMEM, tcgen05.commit multicasts the barrier arrivals to both CTAs, and each
// CTA runs the epilogue of its own 128 rows.  Operand traffic L2 -> SM per output column drops by ~1/3.
// Tiles are ordered n-fastest, so CTAs of one wave share the x-tile through L2.
// Rows beyond M and K beyond the tensor are zero-filled by TMA; stores are predicated at M and N.
//
// GELU on this path: 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with a minimax fit of the exact erf GELU
// (max abs error 5.1e-5 over the real line, below bf16 resolution of the activations), evaluated two at a time in
// packed fp16 when the output is bf16 — the fp32 parity path keeps erff.
#include "psw_common.cuh"

namespace psw {

constexpr int TC_BM = 128;         // UMMA M
constexpr int TC_BK = 64;          // one 128-byte swizzle row of bf16
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_EPI_WARPS = 16;
constexpr int TC_TILE_BYTES = 32 * 64;      // epilogue staging tile: 32 rows x 64 B, SWIZZLE_64B

// Patch-convolution view of the A operand (psw_patch_conv_fwd): token (ty, tx) of a non-overlapping ph x pw patch grid
// is GEMM row ty * wt + tx, its K axis is (dy, dx, c): for a fixed dy the pw * cin values are contiguous in the NHWC
// image, so x-tiles are boxes of a 3-D tensor {pw * cin, wt, B * H}.  tiles_x = 0 means a plain 2-D A operand.
//
// conv3 = 1 (psw_conv3x3_nhwc_fwd): 3 x 3 / stride 1 / pad 1 convolution of an NHWC image without im2col.  Output pixel
// (b, y, x) is GEMM row (b * himg + y) * wt + x and its K axis is (tap = dy * 3 + dx, c): the x-tile of K-block
// (tap, kc) is the box {64 channels, 128 pixels, 1 row, 1 image} of the 4-D tensor {cin, W, H, B} at
// (64 kc, x0 + dx - 1, y + dy - 1, b) -- the same pixels shifted by one; coordinates outside the image are zero-filled
// by TMA, which is the convolution's zero padding.  kb_per_dy = K-blocks per tap (cin / 64).
struct ConvView {
  int tiles_x;            // M-tiles per token row
  int wt;                 // tokens per image row
  int kb_per_dy;          // K-blocks (64 elements) per patch row: pw * cin / 64
  int ph;                 // patch height
  int conv3;              // 3 x 3 stride-1 view (see above)
  int himg;               // conv3: rows per image
  int relu;               // ReLU after the bias (any view)
};

struct alignas(16) TcSmemTail {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint64_t res_bar[TC_MAX_EPI_WARPS][2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of a cluster share one 256-row tile; the leader (cluster rank 0)
// issues the MMAs, which read A / B from both CTAs' shared memory and write both CTAs' tensor memory.
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> rank 0
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2_rt(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// TMA load into this CTA's shared memory whose bytes are counted on the LEADER's mbarrier (same offset)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1) : "memory");
}
// 3-D variants for the patch-convolution view of the A operand (coordinates: k inside the patch row, token x, image row)
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// arrive (when all prior MMAs retire) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader's barrier at this offset (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);           // the fitted polynomial is used on |x| <= 8; beyond, tanh saturates
  const float p = fmaf(fmaf(-3.20974528e-04f, x2, 3.68320430e-02f), x2, 7.97686932e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Two GELUs in packed fp16 arithmetic (the result is rounded to bf16 anyway): same fit, 6 instead of 9.5
// instructions per element.  satfinite keeps |x| > 65504 finite so the tail evaluates to x or 0, not NaN.
__device__ __forceinline__ uint32_t gelu_pair_bf16(float x0, float x1) {
  uint32_t h, t;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
  const __half2 x = *reinterpret_cast<const __half2*>(&h);
  const __half2 x2 = __hmin2(__hmul2(x, x), __float2half2_rn(64.0f));
  const __half2 p = __hfma2(__hfma2(__float2half2_rn(-3.20974528e-04f), x2, __float2half2_rn(3.68320430e-02f)), x2,
                            __float2half2_rn(7.97686932e-01f));
  const __half2 u = __hmul2(x, p);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(*reinterpret_cast<const uint32_t*>(&u)));
  const __half2 hx = __hmul2(x, __float2half2_rn(0.5f));
  const float2 f = __half22float2(__hfma2(hx, *reinterpret_cast<const __half2*>(&t), hx));
  return pack_bf16x2(f.x, f.y);
}

// diagnostics (mode bit 4): SM-cycle totals of CTA 0 -- {producer wait-empty, mma wait-tempty, mma wait-full, mma issue,
// epi wait-tfull, epi tmem-ld, epi math+stage, epi store-issue, tiles of CTA 0}
__device__ long long g_tc_cycles[16];

template <typename TO> struct Chunk;             // CW output columns = one 64-byte row of the staging tile
template <> struct Chunk<float> { static constexpr int CW = 16; };
template <> struct Chunk<bf16> { static constexpr int CW = 32; };

// Two groups of GW epilogue warps (GW = 4 or 8: one or two warps per TMEM lane quadrant); group g drains accumulator
// buffer g, i.e. every second tile of the CTA, so the math of one group overlaps the stores of the other.
template <int GW> struct EpiCfg { static constexpr int EW = 2 * GW; static constexpr int THREADS = 64 + 32 * EW; };

// LayerNorm of the freshly written rows fused into the epilogue (LNF; proj -> norm2 and fc2 -> next block's norm1 when
// one tile holds complete rows, N <= 256): the thread that owns TMEM lane r makes three passes over its row, which
// stays in tensor memory -- (1) x = acc + bias + residual, written out in fp32, summed, and written BACK to TMEM,
// (2) centred sum of squares, (3) (x - mean) * rstd * gamma + beta -> bf16, written to ln_out.  Two-pass statistics in
// fp32 like the standalone kernel; saves the LayerNorm kernel's re-read of the fp32 residual stream.
struct LnFuse {
  const float* gamma;
  const float* beta;
  bf16* out;              // LNF 1: [M, N] bf16
  float eps;
  const float* pos;       // LNF 2: position rows added after the LayerNorm ([pos_rows, N] fp32, row = GEMM row % pos_rows)
  int pos_rows;
  float* out_nchw;        // LNF 3: LN(y) written transposed, fp32 [M / hw, N, hw] (the stage's output map)
  int hw;                 // LNF 3: tokens per image
};
// LNF 1: y = acc + bias + residual (fp32, written) and ln.out = LN(y) (bf16).
// LNF 2: y = LN(acc + bias) * gamma + beta + pos (fp32) only -- the stem: patch conv -> patch_norm -> + abs. position.
// LNF 3: like LNF 1, but LN(y) goes out as the stage's fp32 NCHW feature map: lane r owns token r of the tile, so a
//        warp-wide store of one channel covers 32 consecutive tokens = 128 contiguous bytes -- no transposition needed.

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <bool GELU, bool RES, typename TO, int GW, int CG, int LNF = 0>
__global__ void __launch_bounds__(EpiCfg<GW>::THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_r,
                 const float* __restrict__ bias, TO* __restrict__ y, int64_t M, int N, int K, int block_n, int stages, int mode,
                 const ConvView cv, const LnFuse ln) {
  constexpr int CW = Chunk<TO>::CW;
  constexpr int TC_EPI_WARPS = EpiCfg<GW>::EW;
  constexpr int PER_QUAD = GW / 4;                            // warps of one group sharing a TMEM lane quadrant
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = TC_BM * TC_BK * 2;                 // 16 KiB
  const uint32_t b_bytes = (uint32_t)(block_n / CG) * TC_BK * 2;   // pair mode: each CTA stages half of the w-tile rows
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* epi_smem = smem + (size_t)stages * stage_bytes;    // per warp: 1 out tile (+ 2 residual tiles)
  constexpr int EPI_PER_WARP = (RES ? 3 : 1) * TC_TILE_BYTES;
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(epi_smem + TC_EPI_WARPS * EPI_PER_WARP);
  float* bias_s = reinterpret_cast<float*>(tail + 1);         // bias of all n_tiles * block_n columns, zero padded

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + block_n - 1) / block_n;
  constexpr int TILE_M = TC_BM * CG;                          // rows per tile (of the CTA or of the CTA pair)
  const int64_t m_tiles = cv.tiles_x ? (M / cv.wt) * cv.tiles_x : (M + TILE_M - 1) / TILE_M;
  const int64_t total_tiles = m_tiles * n_tiles;
  const int cta_rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG;                           // index of this CTA (pair) among the tile workers
  const int n_units = gridDim.x / CG;
  const int k_blocks = (K + TC_BK - 1) / TC_BK;
  const uint32_t acc_stride = 256;

  for (int i = threadIdx.x; i < n_tiles * block_n; i += blockDim.x) bias_s[i] = (bias && i < N) ? bias[i] : 0.0f;
  float* gam_s = bias_s + n_tiles * block_n;                  // LNF: gamma, beta of the fused LayerNorm
  float* bet_s = gam_s + N;
  if constexpr (LNF != 0) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) { gam_s[i] = ln.gamma[i]; bet_s[i] = ln.beta[i]; }
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    if (RES) tma_prefetch_desc(&map_r);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->tfull[s], 1);
      mbar_init(&tail->tempty[s], GW * CG);
    }
    for (int w = 0; w < TC_EPI_WARPS; ++w) {
      mbar_init(&tail->res_bar[w][0], 1);
      mbar_init(&tail->res_bar[w][1], 1);
    }
    mbar_fence_init();
  }
  if (CG == 2) cluster_sync_all();                            // peer barriers are initialised before anything remote
  if (warp == 1) {
    if (CG == 2) tmem_alloc2_rt(&tail->tmem_base, 512);
    else tmem_alloc_rt(&tail->tmem_base, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const bool prof = (mode & 16) && blockIdx.x == 0;
      long long c_we = 0;
      for (int64_t tile = unit; tile < total_tiles; tile += n_units) {
        const int m_t = (int)((uint32_t)tile / (uint32_t)n_tiles);
        const int n_t = (int)tile - m_t * n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(&tail->empty[stage], phase ^ 1);
          if (prof) c_we += clock64() - t0;
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if (CG == 2) {                                      // bytes of both CTAs are counted on the leader's barrier
            if (cta_rank == 0) mbar_expect_tx(&tail->full[stage], 2 * stage_bytes);
            if (cv.tiles_x) {
              const int trow = m_t / cv.tiles_x, x0 = (m_t - trow * cv.tiles_x) * TILE_M + cta_rank * TC_BM;
              const int dy = kb / cv.kb_per_dy;
              tma_load_3d_pair(sa, &map_x, &tail->full[stage], (kb - dy * cv.kb_per_dy) * TC_BK, x0, trow * cv.ph + dy);
            } else {
              tma_load_2d_pair(sa, &map_x, &tail->full[stage], kb * TC_BK, m_t * TILE_M + cta_rank * TC_BM);
            }
            tma_load_2d_pair(sa + a_bytes, &map_w, &tail->full[stage], kb * TC_BK, n_t * block_n + cta_rank * (block_n / 2));
          } else if (mode & 2) {                              // diagnostics: no loads
            mbar_arrive(&tail->full[stage]);
          } else {
            mbar_expect_tx(&tail->full[stage], stage_bytes);
            if (cv.conv3) {
              const int trow = m_t / cv.tiles_x, x0 = (m_t - trow * cv.tiles_x) * TILE_M;
              const int tap = kb / cv.kb_per_dy, dy = tap / 3, dx = tap - 3 * dy;
              const int b = trow / cv.himg, yy = trow - b * cv.himg;
              tma_load_4d(sa, &map_x, &tail->full[stage], (kb - tap * cv.kb_per_dy) * TC_BK, x0 + dx - 1, yy + dy - 1, b);
            } else if (cv.tiles_x) {
              const int trow = m_t / cv.tiles_x, x0 = (m_t - trow * cv.tiles_x) * TILE_M;
              const int dy = kb / cv.kb_per_dy;
              tma_load_3d(sa, &map_x, &tail->full[stage], (kb - dy * cv.kb_per_dy) * TC_BK, x0, trow * cv.ph + dy);
            } else {
              tma_load_2d(sa, &map_x, &tail->full[stage], kb * TC_BK, m_t * TILE_M);
            }
            tma_load_2d(sa + a_bytes, &map_w, &tail->full[stage], kb * TC_BK, n_t * block_n);
          }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
      if (prof) g_tc_cycles[0] = c_we;
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(TILE_M, block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const bool prof = (mode & 16) && blockIdx.x == 0;
      long long c_wt = 0, c_wf = 0, c_is = 0, n_t = 0;
      for (int64_t tile = unit; tile < total_tiles; tile += n_units) {
        long long t0 = prof ? clock64() : 0;
        mbar_wait(&tail->tempty[acc], acc_phase ^ 1);         // epilogue has drained this accumulator
        tc_fence_after();
        if (prof) { const long long t1 = clock64(); c_wt += t1 - t0; ++n_t; }
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          t0 = prof ? clock64() : 0;
          mbar_wait(&tail->full[stage], phase);               // TMA bytes have landed
          tc_fence_after();
          const long long t1 = prof ? clock64() : 0;
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = umma_smem_desc(sa, 16, 1024, UMMA_SWIZZLE_128B);
          const uint64_t db = umma_smem_desc(sa + a_bytes, 16, 1024, UMMA_SWIZZLE_128B);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {              // advance 32 B inside the swizzle row: +2 (>>4)
            if (CG == 2) umma_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else if (!(mode & 4)) umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if (CG == 2) umma_commit_pair(&tail->empty[stage]);  // frees the smem slot (both CTAs) when the MMAs retire
          else umma_commit(&tail->empty[stage]);
          if (prof) { c_wf += t1 - t0; c_is += clock64() - t1; }
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        if (CG == 2) umma_commit_pair(&tail->tfull[acc]);     // accumulator complete -> epilogue (of both CTAs)
        else umma_commit(&tail->tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (prof) { g_tc_cycles[1] = c_wt; g_tc_cycles[2] = c_wf; g_tc_cycles[3] = c_is; g_tc_cycles[8] = n_t; }
    }
  } else {
    // ------------------------------- epilogue ------------------------------------
    if constexpr (LNF != 0) {
      // one warp per TMEM lane quadrant and group (GW = 4): lane r owns row r of the tile for all three passes
      static_assert(LNF == 0 || (sizeof(TO) == 4 && GW == 4 && CG == 1 && !GELU && RES == (LNF != 2)), "LNF: fp32 epilogue, GW = 4");
      const int ew = warp - 2;
      const int quad = warp & 3;
      const int grp = ew / GW;
      uint8_t* my_smem = epi_smem + ew * EPI_PER_WARP;
      uint8_t* out_buf = my_smem;
      uint8_t* res_buf0 = my_smem + TC_TILE_BYTES;
      uint64_t* res_bar = tail->res_bar[ew];
      const int n_chunks = N / 16;                            // N % 32 == 0 and N == block_n (checked by the host)
      const int sw = (lane >> 1) & 3;
      uint8_t* my_row_out = out_buf + lane * 64;
      const uint8_t* my_row_res0 = res_buf0 + lane * 64;
      const int t_row = lane >> 2, t_piece = lane & 3;
      const int acc = grp;
      uint32_t acc_phase = 0;
      uint32_t res_issued = 0, res_used = 0;
      const float inv_n = 1.0f / (float)N;
      for (int64_t tile = unit + (int64_t)grp * n_units; tile < total_tiles; tile += 2 * n_units) {
        int row0 = (int)tile * TILE_M + quad * 32;            // n_tiles == 1
        int64_t row_lim = M;
        if (cv.tiles_x) {
          const int trow = (int)tile / cv.tiles_x;
          row0 = trow * cv.wt + ((int)tile - trow * cv.tiles_x) * TILE_M + quad * 32;
          row_lim = (int64_t)(trow + 1) * cv.wt;
        }
        if (LNF != 2 && lane == 0) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (j < n_chunks) {
              const uint32_t b = res_issued & 1;
              mbar_expect_tx(&res_bar[b], TC_TILE_BYTES);
              tma_load_2d(res_buf0 + b * TC_TILE_BYTES, &map_r, &res_bar[b], j * 16, row0);
              ++res_issued;
            }
        }
        mbar_wait(&tail->tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_stride;
        // ---- pass 1: x = acc + bias + residual -> y (fp32) and back into TMEM; row sum
        float sum = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[16];
          tmem_ld_x16(t_addr + (uint32_t)(c * 16), r);
          tmem_ld_wait();
          float v[16];
          const float4* bs = reinterpret_cast<const float4*>(bias_s + c * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = bs[i];
            v[4 * i] = __uint_as_float(r[4 * i]) + b4.x; v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4.y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4.z; v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4.w;
          }
          if constexpr (LNF != 2) {
            const uint32_t b = res_used & 1;
            mbar_wait(&res_bar[b], (res_used >> 1) & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t = *reinterpret_cast<const uint4*>(my_row_res0 + b * TC_TILE_BYTES + ((q ^ sw) << 4));
              v[4 * q] += __uint_as_float(t.x); v[4 * q + 1] += __uint_as_float(t.y);
              v[4 * q + 2] += __uint_as_float(t.z); v[4 * q + 3] += __uint_as_float(t.w);
            }
            ++res_used;
            __syncwarp();
            if (lane == 0 && c + 2 < n_chunks) {
              const uint32_t nb = res_issued & 1;
              mbar_expect_tx(&res_bar[nb], TC_TILE_BYTES);
              tma_load_2d(res_buf0 + nb * TC_TILE_BYTES, &map_r, &res_bar[nb], (c + 2) * 16, row0);
              ++res_issued;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) { sum += v[i]; r[i] = __float_as_uint(v[i]); }
          tmem_st_x16(t_addr + (uint32_t)(c * 16), r);
          if constexpr (LNF != 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(my_row_out + ((q ^ sw) << 4)) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
            __syncwarp();
            {
              uint4 t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int rr = t_row + 8 * j;
                t[j] = *reinterpret_cast<const uint4*>(out_buf + rr * 64 + ((t_piece ^ ((rr >> 1) & 3)) << 4));
              }
              uint8_t* gp = reinterpret_cast<uint8_t*>(y + (int64_t)(row0 + t_row) * N + c * 16) + t_piece * 16;
              const int64_t step = (int64_t)8 * N * 4;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (row0 + t_row + 8 * j < row_lim) *reinterpret_cast<uint4*>(gp + j * step) = t[j];
            }
            __syncwarp();
          }
        }
        tmem_st_wait();
        const float mean = sum * inv_n;
        // ---- pass 2: centred sum of squares
        float qs = 0.f;
        for (int c = 0; c < n_chunks; ++c) {
          uint32_t r[16];
          tmem_ld_x16(t_addr + (uint32_t)(c * 16), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) { const float d = __uint_as_float(r[i]) - mean; qs = fmaf(d, d, qs); }
        }
        const float rstd = rsqrtf(qs * inv_n + ln.eps);
        if constexpr (LNF == 3) {
          // ---- pass 3 (stage output): normalise and write fp32 NCHW; one coalesced 128-byte store per channel and warp
          const int64_t g = (int64_t)row0 + lane;              // my token
          const bool ok = g < row_lim;
          const int64_t bi = g / ln.hw;
          float* op = ln.out_nchw + (bi * N) * ln.hw + (g - bi * ln.hw);
          for (int c = 0; c < n_chunks; ++c) {
            uint32_t r[16];
            tmem_ld_x16(t_addr + (uint32_t)(c * 16), r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c * 16 + i;
              const float o = fmaf((__uint_as_float(r[i]) - mean) * rstd, gam_s[col], bet_s[col]);
              if (ok) op[(int64_t)col * ln.hw] = o;
            }
          }
        } else if constexpr (LNF == 2) {
          // ---- pass 3 (stem): normalise, add the position row, write fp32 -- 16 columns (64 B per row) at a time.  The
          //      position rows are fetched coalesced (four lanes per row) and handed to the row owners through smem.
          for (int c = 0; c < n_chunks; ++c) {
            {
              const uint8_t* pp = reinterpret_cast<const uint8_t*>(ln.pos) + (size_t)c * 64 + t_piece * 16;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int rr = t_row + 8 * j;
                uint4 t = make_uint4(0, 0, 0, 0);
                if (ln.pos != nullptr && row0 + rr < row_lim)
                  t = __ldg(reinterpret_cast<const uint4*>(pp + (size_t)((row0 + rr) % ln.pos_rows) * N * 4));
                *reinterpret_cast<uint4*>(out_buf + rr * 64 + ((t_piece ^ ((rr >> 1) & 3)) << 4)) = t;
              }
            }
            uint32_t r[16];
            tmem_ld_x16(t_addr + (uint32_t)(c * 16), r);
            tmem_ld_wait();
            __syncwarp();
            uint32_t o[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 p4 = *reinterpret_cast<const uint4*>(my_row_out + ((q ^ sw) << 4));
              const uint32_t pw[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int col = c * 16 + 4 * q + i;
                o[4 * q + i] = __float_as_uint(fmaf((__uint_as_float(r[4 * q + i]) - mean) * rstd, gam_s[col], bet_s[col]) +
                                               __uint_as_float(pw[i]));
              }
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(my_row_out + ((q ^ sw) << 4)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
            __syncwarp();
            {
              uint4 t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int rr = t_row + 8 * j;
                t[j] = *reinterpret_cast<const uint4*>(out_buf + rr * 64 + ((t_piece ^ ((rr >> 1) & 3)) << 4));
              }
              uint8_t* gp = reinterpret_cast<uint8_t*>(y + (int64_t)(row0 + t_row) * N + c * 16) + t_piece * 16;
              const int64_t step = (int64_t)8 * N * 4;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (row0 + t_row + 8 * j < row_lim) *reinterpret_cast<uint4*>(gp + j * step) = t[j];
            }
            __syncwarp();
          }
        } else {
        // ---- pass 3: normalise -> bf16, 32 columns (64 B per row) at a time
        for (int c = 0; c < n_chunks / 2; ++c) {
          uint32_t r[32];
          tmem_ld_x32(t_addr + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = c * 32 + 2 * i;
            const float o0 = fmaf((__uint_as_float(r[2 * i]) - mean) * rstd, gam_s[col], bet_s[col]);
            const float o1 = fmaf((__uint_as_float(r[2 * i + 1]) - mean) * rstd, gam_s[col + 1], bet_s[col + 1]);
            pk[i] = pack_bf16x2(o0, o1);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(my_row_out + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          __syncwarp();
          {
            uint4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int rr = t_row + 8 * j;
              t[j] = *reinterpret_cast<const uint4*>(out_buf + rr * 64 + ((t_piece ^ ((rr >> 1) & 3)) << 4));
            }
            uint8_t* gp = reinterpret_cast<uint8_t*>(ln.out + (int64_t)(row0 + t_row) * N + c * 32) + t_piece * 16;
            const int64_t step = (int64_t)8 * N * 2;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (row0 + t_row + 8 * j < row_lim) *reinterpret_cast<uint4*>(gp + j * step) = t[j];
          }
          __syncwarp();
        }
        }
        tc_fence_before();                                    // the accumulator is free again
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->tempty[acc]);
        acc_phase ^= 1;
      }
    } else {
    const int ew = warp - 2;
    const int quad = warp & 3;                                // TMEM lane quadrant this warp may access
    const int grp = ew / GW;                                  // group = accumulator buffer = tile parity of this CTA
    const int part = (ew % GW) >> 2;                          // the PER_QUAD warps of a quadrant interleave chunks
    uint8_t* my_smem = epi_smem + ew * EPI_PER_WARP;
    uint8_t* out_buf = my_smem;                               // [32 rows][64 B], 16-byte piece index ^= (row >> 1) & 3
    uint8_t* res_buf0 = my_smem + TC_TILE_BYTES;             // residual ring: slot b at res_buf0 + b * TC_TILE_BYTES
    uint64_t* res_bar = tail->res_bar[ew];
    const int n_chunks = block_n / CW;
    const int sw = (lane >> 1) & 3;
    uint8_t* my_row_out = out_buf + lane * 64;
    const uint8_t* my_row_res0 = res_buf0 + lane * 64;
    // read-back for the coalesced store: this lane moves 16-byte piece (lane & 3) of rows (lane >> 2) + 8 j
    const int t_row = lane >> 2, t_piece = lane & 3;
    constexpr int PIECE_ELEMS = 16 / (int)sizeof(TO);
    const int acc = grp;
    uint32_t acc_phase = 0;
    uint32_t res_issued = 0, res_used = 0;                    // residual ring counters (2 deep)
    uint32_t tile_iter = 0;                                   // tiles of this group so far
    const bool prof = (mode & 16) && blockIdx.x == 0 && ew == 0;
    long long c_w = 0, c_ld = 0, c_ma = 0, c_st = 0;
    for (int64_t tile = unit + (int64_t)grp * n_units; tile < total_tiles; tile += 2 * n_units, ++tile_iter) {
      const int m_t = (int)((uint32_t)tile / (uint32_t)n_tiles);     // total_tiles < 2^31 (checked on the host)
      const int n_t = (int)tile - m_t * n_tiles;
      int row0 = m_t * TILE_M + cta_rank * TC_BM + quad * 32;
      int64_t row_lim = M;
      if (cv.tiles_x) {                                     // token row `trow`: GEMM rows [trow * wt, trow * wt + wt)
        const int trow = m_t / cv.tiles_x;
        row0 = trow * cv.wt + (m_t - trow * cv.tiles_x) * TILE_M + cta_rank * TC_BM + quad * 32;
        row_lim = (int64_t)(trow + 1) * cv.wt;
      }
      const int col0 = n_t * block_n;
      // chunks of this warp in this tile: c = first, first + PER_QUAD, ... (first rotates with the tile count so a
      // chunk count that is not a multiple of PER_QUAD still balances the warps of a quadrant)
      const int first = (part + (int)(tile_iter % PER_QUAD) * (n_chunks % PER_QUAD)) % PER_QUAD;
      if (RES) {
        // prefetch the first two residual chunks of this tile before waiting for the accumulator
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int c = first + PER_QUAD * j;
            if (c < n_chunks) {
              const uint32_t b = res_issued & 1;
              mbar_expect_tx(&res_bar[b], TC_TILE_BYTES);
              tma_load_2d(res_buf0 + b * TC_TILE_BYTES, &map_r, &res_bar[b], col0 + c * CW, row0);
              ++res_issued;
            }
          }
        }
      }
      long long tp = prof ? clock64() : 0;
      mbar_wait(&tail->tfull[acc], acc_phase);
      tc_fence_after();
      if (prof) { const long long t1 = clock64(); c_w += t1 - tp; tp = t1; }
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_stride;
      if (first >= n_chunks) {                               // no chunk for this warp in this tile: release at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tail->tempty[acc]); else mbar_arrive(&tail->tempty[acc]); }
      }
      for (int c = first; c < n_chunks; c += PER_QUAD) {
        float v[CW];
        if constexpr (CW == 32) {
          uint32_t r[32];
          tmem_ld_x32(t_addr + (uint32_t)(c * CW), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
        } else {
          uint32_t r[16];
          tmem_ld_x16(t_addr + (uint32_t)(c * CW), r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (prof) { const long long t1 = clock64(); c_ld += t1 - tp; tp = t1; }
        if (c + PER_QUAD >= n_chunks) {                       // my last read of this accumulator: hand it back to the
          tc_fence_before();                                  // MMA warp now, the math / stores below no longer need it
          __syncwarp();
          if (lane == 0) { if (CG == 2) mbar_arrive_leader(&tail->tempty[acc]); else mbar_arrive(&tail->tempty[acc]); }
        }
        const int col = col0 + c * CW;
        {                                                     // bias (zero-filled copy in shared memory: broadcast reads)
          const float4* bs = reinterpret_cast<const float4*>(bias_s + col);
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            const float4 b4 = bs[i];
            v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
          }
        }
        if (cv.relu) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        uint32_t packed[16];                                  // the 64-byte output row of this lane
        if constexpr (CW == 32) {
          if (GELU && !RES) {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = gelu_pair_bf16(v[2 * i], v[2 * i + 1]);
          }
        }
        if (!(CW == 32 && GELU && !RES)) {
          if (GELU) {
#pragma unroll
            for (int i = 0; i < CW; ++i) v[i] = gelu_fast(v[i]);
          }
          if (RES) {
            const uint32_t b = res_used & 1;
            mbar_wait(&res_bar[b], (res_used >> 1) & 1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 t = *reinterpret_cast<const uint4*>(my_row_res0 + b * TC_TILE_BYTES + ((q ^ sw) << 4));
              const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
              if constexpr (CW == 16) {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[4 * q + i] += __uint_as_float(w4[i]);
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w4[i]);
                  v[8 * q + 2 * i] += __low2float(h);
                  v[8 * q + 2 * i + 1] += __high2float(h);
                }
              }
            }
            ++res_used;
            __syncwarp();                                     // every lane has read the tile: the slot is free
            if (lane == 0 && c + 2 * PER_QUAD < n_chunks) {   // keep the ring two chunks ahead
              const uint32_t nb = res_issued & 1;
              mbar_expect_tx(&res_bar[nb], TC_TILE_BYTES);
              tma_load_2d(res_buf0 + nb * TC_TILE_BYTES, &map_r, &res_bar[nb], col0 + (c + 2 * PER_QUAD) * CW, row0);
              ++res_issued;
            }
          }
          if constexpr (CW == 16) {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = __float_as_uint(v[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
          }
        }
        if (prof) { const long long t1 = clock64(); c_ma += t1 - tp; tp = t1; }
        // transpose through shared memory (generic proxy only: no fence, no TMA): each lane writes its 64-byte row,
        // then four lanes move one row, so every 16-byte store of the warp fills whole sectors of 8 output rows
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(my_row_out + ((q ^ sw) << 4)) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        __syncwarp();
        {
          uint4 t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {                       // all four reads first: independent registers
            const int r = t_row + 8 * j;
            t[j] = *reinterpret_cast<const uint4*>(out_buf + r * 64 + ((t_piece ^ ((r >> 1) & 3)) << 4));
          }
          if (!(mode & 1) && col + t_piece * PIECE_ELEMS < N) {
            uint8_t* gp = reinterpret_cast<uint8_t*>(y + (int64_t)(row0 + t_row) * N + col) + t_piece * 16;
            const int64_t step = (int64_t)8 * N * (int64_t)sizeof(TO);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (row0 + t_row + 8 * j < row_lim) *reinterpret_cast<uint4*>(gp + j * step) = t[j];
          }
        }
        __syncwarp();                                         // the staging row may be overwritten by the next chunk
        if (prof) { const long long t1 = clock64(); c_st += t1 - tp; tp = t1; }
      }
      acc_phase ^= 1;
    }
    if (prof && lane == 0) { g_tc_cycles[4] = c_w; g_tc_cycles[5] = c_ld; g_tc_cycles[6] = c_ma; g_tc_cycles[7] = c_st; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                            // the peer may still read this CTA's smem / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc2_rt(tmem_base, 512);
    else tmem_dealloc_rt(tmem_base, 512);
  }
}

// Tile width: a multiple of 32 (epilogue chunks), <= 256 (TMEM double buffer); 192 keeps four pipeline stages
// next to the epilogue staging, so it is preferred whenever it divides N.
// Diagnostics switch (psw_diag_linear_mode, include/panoswin_b200_debug.h): bit0 no stores, bit1 no loads, bit2 no MMAs,
// bits [8,12) stage-count override, bits [16,25) tile-width override.  It exists -- as mutable state -- only in a
// -DPSW_DIAGNOSTICS build; the product library compiles it to the constant 0, so kernel selection is stateless.
#ifdef PSW_DIAGNOSTICS
static int g_tc_mode = 0;
#else
static constexpr int g_tc_mode = 0;
#endif

static int pick_block_n(int N) {
  if ((g_tc_mode >> 16) & 0x1ff) return (g_tc_mode >> 16) & 0x1ff;
  if (N <= 256) return (N + 31) / 32 * 32;
  const int prefs[] = {192, 256, 224, 160, 128, 96, 64};
  for (int bn : prefs)
    if (N % bn == 0) return bn;
  return 192;                                                 // ragged last tile: zero-filled loads, clipped stores
}

template <bool GELU, bool RES, typename TO, int GW, int CG>
static int launch_tc_gw(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& mr, const float* bias, void* y,
                        int64_t M, int N, int K, int block_n, int stages, size_t smem, ConvView cv, cudaStream_t st) {
  if (cv.tiles_x) cv.tiles_x = (cv.wt + TC_BM * CG - 1) / (TC_BM * CG);
  const int64_t m_tiles = cv.tiles_x ? (M / cv.wt) * cv.tiles_x : (M + TC_BM * CG - 1) / (TC_BM * CG);
  const int64_t tiles = m_tiles * ((N + block_n - 1) / block_n);
  const int units = num_sms() / CG;
  const int grid = CG * (int)(tiles < units ? tiles : units);
  auto kern = linear_tc_kernel<GELU, RES, TO, GW, CG>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(EpiCfg<GW>::THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  const LnFuse no_ln = {nullptr, nullptr, nullptr, 0.f, nullptr, 1, nullptr, 1};
  PSW_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mw, mr, bias, (TO*)y, M, N, K, block_n, stages, g_tc_mode & 31, cv, no_ln));
  return launch_status("linear_tc_kernel");
}

// CTA-pair tiles (256 x block_n): worth it when the GEMM is bound by L2 -> SM operand traffic, i.e. deep K, and the
// pair tiles still fill the machine.
static bool use_pair(int64_t M, int N, int K, int* block_n) {
  if ((g_tc_mode >> 26) & 1) return false;
  const bool force = (g_tc_mode >> 27) & 1;
  if (N % 32 != 0) return false;
  int bn = 0;
  const int prefs[] = {256, 192, 128, 64};
  for (int c : prefs)
    if (N % c == 0) { bn = c; break; }
  if (!bn) return false;
  if ((g_tc_mode >> 16) & 0x1ff) bn = (g_tc_mode >> 16) & 0x1ff;
  const int64_t tiles = ((M + 255) / 256) * (N / bn);
  if (!force && (K < 768 || tiles < 2 * (num_sms() / 2))) return false;   // measured: profitable from K = 768 up
  *block_n = bn;
  return true;
}

// conv != nullptr: x is an NHWC image [rows = B*H][W][cin] viewed through ConvView (wt tokens per row, ph x pw patches)
struct ConvArgs { int BH, W, cin, ph, pw, conv3, himg, relu; };

template <bool GELU, bool RES, typename TO>
static int launch_tc(const void* x, const void* w, const float* bias, const void* residual, void* y, int64_t M, int N,
                     int K, cudaStream_t st, const ConvArgs* conv = nullptr) {
  constexpr int CW = Chunk<TO>::CW;
  int block_n = pick_block_n(N);
  const bool pair = !(conv && conv->conv3) && use_pair(M, N, K, &block_n);
  const int cg = pair ? 2 : 1;
  CUtensorMap mx, mw, mr;
  ConvView cv = {0, 0, 0, 0, 0, 0, 0};
  int rc;
  if (conv && conv->conv3) {
    const uint64_t dims[4] = {(uint64_t)conv->cin, (uint64_t)conv->W, (uint64_t)conv->himg, (uint64_t)(conv->BH / conv->himg)};
    const uint64_t strides[3] = {(uint64_t)conv->cin * 2, (uint64_t)conv->W * conv->cin * 2,
                                 (uint64_t)conv->himg * conv->W * conv->cin * 2};
    const uint32_t box[4] = {TC_BK, TC_BM, 1, 1};
    rc = make_tensor_map_nd(&mx, x, 4, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_128B);
    cv.tiles_x = 1;                                         // finalised per CTA-group size in launch_tc_gw
    cv.wt = conv->W;
    cv.kb_per_dy = conv->cin / TC_BK;
    cv.ph = 1;
    cv.conv3 = 1;
    cv.himg = conv->himg;
    cv.relu = conv->relu;
  } else if (conv) {
    const uint64_t dims[3] = {(uint64_t)conv->pw * conv->cin, (uint64_t)(conv->W / conv->pw), (uint64_t)conv->BH};
    const uint64_t strides[2] = {(uint64_t)conv->pw * conv->cin * 2, (uint64_t)conv->W * conv->cin * 2};
    const uint32_t box[3] = {TC_BK, TC_BM, 1};
    rc = make_tensor_map_nd(&mx, x, 3, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_128B);
    cv.tiles_x = 1;                                         // finalised per CTA-group size in launch_tc_gw
    cv.wt = conv->W / conv->pw;
    cv.kb_per_dy = conv->pw * conv->cin / TC_BK;
    cv.ph = conv->ph;
  } else {
    rc = make_tensor_map_2d(&mx, x, (uint64_t)M, (uint64_t)K, TC_BM, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw, w, (uint64_t)N, (uint64_t)K, (uint32_t)(block_n / cg), TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mr, RES ? residual : y, (uint64_t)M, (uint64_t)N, 32, CW, (int)sizeof(TO), CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)(block_n / cg) * TC_BK * 2;
  const size_t bias_bytes = (size_t)((N + block_n - 1) / block_n) * block_n * sizeof(float);
  const int k_blocks = (K + TC_BK - 1) / TC_BK;
  // 16 epilogue warps when the pipeline still gets enough stages next to their staging tiles, else 8
  int gw = 8, stages = 0;
  size_t fixed = 0;
  for (;; gw = 4) {
    const size_t epi_bytes = (size_t)2 * gw * (RES ? 3 : 1) * TC_TILE_BYTES;
    fixed = 1024 + epi_bytes + sizeof(TcSmemTail) + bias_bytes;
    stages = (int)((227 * 1024 - fixed) / stage_bytes);
    if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
    if (gw == 4 || stages >= 4 || stages >= 2 * k_blocks) break;
  }
  if ((g_tc_mode >> 25) & 1) gw = 4;
  if (((g_tc_mode >> 8) & 15) && ((g_tc_mode >> 8) & 15) < stages) stages = (g_tc_mode >> 8) & 15;
  PSW_REQUIRE(stages >= 2, PSW_ERR_UNSUPPORTED, "psw_linear_fwd(bf16): tile too large for shared memory");
  const size_t smem = fixed + stages * stage_bytes;
  if (pair) {
    if (gw == 8) return launch_tc_gw<GELU, RES, TO, 8, 2>(mx, mw, mr, bias, y, M, N, K, block_n, stages, smem, cv, st);
    return launch_tc_gw<GELU, RES, TO, 4, 2>(mx, mw, mr, bias, y, M, N, K, block_n, stages, smem, cv, st);
  }
  if (gw == 8) return launch_tc_gw<GELU, RES, TO, 8, 1>(mx, mw, mr, bias, y, M, N, K, block_n, stages, smem, cv, st);
  return launch_tc_gw<GELU, RES, TO, 4, 1>(mx, mw, mr, bias, y, M, N, K, block_n, stages, smem, cv, st);
}

int linear_f32(const float* x, const float* w, const float* bias, const float* residual, float* y, int64_t M, int N,
               int K, int flags, cudaStream_t st);

}  // namespace psw

using namespace psw;

#ifdef PSW_DIAGNOSTICS
extern "C" PSW_API int psw_diag_linear_mode(int mode) {
  const int old = g_tc_mode;
  g_tc_mode = mode;
  return old;
}
#endif

// LNF 1: y = x . w^T + bias + residual (fp32, may alias residual) and ln_out = LayerNorm(y) * gamma + beta (bf16).
// LNF 2 (residual == nullptr): y = LayerNorm(x . w^T + bias) * gamma + beta + pos (fp32); x may be a patch-conv view.
static int launch_tc_lnf(const void* x, const void* w, const float* bias, const void* residual, void* y, const LnFuse& ln,
                         int64_t M, int N, int K, cudaStream_t st, const ConvArgs* conv = nullptr) {
  const int block_n = N;
  CUtensorMap mx, mw, mr;
  ConvView cv = {0, 0, 0, 0};
  int rc;
  if (conv) {
    const uint64_t dims[3] = {(uint64_t)conv->pw * conv->cin, (uint64_t)(conv->W / conv->pw), (uint64_t)conv->BH};
    const uint64_t strides[2] = {(uint64_t)conv->pw * conv->cin * 2, (uint64_t)conv->W * conv->cin * 2};
    const uint32_t box[3] = {TC_BK, TC_BM, 1};
    rc = make_tensor_map_nd(&mx, x, 3, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_128B);
    cv.wt = conv->W / conv->pw;
    cv.tiles_x = (cv.wt + TC_BM - 1) / TC_BM;
    cv.kb_per_dy = conv->pw * conv->cin / TC_BK;
    cv.ph = conv->ph;
  } else {
    rc = make_tensor_map_2d(&mx, x, (uint64_t)M, (uint64_t)K, TC_BM, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw, w, (uint64_t)N, (uint64_t)K, (uint32_t)block_n, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mr, residual ? residual : y, (uint64_t)M, (uint64_t)N, 32, 16, 4, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  constexpr int GW = 4;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)block_n * TC_BK * 2;
  const size_t fixed = 1024 + (size_t)2 * GW * 3 * TC_TILE_BYTES + sizeof(TcSmemTail) + (size_t)3 * N * sizeof(float);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  PSW_REQUIRE(stages >= 2, PSW_ERR_UNSUPPORTED, "psw_linear_ln_fwd: tile too large for shared memory");
  const size_t smem = fixed + stages * stage_bytes;
  const int64_t tiles = cv.tiles_x ? (M / cv.wt) * cv.tiles_x : (M + TC_BM - 1) / TC_BM;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  if (residual && ln.out_nchw) {
    auto kern = linear_tc_kernel<false, true, float, GW, 1, 3>;
    PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, EpiCfg<GW>::THREADS, smem, st>>>(mx, mw, mr, bias, (float*)y, M, N, K, block_n, stages, 0, cv, ln);
  } else if (residual) {
    auto kern = linear_tc_kernel<false, true, float, GW, 1, 1>;
    PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, EpiCfg<GW>::THREADS, smem, st>>>(mx, mw, mr, bias, (float*)y, M, N, K, block_n, stages, 0, cv, ln);
  } else {
    // LNF 2 (no residual: LayerNorm(x . w^T + bias) + pos, the patch conv + patch_norm fusion) was measured slower than
    // the two kernels it replaces and is not instantiated
    PSW_REQUIRE(false, PSW_ERR_UNSUPPORTED, "psw_linear_ln_fwd: needs a residual");
  }
  return launch_status("linear_tc_kernel<LNF>");
}

extern "C" PSW_API int psw_linear_ln_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                         const float* ln_gamma, const float* ln_beta, float ln_eps, void* ln_out,
                                         int64_t M, int N, int K, void* stream) {
  PSW_REQUIRE(x && w && residual && y && ln_gamma && ln_beta && ln_out, PSW_ERR_BAD_ARG, "psw_linear_ln_fwd: null pointer");
  PSW_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31), PSW_ERR_BAD_ARG, "psw_linear_ln_fwd: M=%lld N=%d K=%d", (long long)M, N, K);
  PSW_REQUIRE(N % 32 == 0 && N <= 256 && K % 8 == 0, PSW_ERR_UNSUPPORTED,
              "psw_linear_ln_fwd: needs N %% 32 == 0, N <= 256 (one tile per row) and K %% 8 == 0 (N=%d K=%d)", N, K);
  PSW_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias) && aligned16(residual) && aligned16(ln_out),
              PSW_ERR_BAD_ARG, "psw_linear_ln_fwd: pointers must be 16-byte aligned");
  const LnFuse ln = {ln_gamma, ln_beta, (bf16*)ln_out, ln_eps, nullptr, 1, nullptr, 1};
  return launch_tc_lnf(x, w, bias, residual, y, ln, M, N, K, (cudaStream_t)stream);
}

// The last fc2 of a stage: y = x . w^T + bias + residual (fp32) and the stage's output map LN(y) as fp32 NCHW
// [M / HW, N, HW] (SimplePanoSwinTransformer.forward :974-978) from the same epilogue.
extern "C" PSW_API int psw_linear_ln_nchw_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                                              const float* ln_gamma, const float* ln_beta, float ln_eps, float* out_nchw,
                                              int64_t HW, int64_t M, int N, int K, void* stream) {
  PSW_REQUIRE(x && w && residual && y && ln_gamma && ln_beta && out_nchw, PSW_ERR_BAD_ARG, "psw_linear_ln_nchw_fwd: null pointer");
  PSW_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && HW > 0 && HW < (1ll << 31) && M % HW == 0, PSW_ERR_BAD_ARG,
              "psw_linear_ln_nchw_fwd: M=%lld N=%d K=%d HW=%lld (M must be a multiple of HW)", (long long)M, N, K, (long long)HW);
  PSW_REQUIRE(N % 32 == 0 && N <= 256 && K % 8 == 0, PSW_ERR_UNSUPPORTED,
              "psw_linear_ln_nchw_fwd: needs N %% 32 == 0, N <= 256 (one tile per row) and K %% 8 == 0 (N=%d K=%d)", N, K);
  PSW_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias) && aligned16(residual), PSW_ERR_BAD_ARG,
              "psw_linear_ln_nchw_fwd: pointers must be 16-byte aligned");
  const LnFuse ln = {ln_gamma, ln_beta, nullptr, ln_eps, nullptr, 1, out_nchw, (int)HW};
  return launch_tc_lnf(x, w, bias, residual, y, ln, M, N, K, (cudaStream_t)stream);
}

// Non-overlapping patch convolution (PatchEmbed.proj[6], reference :749: conv(kernel = stride = patch)) as a GEMM over
// a 3-D TMA view of the NHWC image: no im2col, bias fused in the epilogue.
extern "C" PSW_API int psw_patch_conv_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                                          int cin, int cout, int patch_h, int patch_w, void* stream) {
  PSW_REQUIRE(x && w && out, PSW_ERR_BAD_ARG, "psw_patch_conv_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0 && patch_h > 0 && patch_w > 0, PSW_ERR_BAD_ARG,
              "psw_patch_conv_fwd: bad dims");
  PSW_REQUIRE(H % patch_h == 0 && W % patch_w == 0, PSW_ERR_BAD_ARG,
              "psw_patch_conv_fwd: H=%d W=%d must be multiples of the patch %dx%d (pad the image first)", H, W, patch_h, patch_w);
  PSW_REQUIRE((patch_w * cin) % TC_BK == 0 && cout % 16 == 0, PSW_ERR_UNSUPPORTED,
              "psw_patch_conv_fwd: needs patch_w * cin %% 64 == 0 and cout %% 16 == 0 (cin=%d patch_w=%d cout=%d)", cin, patch_w, cout);
  PSW_REQUIRE(aligned16(x) && aligned16(w) && aligned16(out) && aligned16(bias), PSW_ERR_BAD_ARG,
              "psw_patch_conv_fwd: pointers must be 16-byte aligned");
  const int64_t M = (int64_t)B * (H / patch_h) * (W / patch_w);
  PSW_REQUIRE(M < (1ll << 31) && (int64_t)B * H < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_patch_conv_fwd: too many tokens");
  const ConvArgs conv = {B * H, W, cin, patch_h, patch_w, 0, 0, 0};
  return launch_tc<false, false, bf16>(x, w, bias, nullptr, out, M, cout, patch_h * patch_w * cin, (cudaStream_t)stream, &conv);
}

// 3 x 3 / stride 1 / pad 1 convolution (+ bias, optional ReLU) of an NHWC bf16 image as a GEMM over shifted TMA views
// (ConvView.conv3): the second stem convolution of the models whose stem widths the dedicated kernels of psw_stem2.cu
// are not instantiated for (PatchEmbed.proj[3:6], reference :746-748; PanoSwin-B: 42 -> 84 channels, zero-padded to
// 64 -> 96 by the caller).  w [cout][3][3][cin] bf16, out [B, H, W, cout] bf16.
extern "C" PSW_API int psw_conv3x3_nhwc_fwd(const void* x, const void* w, const float* bias, void* out, int B, int H, int W,
                                            int cin, int cout, int relu, void* stream) {
  PSW_REQUIRE(x && w && out, PSW_ERR_BAD_ARG, "psw_conv3x3_nhwc_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && cin > 0 && cout > 0, PSW_ERR_BAD_ARG, "psw_conv3x3_nhwc_fwd: bad dims");
  PSW_REQUIRE(cin % TC_BK == 0 && cout % 16 == 0, PSW_ERR_UNSUPPORTED,
              "psw_conv3x3_nhwc_fwd: needs cin %% 64 == 0 and cout %% 16 == 0 (zero-pad the channels); got cin=%d cout=%d", cin, cout);
  PSW_REQUIRE(aligned16(x) && aligned16(w) && aligned16(out) && aligned16(bias), PSW_ERR_BAD_ARG,
              "psw_conv3x3_nhwc_fwd: pointers must be 16-byte aligned");
  const int64_t M = (int64_t)B * H * W;
  PSW_REQUIRE(M < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_conv3x3_nhwc_fwd: too many pixels");
  const ConvArgs conv = {B * H, W, cin, 1, 1, 1, H, relu ? 1 : 0};
  return launch_tc<false, false, bf16>(x, w, bias, nullptr, out, M, cout, 9 * cin, (cudaStream_t)stream, &conv);
}

#ifdef PSW_DIAGNOSTICS
extern "C" PSW_API int psw_diag_linear_cycles(long long* host_out16) {
  PSW_REQUIRE(host_out16, PSW_ERR_BAD_ARG, "psw_diag_linear_cycles: null pointer");
  PSW_CUDA(cudaMemcpyFromSymbol(host_out16, g_tc_cycles, sizeof(long long) * 16));
  return PSW_OK;
}
#endif

extern "C" PSW_API int psw_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y,
                              int64_t M, int N, int K, int flags, int dtype, int out_dtype, void* stream) {
  PSW_REQUIRE(x && w && y, PSW_ERR_BAD_ARG, "psw_linear_fwd: null pointer");
  PSW_REQUIRE(M > 0 && N > 0 && K > 0, PSW_ERR_BAD_ARG, "psw_linear_fwd: M=%lld N=%d K=%d", (long long)M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSW_F32) {
    PSW_REQUIRE(out_dtype == PSW_F32, PSW_ERR_BAD_ARG, "psw_linear_fwd: fp32 path writes fp32");
    return linear_f32((const float*)x, (const float*)w, bias, (const float*)residual, (float*)y, M, N, K, flags, st);
  }
  PSW_REQUIRE(dtype == PSW_BF16, PSW_ERR_BAD_ARG, "psw_linear_fwd: unknown dtype %d", dtype);
  PSW_REQUIRE(K % 8 == 0 && N % 16 == 0, PSW_ERR_UNSUPPORTED, "psw_linear_fwd(bf16): need K %% 8 == 0 and N %% 16 == 0 (K=%d N=%d)", K, N);
  PSW_REQUIRE(M < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_linear_fwd(bf16): M too large");
  PSW_REQUIRE(aligned16(x) && aligned16(w) && aligned16(y) && aligned16(bias) && aligned16(residual), PSW_ERR_BAD_ARG,
              "psw_linear_fwd(bf16): pointers must be 16-byte aligned");
  PSW_REQUIRE(out_dtype == PSW_BF16 || out_dtype == PSW_F32, PSW_ERR_BAD_ARG, "psw_linear_fwd: unknown out_dtype %d", out_dtype);
  const bool gelu = (flags & PSW_EPI_GELU) != 0;
  const bool res = residual != nullptr;
  const int sel = (gelu ? 4 : 0) | (res ? 2 : 0) | (out_dtype == PSW_F32 ? 1 : 0);
  switch (sel) {
    case 0: return launch_tc<false, false, bf16>(x, w, bias, residual, y, M, N, K, st);
    case 1: return launch_tc<false, false, float>(x, w, bias, residual, y, M, N, K, st);
    case 2: return launch_tc<false, true, bf16>(x, w, bias, residual, y, M, N, K, st);
    case 3: return launch_tc<false, true, float>(x, w, bias, residual, y, M, N, K, st);
    case 4: return launch_tc<true, false, bf16>(x, w, bias, residual, y, M, N, K, st);
    case 5: return launch_tc<true, false, float>(x, w, bias, residual, y, M, N, K, st);
    case 6: return launch_tc<true, true, bf16>(x, w, bias, residual, y, M, N, K, st);
    default: return launch_tc<true, true, float>(x, w, bias, residual, y, M, N, K, st);
  }
}
