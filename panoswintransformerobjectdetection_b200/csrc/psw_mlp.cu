// Fused MLP of a PanoSwin block (Mlp.forward + shortcut, reference simple_panoswin_transformer.py:44-61, :534):
//   x <- x + fc2(GELU(fc1(xn)))        xn [M, C] bf16 (= norm2(x)), x [M, C] fp32 in place
// for C = 96, hidden = 384 (stage 0 of every embed_dim-96 model), where the 4C hidden activation is the largest tensor
// of the block: it never leaves the SM.  Both weight matrices stay in shared memory for the kernel's lifetime.
//
// Per 128-token tile the hidden dimension is processed in FOUR chunks of 96 columns:
//   fc1 chunk   : tcgen05.mma SS, A = xn tile (K = 96 as a 64-wide SWIZZLE_128B block + a 32-wide SWIZZLE_64B block),
//                 B = 96 rows of W1, accumulator = one of two TMEM buffers H
//   GELU        : 12 warps (three per TMEM lane quadrant, 32 columns each): tcgen05.ld -> + b1 -> GELU (packed fp16,
//                 tanh fit) -> bf16 pairs -> tcgen05.st into a TMEM region P
//   fc2 partial : tcgen05.mma TS, A = P from tensor memory, B = K-blocks of W2, accumulating into D[tile & 1]
// and four dedicated final-epilogue warps add b2 and the residual and write x (fp32) with coalesced 16-byte stores
// while the GELU warps work on the next tile (details at the kernel).  warp 0 = TMA producer, warp 1 = MMA issuer.
#include "psw_common.cuh"
#include <cuda_fp16.h>
#include <type_traits>

namespace psw {

constexpr int ML_C = 96;
constexpr int ML_HID = 384;
constexpr int ML_BM = 128;
constexpr uint32_t ML_W1A = 0;                               // W1 K-block 0: [384][128 B] SWIZZLE_128B
constexpr uint32_t ML_W1B = ML_W1A + ML_HID * 128;           // W1 K-block 1: [384][64 B]  SWIZZLE_64B
constexpr uint32_t ML_W2 = ML_W1B + ML_HID * 64;             // W2: 6 K-blocks [96][128 B] SWIZZLE_128B
constexpr uint32_t ML_X = ML_W2 + 6 * ML_C * 128;            // 2 stages x ([128][128 B] + [128][64 B])
constexpr uint32_t ML_XSTAGE = ML_BM * 128 + ML_BM * 64;
constexpr uint32_t ML_STG = ML_X + 2 * ML_XSTAGE;            // 12 warps (final epilogue) x 2 KB staging
constexpr uint32_t ML_TAIL = ML_STG + 12 * 2048;


__device__ __forceinline__ uint32_t mlp_gelu_pair(float x0, float x1) {
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

// volatile: keeps the residual prefetch where it is written (ahead of the GELU work), the compiler would sink it
__device__ __forceinline__ uint4 ml_ldg_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void ml_tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// The final epilogue has its own four warps.  (A first version let sixteen epilogue warps do GELU chunks AND the final
// epilogue one after the other: MMA/loads ~76 us + GELU ~90 us (MUFU-bound) + final epilogue ~160 us (DRAM-bound) added
// up instead of overlapping: 328 us.)  The hidden dimension is cut into FOUR chunks of 96 columns, which
// frees enough tensor memory for a second accumulator: H0 [0,96) H1 [96,192) P0 [192,240) P1 [256,304) D0 [320,416)
// D1 [416,512).  fc2 of tile t accumulates into D[t & 1]; warps 14-17 (one per TMEM lane quadrant) drain it during tile
// t + 1: the accumulator row of a lane goes through a 2 KB staging tile into a four-lanes-per-row layout, where b2 and the
// residual (prefetched three 16-column chunks ahead in registers, the tile's 48 KB pulled into L2 by the producer with
// cp.async.bulk.prefetch.L2 when it loads the tile's xn) are added and written with coalesced 16-byte stores -- while
// warps 2-13 (three per quadrant, 32 columns each) keep the GELU chunks going.  The MMA-issuing thread runs a fully
// unrolled per-tile schedule with prebuilt descriptors: with run-time chunk indices its instruction stream was the
// critical path (195 -> 139 us for the MMA / synchronisation skeleton alone).
// Measured (M = 1 Mi rows): 270 us; 197 us without the final epilogue, 231 us without the
// GELU arithmetic, 144 us with neither: the final-epilogue warps are bound by bytes in flight x memory latency.
// ---------------------------------------------------------------------------------------------------------------------
// The kernel is close to issue-bound (GELU arithmetic), so a warp that waits must not spin at full rate: back off between
// polls (the producer and the final-epilogue warps are far from the critical path, the GELU warps are given a short nap).
__device__ __forceinline__ void m2_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

constexpr int M2_CH = 96;                        // hidden columns per chunk
constexpr int M2_NCH = ML_HID / M2_CH;           // 4
constexpr int M2_GELU_WARPS = 12;
constexpr int M2_F_WARPS = 4;                    // final epilogue: one per TMEM lane quadrant
constexpr int M2_THREADS = 64 + 32 * (M2_GELU_WARPS + M2_F_WARPS);
constexpr uint32_t M2_COL_H = 0, M2_COL_P0 = 192, M2_COL_P1 = 256, M2_COL_D = 320;

struct alignas(16) M2Tail {
  uint64_t w_full, x_full[2], x_empty[2], h_full[2], h_free[2], p_full[2], p_free[2], d_full[2], d_free[2];
  uint32_t tmem_base;
  alignas(16) float b1[ML_HID];
  alignas(16) float b2[ML_C];
};
constexpr size_t M2_SMEM = 1024 + ML_TAIL + sizeof(M2Tail);

__global__ void __launch_bounds__(M2_THREADS, 1)
mlp_fused_v2_kernel(const __grid_constant__ CUtensorMap map_xa, const __grid_constant__ CUtensorMap map_xb,
                    const __grid_constant__ CUtensorMap map_w1a, const __grid_constant__ CUtensorMap map_w1b,
                    const __grid_constant__ CUtensorMap map_w2, const float* __restrict__ b1, const float* __restrict__ b2,
                    float* __restrict__ x, int64_t M, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  M2Tail* tail = reinterpret_cast<M2Tail*>(smem + ML_TAIL);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles = (int)((M + ML_BM - 1) / ML_BM);

  for (int i = threadIdx.x; i < ML_HID; i += blockDim.x) tail->b1[i] = b1 ? b1[i] : 0.f;
  for (int i = threadIdx.x; i < ML_C; i += blockDim.x) tail->b2[i] = b2 ? b2[i] : 0.f;
  if (threadIdx.x == 0) {
    mbar_init(&tail->w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->x_full[s], 1);
      mbar_init(&tail->x_empty[s], 1);
      mbar_init(&tail->h_full[s], 1);
      mbar_init(&tail->h_free[s], M2_GELU_WARPS);
      mbar_init(&tail->p_full[s], M2_GELU_WARPS);
      mbar_init(&tail->p_free[s], 1);
      mbar_init(&tail->d_full[s], 1);
      mbar_init(&tail->d_free[s], M2_F_WARPS);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&tail->w_full, ML_HID * 128 + ML_HID * 64 + 6 * ML_C * 128);
      for (int j = 0; j < 2; ++j) {
        tma_load_2d(smem + ML_W1A + j * 192 * 128, &map_w1a, &tail->w_full, 0, j * 192);
        tma_load_2d(smem + ML_W1B + j * 192 * 64, &map_w1b, &tail->w_full, 64, j * 192);
      }
      for (int j = 0; j < 6; ++j) tma_load_2d(smem + ML_W2 + j * ML_C * 128, &map_w2, &tail->w_full, j * 64, 0);
      uint32_t it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int s = it & 1;
        m2_wait_backoff(&tail->x_empty[s], ((it >> 1) & 1) ^ 1, 256);
        uint8_t* xs = smem + ML_X + s * ML_XSTAGE;
        mbar_expect_tx(&tail->x_full[s], ML_XSTAGE);
        tma_load_2d(xs, &map_xa, &tail->x_full[s], 0, t * ML_BM);
        tma_load_2d(xs + ML_BM * 128, &map_xb, &tail->x_full[s], 64, t * ML_BM);
        // the tile's residual rows (48 KB, contiguous) into L2: the final-epilogue warps fetch them a tile or two later with
        // plain loads, three chunks ahead -- not enough bytes in flight to cover DRAM latency, plenty for L2 latency
        const int64_t rows = M - (int64_t)t * ML_BM < ML_BM ? M - (int64_t)t * ML_BM : ML_BM;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(x + (int64_t)t * ML_BM * ML_C),
                     "r"((uint32_t)(rows * ML_C * 4)) : "memory");
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // One thread feeds the tensor core, and its instruction stream is the kernel's critical path (with run-time chunk
      // indices -- descriptor assembly, 64-bit adds, parity arithmetic around every MMA -- the pipe sat idle 70% of the
      // time waiting for it).  So: every descriptor is built once, the four chunk steps of a tile are unrolled with
      // compile-time chunk indices (offsets, accumulate flags and most barrier parities become immediates).
      const uint32_t idesc = umma_idesc_bf16(128, M2_CH, 0, 0);          // fc1 chunk and fc2 both have N = 96
      const uint32_t sb = smem_u32(smem);
      const uint64_t d_xa = umma_smem_desc(sb + ML_X, 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t d_xb = umma_smem_desc(sb + ML_X + ML_BM * 128, 16, 512, UMMA_SWIZZLE_64B);
      const uint64_t d_w1a = umma_smem_desc(sb + ML_W1A, 16, 1024, UMMA_SWIZZLE_128B);
      const uint64_t d_w1b = umma_smem_desc(sb + ML_W1B, 16, 512, UMMA_SWIZZLE_64B);
      const uint64_t d_w2 = umma_smem_desc(sb + ML_W2, 16, 1024, UMMA_SWIZZLE_128B);
      mbar_wait(&tail->w_full, 0);
      tc_fence_after();
      const uint32_t my_tiles = (int)blockIdx.x < tiles ? (uint32_t)((tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u;
      // chunk h of the CTA's j-th tile is chunk G = 4 j + h of its sequence: buffer G & 1 = h & 1, parity (G >> 1) & 1 = (h >> 1) & 1
      auto fc1 = [&](uint32_t j, auto hc) {
        constexpr int h = decltype(hc)::value;
        const uint32_t s = j & 1;
        if (h == 0) {
          mbar_wait(&tail->x_full[s], (j >> 1) & 1);
          tc_fence_after();
        }
        mbar_wait(&tail->h_free[h & 1], ((h >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t da = d_xa + (uint64_t)(s * (ML_XSTAGE >> 4)), db = d_xb + (uint64_t)(s * (ML_XSTAGE >> 4));
        const uint32_t dH = tmem + M2_COL_H + (uint32_t)(h & 1) * M2_CH;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(dH, da + 2 * k, d_w1a + (uint64_t)(h * (M2_CH * 128 >> 4) + 2 * k), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss(dH, db + 2 * k, d_w1b + (uint64_t)(h * (M2_CH * 64 >> 4) + 2 * k), idesc, 1);
        umma_commit(&tail->h_full[h & 1]);
        if (h == M2_NCH - 1) umma_commit(&tail->x_empty[s]);
      };
      auto fc2 = [&](uint32_t jp, auto hc) {
        constexpr int hp = decltype(hc)::value;
        const uint32_t db2 = jp & 1;
        if (hp == 0) {                                       // D[jp & 1] was last used by tile jp - 2
          mbar_wait(&tail->d_free[db2], ((jp >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        mbar_wait(&tail->p_full[hp & 1], (hp >> 1) & 1);
        tc_fence_after();
        const uint32_t pA = tmem + ((hp & 1) ? M2_COL_P1 : M2_COL_P0);
        const uint32_t dD = tmem + M2_COL_D + db2 * ML_C;
#pragma unroll
        for (int i = 0; i < 6; ++i) {                        // K = 96 of this chunk: k16 steps 6 hp .. 6 hp + 5 of W2's K-blocks
          constexpr int dummy = 0; (void)dummy;
          const int kk = 6 * hp + i;
          umma_ts(dD, pA + (uint32_t)(i * 8), d_w2 + (uint64_t)((kk >> 2) * (ML_C * 128 >> 4) + 2 * (kk & 3)), idesc, (hp | i) != 0);
        }
        umma_commit(&tail->p_free[hp & 1]);
        if (hp == M2_NCH - 1) umma_commit(&tail->d_full[db2]);
      };
      using C0 = std::integral_constant<int, 0>;
      using C1 = std::integral_constant<int, 1>;
      using C2 = std::integral_constant<int, 2>;
      using C3 = std::integral_constant<int, 3>;
      for (uint32_t j = 0; j < my_tiles; ++j) {              // step: fc1 of a chunk, then fc2 of the chunk before it
        fc1(j, C0{});
        if (j > 0) fc2(j - 1, C3{});
        fc1(j, C1{});
        fc2(j, C0{});
        fc1(j, C2{});
        fc2(j, C1{});
        fc1(j, C3{});
        fc2(j, C2{});
      }
      if (my_tiles) fc2(my_tiles - 1, C3{});
    }
  } else if (warp < 2 + M2_GELU_WARPS) {
    // ------------------------------- GELU warps ---------------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int part = ew >> 2;                                // 0..2: hidden columns [32 part, 32 part + 32) of a chunk
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t G = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      for (int h = 0; h < M2_NCH; ++h, ++G) {
        const int hb = G & 1;
        m2_wait_backoff(&tail->h_full[hb], (G >> 1) & 1, 32);
        tc_fence_after();
        uint32_t r0[32];
        tmem_ld_x32(tmem + lane_base + M2_COL_H + (uint32_t)hb * M2_CH + (uint32_t)part * 32, r0);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->h_free[hb]);
        uint32_t pk[16];
        const float* bb = tail->b1 + h * M2_CH + part * 32;
        if (mode & 8) {                                      // diagnostics: no GELU arithmetic (plain bf16 rounding)
#pragma unroll
          for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            pk[i] = mlp_gelu_pair(__uint_as_float(r0[2 * i]) + bb[2 * i], __uint_as_float(r0[2 * i + 1]) + bb[2 * i + 1]);
        }
        m2_wait_backoff(&tail->p_free[hb], ((G >> 1) & 1) ^ 1, 32);    // fc2 two chunks ago has consumed this P buffer
        tc_fence_after();
        ml_tmem_st_x16(tmem + lane_base + (hb ? M2_COL_P1 : M2_COL_P0) + (uint32_t)part * 16, pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->p_full[hb]);
      }
    }
  } else {
    // ------------------------------- final-epilogue warps -----------------------
    // (Eight such warps -- two per quadrant, the whole next tile's residual in flight -- drain faster on their own, 213 vs
    // 231 us without the GELU arithmetic, but the 80-register cap of 704 threads costs the GELU warps more: 282 vs 270 us.)
    const int quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t* stg = smem + ML_STG + quad * 4096;             // two 2 KB staging tiles
    const int sw = (lane >> 1) & 3;
    const int t_row = lane >> 2, t_piece = lane & 3;
    uint4 resid[3][4];                                       // residual chunks in flight: [chunk % 3][row group]
    auto load_chunk = [&](int t, int c, uint4 (&dst)[4]) {   // 16 columns of my 32 rows, four lanes per row
      const int row0 = t * ML_BM + quad * 32;
      const float* gx = x + (int64_t)(row0 + t_row) * ML_C + c * 16 + t_piece * 4;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        dst[jj] = make_uint4(0, 0, 0, 0);
        if (t < tiles && row0 + t_row + 8 * jj < M) dst[jj] = ml_ldg_v4(gx + (int64_t)8 * jj * ML_C);
      }
    };
    uint32_t j = 0;
    int t = blockIdx.x;
    if (t < tiles && !(mode & 2)) {
#pragma unroll
      for (int c = 0; c < 3; ++c) load_chunk(t, c, resid[c]);
    }
    for (; t < tiles; t += gridDim.x, ++j) {
      const int db2 = j & 1;
      m2_wait_backoff(&tail->d_full[db2], (j >> 1) & 1, 64);
      tc_fence_after();
      if (mode & 2) {                                        // diagnostics: no final epilogue traffic
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->d_free[db2]);
        continue;
      }
      const int row0 = t * ML_BM + quad * 32;
      float* gx = x + (int64_t)(row0 + t_row) * ML_C + t_piece * 4;
      const uint32_t dD = tmem + lane_base + M2_COL_D + (uint32_t)db2 * ML_C;
      // Per 16-column chunk: my accumulator row (lane = row) goes through a 2 KB staging tile into the four-lanes-per-row
      // layout of the residual registers, where b2 and the residual are added and the 16-byte stores are coalesced.
      // One shared-memory round trip and one __syncwarp per chunk (two staging tiles), the TMEM load one chunk ahead.
      uint32_t ra[16], rb[16];
      tmem_ld_x16(dD, ra);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        uint32_t (&r)[16] = (c & 1) ? rb : ra;
        uint8_t* sbuf = stg + (c & 1) * 2048;
        tmem_ld_wait();
        if (c < 5) {
          tmem_ld_x16(dD + (uint32_t)((c + 1) * 16), (c & 1) ? ra : rb);
        } else {                                             // the accumulator has been read completely
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tail->d_free[db2]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(sbuf + lane * 64 + ((q ^ sw) << 4)) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        const float4 bias4 = *reinterpret_cast<const float4*>(tail->b2 + c * 16 + t_piece * 4);
        __syncwarp();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int rr = t_row + 8 * jj;
          const uint4 d4 = *reinterpret_cast<const uint4*>(sbuf + rr * 64 + ((t_piece ^ ((rr >> 1) & 3)) << 4));
          const uint4 x4 = resid[c % 3][jj];
          uint4 v;
          v.x = __float_as_uint(__uint_as_float(d4.x) + bias4.x + __uint_as_float(x4.x));
          v.y = __float_as_uint(__uint_as_float(d4.y) + bias4.y + __uint_as_float(x4.y));
          v.z = __float_as_uint(__uint_as_float(d4.z) + bias4.z + __uint_as_float(x4.z));
          v.w = __float_as_uint(__uint_as_float(d4.w) + bias4.w + __uint_as_float(x4.w));
          if (row0 + rr < M) *reinterpret_cast<uint4*>(gx + c * 16 + (int64_t)8 * jj * ML_C) = v;
        }
        // refill the slot: chunk c + 3 of this tile, or chunk c - 3 of my next tile
        if (c < 3) load_chunk(t, c + 3, resid[c % 3]);
        else load_chunk(t + (int)gridDim.x, c - 3, resid[c % 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace psw

using namespace psw;

// Diagnostics switch (psw_diag_mlp_mode, -DPSW_DIAGNOSTICS builds only): bit1 no final epilogue, bit3 no GELU arithmetic.
#ifdef PSW_DIAGNOSTICS
static int g_mlp_mode = 0;
extern "C" PSW_API int psw_diag_mlp_mode(int mode) {
  const int old = g_mlp_mode;
  g_mlp_mode = mode;
  return old;
}
#else
static constexpr int g_mlp_mode = 0;
#endif

extern "C" PSW_API int psw_mlp_fused_fwd(const void* xn, const void* w1, const float* b1, const void* w2, const float* b2,
                                         void* x, int64_t M, int C, int hidden, void* stream) {
  PSW_REQUIRE(xn && w1 && b1 && w2 && b2 && x, PSW_ERR_BAD_ARG, "psw_mlp_fused_fwd: null pointer");
  PSW_REQUIRE(C == ML_C && hidden == ML_HID, PSW_ERR_UNSUPPORTED,
              "psw_mlp_fused_fwd: instantiated for C = 96, hidden = 384 (got C=%d hidden=%d)", C, hidden);
  PSW_REQUIRE(M > 0 && M < (1ll << 31), PSW_ERR_BAD_ARG, "psw_mlp_fused_fwd: M=%lld", (long long)M);
  PSW_REQUIRE(aligned16(xn) && aligned16(w1) && aligned16(w2) && aligned16(x), PSW_ERR_BAD_ARG,
              "psw_mlp_fused_fwd: pointers must be 16-byte aligned");
  CUtensorMap mxa, mxb, mw1a, mw1b, mw2;
  int rc = make_tensor_map_2d(&mxa, xn, (uint64_t)M, ML_C, ML_BM, 64, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mxb, xn, (uint64_t)M, ML_C, ML_BM, 32, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw1a, w1, ML_HID, ML_C, 192, 64, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw1b, w1, ML_HID, ML_C, 192, 32, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw2, w2, ML_C, ML_HID, ML_C, 64, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int tiles = (int)((M + ML_BM - 1) / ML_BM);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  PSW_CUDA(cudaFuncSetAttribute(mlp_fused_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M2_SMEM));
  mlp_fused_v2_kernel<<<grid, M2_THREADS, M2_SMEM, (cudaStream_t)stream>>>(mxa, mxb, mw1a, mw1b, mw2, b1, b2, (float*)x, M,
                                                                          g_mlp_mode);
  return launch_status("mlp_fused_v2_kernel");
}
