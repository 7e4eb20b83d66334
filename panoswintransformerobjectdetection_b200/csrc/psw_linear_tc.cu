// K2 (throughput path): y = act(x . w^T + bias) (+ residual) with bf16 operands on the 5th-generation
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM), every global access a TMA transfer.
//
// Persistent, warp-specialised CTA (one per SM, 320 or 448 threads):
//   warp 0      TMA producer : [128 x 64] x-tile and [BLOCK_N x 64] w-tile per stage, SWIZZLE_128B, mbarrier ring
//   warp 1      MMA issuer   : one thread issues 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per stage; commits free
//                              the smem slot and, per tile, publish the accumulator
//   warps 2..   epilogue     : 8 or 12 warps, two or three per TMEM lane quadrant, interleaving 64-byte-wide chunks:
//                              tcgen05.ld 32 rows x CW columns -> +bias -> [GELU] -> [+residual] -> swizzled smem
//                              tile -> TMA store (coalesced, clipped at M / N).  Residual tiles are TMA-loaded two
//                              chunks ahead into a per-warp ring, so no thread ever issues a strided global access.
// The accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Tiles are ordered n-fastest, so CTAs of one wave share the x-tile through L2.
// Rows beyond M and K beyond the tensor are zero-filled by TMA.
//
// GELU on this path: 0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4))) with a minimax fit of the exact erf GELU
// (max abs error 5.1e-5 over the real line, below bf16 resolution of the activations) and MUFU.TANH — the
// fp32 parity path keeps erff.
#include "psw_common.cuh"

namespace psw {

constexpr int TC_BM = 128;         // UMMA M
constexpr int TC_BK = 64;          // one 128-byte swizzle row of bf16
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_EPI_WARPS = 12;
constexpr int TC_TILE_BYTES = 32 * 64;      // epilogue staging tile: 32 rows x 64 B, SWIZZLE_64B

struct TcSmemTail {
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
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);           // the fitted polynomial is used on |x| <= 8; beyond, tanh saturates
  const float p = fmaf(fmaf(-3.20974528e-04f, x2, 3.68320430e-02f), x2, 7.97686932e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

template <typename TO> struct Chunk;             // CW output columns = one 64-byte row of the staging tile
template <> struct Chunk<float> { static constexpr int CW = 16; };
template <> struct Chunk<bf16> { static constexpr int CW = 32; };

// EW epilogue warps: 12 (three per TMEM lane quadrant) when there is no residual ring to stage, else 8
template <bool RES> struct EpiCfg { static constexpr int EW = RES ? 8 : 12; static constexpr int THREADS = 64 + 32 * EW; };

template <bool GELU, bool RES, typename TO>
__global__ void __launch_bounds__(EpiCfg<RES>::THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r,
                 const float* __restrict__ bias, int64_t M, int N, int K, int block_n, int stages) {
  constexpr int CW = Chunk<TO>::CW;
  constexpr int TC_EPI_WARPS = EpiCfg<RES>::EW;
  constexpr int PER_QUAD = TC_EPI_WARPS / 4;                  // warps sharing one TMEM lane quadrant
  extern __shared__ uint8_t smem_raw[];
  // align to 1024 B by OFFSETTING the __shared__ array (keeps the shared address space: LDS/STS, not generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = TC_BM * TC_BK * 2;                 // 16 KiB
  const uint32_t b_bytes = (uint32_t)block_n * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* epi_smem = smem + (size_t)stages * stage_bytes;    // per warp: 2 out tiles (+ 2 residual tiles)
  constexpr int EPI_PER_WARP = (RES ? 4 : 2) * TC_TILE_BYTES;
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(epi_smem + TC_EPI_WARPS * EPI_PER_WARP);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + block_n - 1) / block_n;
  const int64_t m_tiles = (M + TC_BM - 1) / TC_BM;
  const int64_t total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + TC_BK - 1) / TC_BK;
  const uint32_t acc_stride = 256;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_y);
    if (RES) tma_prefetch_desc(&map_r);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->tfull[s], 1);
      mbar_init(&tail->tempty[s], TC_EPI_WARPS);
    }
    for (int w = 0; w < TC_EPI_WARPS; ++w) {
      mbar_init(&tail->res_bar[w][0], 1);
      mbar_init(&tail->res_bar[w][1], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_rt(&tail->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int64_t m_t = tile / n_tiles;
        const int n_t = (int)(tile - m_t * n_tiles);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(&tail->full[stage], stage_bytes);
          tma_load_2d(sa, &map_x, &tail->full[stage], kb * TC_BK, (int)(m_t * TC_BM));
          tma_load_2d(sa + a_bytes, &map_w, &tail->full[stage], kb * TC_BK, n_t * block_n);
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer ---------------------------------
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(&tail->tempty[acc], acc_phase ^ 1);         // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_stride;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&tail->full[stage], phase);               // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t da = umma_smem_desc(sa, 16, 1024, UMMA_SWIZZLE_128B);
          const uint64_t db = umma_smem_desc(sa + a_bytes, 16, 1024, UMMA_SWIZZLE_128B);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)                // advance 32 B inside the swizzle row: +2 (>>4)
            umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&tail->empty[stage]);                   // frees the smem slot when the MMAs retire
          if (++stage == stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tail->tfull[acc]);                       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue (8 warps) --------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;                                // TMEM lane quadrant this warp may access
    const int part = ew >> 2;                                 // the PER_QUAD warps of a quadrant interleave chunks
    uint8_t* my_smem = epi_smem + ew * EPI_PER_WARP;
    uint8_t* out_buf[2] = {my_smem, my_smem + TC_TILE_BYTES};
    uint8_t* res_buf[2] = {my_smem + 2 * TC_TILE_BYTES, my_smem + 3 * TC_TILE_BYTES};
    uint64_t* res_bar = tail->res_bar[ew];
    const int n_chunks = block_n / CW;
    const int sw = (lane >> 1) & 3;                           // SWIZZLE_64B: 16-byte chunk index ^= (row >> 1) & 3
    uint8_t* my_row_out[2] = {out_buf[0] + lane * 64, out_buf[1] + lane * 64};
    const uint8_t* my_row_res[2] = {res_buf[0] + lane * 64, res_buf[1] + lane * 64};
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t out_cnt = 0;                                     // chunks stored so far (selects the out buffer)
    uint32_t res_issued = 0, res_used = 0;                    // residual ring counters (2 deep)
    uint32_t tile_iter = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_iter) {
      const int64_t m_t = tile / n_tiles;
      const int n_t = (int)(tile - m_t * n_tiles);
      const int row0 = (int)(m_t * TC_BM) + quad * 32;
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
              tma_load_2d(res_buf[b], &map_r, &res_bar[b], col0 + c * CW, row0);
              ++res_issued;
            }
          }
        }
      }
      mbar_wait(&tail->tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_stride;
      if (first >= n_chunks) {                               // no chunk for this warp in this tile: release at once
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->tempty[acc]);
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
        if (c + PER_QUAD >= n_chunks) {                       // my last read of this accumulator: hand it back to the
          tc_fence_before();                                  // MMA warp now, the math / stores below no longer need it
          __syncwarp();
          if (lane == 0) mbar_arrive(&tail->tempty[acc]);
        }
        const int col = col0 + c * CW;
        if (bias) {
#pragma unroll
          for (int i = 0; i < CW / 4; ++i) {
            if (col + 4 * i < N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
        }
        if (GELU) {
#pragma unroll
          for (int i = 0; i < CW; ++i) v[i] = gelu_fast(v[i]);
        }
        if (RES) {
          const uint32_t b = res_used & 1;
          mbar_wait(&res_bar[b], (res_used >> 1) & 1);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 t = *reinterpret_cast<const uint4*>(my_row_res[b] + ((q ^ sw) << 4));
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
          __syncwarp();                                       // every lane has read the tile: the slot is free
          if (lane == 0 && c + 2 * PER_QUAD < n_chunks) {     // keep the ring two chunks ahead
            const uint32_t nb = res_issued & 1;
            mbar_expect_tx(&res_bar[nb], TC_TILE_BYTES);
            tma_load_2d(res_buf[nb], &map_r, &res_bar[nb], col0 + (c + 2 * PER_QUAD) * CW, row0);
            ++res_issued;
          }
        }
        // stage the chunk in shared memory (64-byte rows, 64B swizzle -> conflict-free 16-byte stores)
        const uint32_t ob = out_cnt & 1;
        if (lane == 0) tma_store_wait_read<1>();              // the store issued two chunks ago has left this buffer
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 t;
          if constexpr (CW == 16) {
            t = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                           __float_as_uint(v[4 * q + 3]));
          } else {
            t = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                           pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
          }
          *reinterpret_cast<uint4*>(my_row_out[ob] + ((q ^ sw) << 4)) = t;
        }
        fence_async_shared();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_y, out_buf[ob], col, row0);       // clipped at M and N by the tensor map
          tma_store_commit();
        }
        ++out_cnt;
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();                       // all stores of this warp are complete
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, 512);
  }
}

// Tile width: a multiple of 32 (epilogue chunks), <= 256 (TMEM double buffer); 192 keeps four pipeline stages
// next to the epilogue staging, so it is preferred whenever it divides N.
static int pick_block_n(int N) {
  if (N <= 256) return (N + 31) / 32 * 32;
  const int prefs[] = {192, 256, 224, 160, 128, 96, 64};
  for (int bn : prefs)
    if (N % bn == 0) return bn;
  return 192;                                                 // ragged last tile: zero-filled loads, clipped stores
}

template <bool GELU, bool RES, typename TO>
static int launch_tc(const void* x, const void* w, const float* bias, const void* residual, void* y, int64_t M, int N,
                     int K, cudaStream_t st) {
  constexpr int CW = Chunk<TO>::CW;
  const int block_n = pick_block_n(N);
  CUtensorMap mx, mw, my, mr;
  int rc = make_tensor_map_2d(&mx, x, (uint64_t)M, (uint64_t)K, TC_BM, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw, w, (uint64_t)N, (uint64_t)K, (uint32_t)block_n, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&my, y, (uint64_t)M, (uint64_t)N, 32, CW, (int)sizeof(TO), CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mr, RES ? residual : y, (uint64_t)M, (uint64_t)N, 32, CW, (int)sizeof(TO), CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)block_n * TC_BK * 2;
  const size_t epi_bytes = (size_t)EpiCfg<RES>::EW * (RES ? 4 : 2) * TC_TILE_BYTES;
  const size_t fixed = 1024 + epi_bytes + sizeof(TcSmemTail);
  int stages = (int)((227 * 1024 - fixed) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  PSW_REQUIRE(stages >= 2, PSW_ERR_UNSUPPORTED, "psw_linear_fwd(bf16): tile too large for shared memory");
  const size_t smem = fixed + stages * stage_bytes;
  const int64_t tiles = ((M + TC_BM - 1) / TC_BM) * ((N + block_n - 1) / block_n);
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  auto kern = linear_tc_kernel<GELU, RES, TO>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, EpiCfg<RES>::THREADS, smem, st>>>(mx, mw, my, mr, bias, M, N, K, block_n, stages);
  return launch_status("linear_tc_kernel");
}

int linear_f32(const float* x, const float* w, const float* bias, const float* residual, float* y, int64_t M, int N,
               int K, int flags, cudaStream_t st);

}  // namespace psw

using namespace psw;

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
