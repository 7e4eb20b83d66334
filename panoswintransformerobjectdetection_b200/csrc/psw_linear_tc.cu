// K2 (throughput path): y = act(x . w^T + bias) (+ residual) with bf16 operands on the 5th-generation
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged by TMA.
//
// Persistent, warp-specialised CTA (one per SM, 192 threads):
//   warp 0      TMA producer : [128 x 64] x-tile and [BLOCK_N x 64] w-tile per stage, SWIZZLE_128B, mbarrier ring
//   warp 1      MMA issuer   : one thread issues 4 x tcgen05.mma (M=128, N=BLOCK_N, K=16) per stage; commits free
//                              the smem slot and, per tile, publish the accumulator
//   warps 2..5  epilogue     : tcgen05.ld accumulator rows -> +bias -> [GELU] -> [+residual] -> global store
// The accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1.  Tiles are ordered n-fastest, so CTAs of one wave share the x-tile through L2.
// Rows beyond M and K beyond the tensor are zero-filled by TMA; stores are guarded.
#include "psw_common.cuh"

namespace psw {

constexpr int TC_BM = 128;         // UMMA M
constexpr int TC_BK = 64;          // one 128-byte swizzle row of bf16
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

struct TcSmemTail {
  uint64_t full[TC_MAX_STAGES];
  uint64_t empty[TC_MAX_STAGES];
  uint64_t tfull[2];
  uint64_t tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tmem_alloc_rt(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

template <typename TO> struct OutVec;
template <> struct OutVec<float> {
  static __device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p) + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  }
  static __device__ __forceinline__ void store16(float* p, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(p)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
};
template <> struct OutVec<bf16> {
  static __device__ __forceinline__ void load16(const bf16* p, float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p) + i);
      const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[8 * i + 2 * j] = __low2float(h);
        v[8 * i + 2 * j + 1] = __high2float(h);
      }
    }
  }
  static __device__ __forceinline__ void store16(bf16* p, const float (&v)[16]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      uint4 t;
      t.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
      t.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      t.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
      t.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      reinterpret_cast<uint4*>(p)[i] = t;
    }
  }
};

template <bool GELU, typename TO>
__global__ void __launch_bounds__(TC_THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                 const float* __restrict__ bias, const TO* __restrict__ residual, TO* __restrict__ y, int64_t M, int N,
                 int K, int block_n, int stages, int tmem_cols) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = TC_BM * TC_BK * 2;                 // 16 KiB
  const uint32_t b_bytes = (uint32_t)block_n * TC_BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  TcSmemTail* tail = reinterpret_cast<TcSmemTail*>(smem + (size_t)stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (N + block_n - 1) / block_n;
  const int64_t m_tiles = (M + TC_BM - 1) / TC_BM;
  const int64_t total_tiles = m_tiles * n_tiles;
  const int k_blocks = (K + TC_BK - 1) / TC_BK;
  const uint32_t acc_stride = (uint32_t)tmem_cols >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&tail->full[s], 1);
      mbar_init(&tail->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tail->tfull[s], 1);
      mbar_init(&tail->tempty[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc_rt(&tail->tmem_base, (uint32_t)tmem_cols);
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
    // ------------------------------- epilogue (4 warps) --------------------------
    const int quad = warp & 3;                                // TMEM lane quadrant this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int64_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int64_t m_t = tile / n_tiles;
      const int n_t = (int)(tile - m_t * n_tiles);
      mbar_wait(&tail->tfull[acc], acc_phase);
      tc_fence_after();
      const int64_t row = m_t * TC_BM + quad * 32 + lane;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc * acc_stride;
      for (int c0 = 0; c0 < block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld_x16(t_addr + (uint32_t)c0, r);
        tmem_ld_wait();
        const int col = n_t * block_n + c0;
        if (row < M && col < N) {
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
          if (bias) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col) + i);
              v[4 * i] += b4.x; v[4 * i + 1] += b4.y; v[4 * i + 2] += b4.z; v[4 * i + 3] += b4.w;
            }
          }
          if (GELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = gelu_erf(v[i]);
          }
          if (residual) {
            float rr[16];
            OutVec<TO>::load16(residual + row * N + col, rr);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += rr[i];
          }
          OutVec<TO>::store16(y + row * N + col, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_rt(tmem_base, (uint32_t)tmem_cols);
  }
}

// Largest tile width <= 256 that is a multiple of 16 and divides N (falls back to 256/zero-filled tail).
static int pick_block_n(int N) {
  if (N <= 256) return N;
  for (int bn = 256; bn >= 64; bn -= 16)
    if (N % bn == 0) return bn;
  return 256;
}

template <bool GELU, typename TO>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const float* bias, const void* residual, void* y,
                     int64_t M, int N, int K, int block_n, cudaStream_t st) {
  const size_t stage_bytes = (size_t)TC_BM * TC_BK * 2 + (size_t)block_n * TC_BK * 2;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  PSW_REQUIRE(stages >= 2, PSW_ERR_UNSUPPORTED, "psw_linear_fwd(bf16): tile too large for shared memory");
  const size_t smem = 1024 + stages * stage_bytes + sizeof(TcSmemTail);
  int pow2 = 32;
  while (pow2 < block_n) pow2 <<= 1;
  const int tmem_cols = 2 * pow2;                                      // double-buffered accumulator
  const int64_t tiles = ((M + TC_BM - 1) / TC_BM) * ((N + block_n - 1) / block_n);
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  auto kern = linear_tc_kernel<GELU, TO>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, TC_THREADS, smem, st>>>(mx, mw, bias, (const TO*)residual, (TO*)y, M, N, K, block_n, stages, tmem_cols);
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
  const int block_n = pick_block_n(N);
  CUtensorMap mx, mw;
  int rc = make_tensor_map_2d(&mx, x, (uint64_t)M, (uint64_t)K, TC_BM, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mw, w, (uint64_t)N, (uint64_t)K, (uint32_t)block_n, TC_BK, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const bool gelu = (flags & PSW_EPI_GELU) != 0;
  if (out_dtype == PSW_BF16)
    return gelu ? launch_tc<true, bf16>(mx, mw, bias, residual, y, M, N, K, block_n, st)
                : launch_tc<false, bf16>(mx, mw, bias, residual, y, M, N, K, block_n, st);
  PSW_REQUIRE(out_dtype == PSW_F32, PSW_ERR_BAD_ARG, "psw_linear_fwd: unknown out_dtype %d", out_dtype);
  return gelu ? launch_tc<true, float>(mx, mw, bias, residual, y, M, N, K, block_n, st)
              : launch_tc<false, float>(mx, mw, bias, residual, y, M, N, K, block_n, st);
}
