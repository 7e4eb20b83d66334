// Weight gradient of a linear layer on tcgen05:  dW[N, K] = dy[M, N]^T . x[M, K], bf16 operands,
// fp32 accumulation in tensor memory.  The contraction runs over the ROWS of two row-major tensors, so both MMA
// operands are MN-major: a TMA box {64 columns, 64 rows} of dy (resp. x) lands in shared memory as [64 m][128 B] with
// the 128-byte swizzle, which is exactly the canonical MN-major SWIZZLE_128B atom stack (8 m-rows x 64 elements per
// atom; SBO = 1024 B between 8-row groups, LBO = 8192 B between 64-column blocks) -- no transposed copy of the
// activations is ever made.  One CTA owns a 128 (n) x 128 (k) tile of dW for one slice of the rows (the output is tiny
// and the contraction huge, so the rows are split over the grid) and adds its tile to dW with fp32 reductions.
//   warp 0: TMA producer (4-stage mbarrier ring, 32 KB per stage)   warp 1: MMA issuer (4 x K=16 per stage), TMEM owner
//   warps 2-5: epilogue, one per TMEM lane quadrant: tcgen05.ld -> red.global.add.v4.f32 (16-byte vector reductions)
#include "psw_common.cuh"

namespace psw {

constexpr int WG_TM = 128, WG_TN = 128, WG_BK = 64, WG_STAGES = 4;
constexpr int WG_BLOCK = 64 * 128;                    // one box: 64 rows x 128 B
constexpr int WG_STAGE_BYTES = 4 * WG_BLOCK;          // A: 2 column blocks, B: 2 column blocks
constexpr int WG_THREADS = 6 * 32;

struct alignas(16) WgTail {
  uint64_t full[WG_STAGES], empty[WG_STAGES], acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, float* __restrict__ dw,
                int N, int K, int chunks, int chunks_per_split) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  WgTail* tail = reinterpret_cast<WgTail*>(smem + WG_STAGES * WG_STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * WG_TM, k0 = blockIdx.y * WG_TN;
  const int c_begin = blockIdx.z * chunks_per_split;
  const int c_end = c_begin + chunks_per_split < chunks ? c_begin + chunks_per_split : chunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    mbar_init(&tail->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<WG_TN>(&tail->tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&tail->empty[stage], phase ^ 1);
        mbar_expect_tx(&tail->full[stage], WG_STAGE_BYTES);
        uint8_t* st = smem + stage * WG_STAGE_BYTES;
        const int m0 = c * WG_BK;
        tma_load_2d(st, &map_dy, &tail->full[stage], n0, m0);                       // out-of-range boxes are zero-filled
        tma_load_2d(st + WG_BLOCK, &map_dy, &tail->full[stage], n0 + 64, m0);
        tma_load_2d(st + 2 * WG_BLOCK, &map_x, &tail->full[stage], k0, m0);
        tma_load_2d(st + 3 * WG_BLOCK, &map_x, &tail->full[stage], k0 + 64, m0);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(WG_TM, WG_TN, 1, 1);                    // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; ++c) {
        mbar_wait(&tail->full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * WG_STAGE_BYTES);
#pragma unroll
        for (int kk = 0; kk < WG_BK / 16; ++kk) {                                    // 16 rows = two 8-row groups: +2048 B
          const uint64_t da = umma_smem_desc(sa + kk * 2048, WG_BLOCK, 1024, UMMA_SWIZZLE_128B);
          const uint64_t db = umma_smem_desc(sa + 2 * WG_BLOCK + kk * 2048, WG_BLOCK, 1024, UMMA_SWIZZLE_128B);
          umma_ss(tmem_base, da, db, idesc, (c > c_begin || kk > 0) ? 1u : 0u);
        }
        umma_commit(&tail->empty[stage]);
        if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tail->acc_full);
    }
  } else {
    const int q = warp & 3;                                                          // TMEM lane quadrant of this warp
    const int n = n0 + q * 32 + lane;
    if (c_end > c_begin) {
      mbar_wait(&tail->acc_full, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < WG_TN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v);
        tmem_ld_wait();
        if (n < N) {                                         // K % 8 == 0: a group of four columns is inside or outside
          float* row = dw + (size_t)n * K + k0 + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (k0 + cc * 32 + j < K)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + j), "f"(__uint_as_float(v[j])),
                           "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3])) : "memory");
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<WG_TN>(tmem_base);
  }
}

// dw must be zero on entry (the caller clears it); N % 8 == 0 and K % 8 == 0 (16-byte row pitch for TMA)
int wgrad_tc(const bf16* dy, const bf16* x, float* dw, int64_t M, int N, int K, cudaStream_t st) {
  CUtensorMap mdy, mx;
  int rc = make_tensor_map_2d(&mdy, dy, (uint64_t)M, (uint64_t)N, 64, 64, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tensor_map_2d(&mx, x, (uint64_t)M, (uint64_t)K, 64, 64, 2, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const int tiles = ((N + WG_TM - 1) / WG_TM) * ((K + WG_TN - 1) / WG_TN);
  const int chunks = (int)((M + WG_BK - 1) / WG_BK);
  int splits = (3 * num_sms() + tiles - 1) / tiles;                                  // ~3 CTAs per SM in total
  if (splits > (chunks + 7) / 8) splits = (chunks + 7) / 8;                          // at least 8 chunks (512 rows) per CTA
  if (splits < 1) splits = 1;
  const int per = (chunks + splits - 1) / splits;
  splits = (chunks + per - 1) / per;
  const size_t smem = 1024 + (size_t)WG_STAGES * WG_STAGE_BYTES + sizeof(WgTail);
  PSW_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((N + WG_TM - 1) / WG_TM, (K + WG_TN - 1) / WG_TN, splits);
  wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(mdy, mx, dw, N, K, chunks, per);
  return launch_status("wgrad_tc_kernel");
}

}  // namespace psw
