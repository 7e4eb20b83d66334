// Stem, second convolution: conv3x3 (32 -> COUT = 32 or 64, pad 1) + folded BatchNorm + ReLU on NHWC bf16 activations
// (PatchEmbed.proj[3:6], reference simple_panoswin_transformer.py:744-746) as a tcgen05 implicit GEMM WITHOUT im2col.
//
// One pixel is one 64-byte K-major row (32 channels).  A [rows+2] x [cols+2] x 32 patch of the input is one 4-D TMA
// box (zero-filled outside the image = the convolution's zero padding), stored with the 64-byte swizzle.  For tap
// (dy, dx) the A operand of 128 consecutive output pixels of one image row is simply the patch viewed from pixel
// (row + dy, dx) on: the shared-memory descriptor start address moves by whole 64-byte rows.  tcgen05 applies the
// swizzle XOR to absolute shared-memory address bits (tools/probes/umma_shift_probe.cu), so such shifted views read
// exactly what TMA wrote.  Per 128 pixels: 9 taps x 2 K-halves = 18 MMAs (M=128, N=COUT, K=16) into one COUT-column
// fp32 accumulator in tensor memory; weights (9 x [COUT out][32 in] bf16) stay in shared memory.
//
//   warp 0     TMA producer: one box per tile (4 rows x 128 px of output -> 6 x 130 px patch), 3-deep ring
//   warp 1     MMA issuer  : 72 MMAs per tile into 4 of 8 accumulators (tile parity), commits per image row
//   warps 2-9  epilogue    : warp (quadrant q, group g) drains rows r = g, g + 2 of a tile: tcgen05.ld -> + bias ->
//                            ReLU -> bf16 -> transposed through a swizzled smem tile -> 16-byte stores covering
//                            contiguous 512-byte runs of the NHWC output
// HBM-bound: 64 B in + 64 B out per pixel.
#include "psw_common.cuh"

namespace psw {

constexpr int S2_ROWS = 4;                    // output rows per tile
constexpr int S2_COLS = 128;                  // output pixels per row-tile = UMMA M
constexpr int S2_PW = S2_COLS + 2;            // patch width
constexpr int S2_PR = S2_ROWS + 2;            // patch rows
constexpr int S2_PATCH_BYTES = S2_PR * S2_PW * 64;                       // 49,920
constexpr int S2_PATCH_PITCH = (S2_PATCH_BYTES + 1023) / 1024 * 1024;    // 50,176
constexpr int S2_STAGES = 3;
constexpr int S2_EPI_WARPS = 8;
constexpr int S2_THREADS = 64 + 32 * S2_EPI_WARPS;

struct alignas(16) S2Tail {
  uint64_t full[S2_STAGES];
  uint64_t empty[S2_STAGES];
  uint64_t acc_full[8];
  uint64_t acc_empty[8];
  float bias[64];
  uint32_t tmem_base;
};

template <int COUT> constexpr size_t s2_smem() {
  return 1024 + (size_t)S2_STAGES * S2_PATCH_PITCH + 9 * COUT * 64 + S2_EPI_WARPS * 32 * COUT * 2 + sizeof(S2Tail);
}

template <int COUT>
__global__ void __launch_bounds__(S2_THREADS, 1)
stem_conv2_kernel(const __grid_constant__ CUtensorMap map_in, const bf16* __restrict__ w_taps, const float* __restrict__ bias,
                  bf16* __restrict__ out, int B, int H, int W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* patches = smem;
  uint8_t* wsm = smem + (size_t)S2_STAGES * S2_PATCH_PITCH;                   // 9 x [32 out][64 B], 64B swizzle
  constexpr int W_BYTES = 9 * COUT * 64;
  constexpr int STG_BYTES = 32 * COUT * 2;                                    // per epilogue warp: COUT/32 x [32 px][64 B]
  uint8_t* epi = wsm + W_BYTES;
  S2Tail* tail = reinterpret_cast<S2Tail*>(epi + S2_EPI_WARPS * STG_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_x = (W + S2_COLS - 1) / S2_COLS;
  const int tiles_y = (H + S2_ROWS - 1) / S2_ROWS;
  const int total = B * tiles_y * tiles_x;

  // weights: global [tap][out][in] bf16 -> swizzled K-major rows (16-byte chunk ^= (address >> 7) & 3)
  for (int i = threadIdx.x; i < 9 * COUT * 4; i += blockDim.x) {
    const int row = i >> 2, ch = i & 3;                                        // row = tap * COUT + out
    const uint32_t off = (uint32_t)row * 64;
    const uint4 v = *reinterpret_cast<const uint4*>(w_taps + (size_t)row * 32 + ch * 8);
    *reinterpret_cast<uint4*>(wsm + off + ((ch ^ ((off >> 7) & 3)) << 4)) = v;
  }
  if (threadIdx.x < COUT) tail->bias[threadIdx.x] = bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_in);
    for (int s = 0; s < S2_STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    for (int a = 0; a < 8; ++a) { mbar_init(&tail->acc_full[a], 1); mbar_init(&tail->acc_empty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tail->tmem_base)), "r"((uint32_t)(8 * COUT)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_shared();                       // the weight tile was written through the generic proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int tx = tile % tiles_x;
        const int ty = (tile / tiles_x) % tiles_y;
        const int b = tile / (tiles_x * tiles_y);
        mbar_wait(&tail->empty[stage], phase ^ 1);
        mbar_expect_tx(&tail->full[stage], S2_PATCH_BYTES);
        tma_load_4d(patches + (size_t)stage * S2_PATCH_PITCH, &map_in, &tail->full[stage], 0, tx * S2_COLS - 1,
                    ty * S2_ROWS - 1, b);
        if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, COUT, 0, 0);
      const uint64_t dw0 = umma_smem_desc(smem_u32(wsm), 16, 512, UMMA_SWIZZLE_64B);
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        mbar_wait(&tail->full[stage], phase);
        tc_fence_after();
        const uint32_t pa = smem_u32(patches + (size_t)stage * S2_PATCH_PITCH);
        const uint32_t acc_phase = (it >> 1) & 1;
#pragma unroll 1
        for (int r = 0; r < S2_ROWS; ++r) {
          const int a = (int)(it & 1) * 4 + r;
          mbar_wait(&tail->acc_empty[a], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)a * COUT;
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int dy = t / 3, dx = t % 3;
            const uint64_t da = umma_smem_desc(pa + (uint32_t)(((r + dy) * S2_PW + dx) * 64), 16, 512, UMMA_SWIZZLE_64B);
            const uint64_t dw = dw0 + (uint64_t)((t * COUT * 64) >> 4);
            umma_ss(d_tmem, da, dw, idesc, t != 0);
            umma_ss(d_tmem, da + 2, dw + 2, idesc, true);
          }
          umma_commit(&tail->acc_full[a]);
        }
        umma_commit(&tail->empty[stage]);
        if (++stage == S2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int grp = ew >> 2;                                   // rows grp, grp + 2 of every tile
    uint8_t* stg = epi + ew * STG_BYTES;
    const uint32_t stg_a = smem_u32(stg);
    const int t_row = lane >> 2, t_piece = lane & 3;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int b = tile / (tiles_x * tiles_y);
      const uint32_t acc_phase = (it >> 1) & 1;
#pragma unroll 1
      for (int r = grp; r < S2_ROWS; r += 2) {
        const int a = (int)(it & 1) * 4 + r;
        mbar_wait(&tail->acc_full[a], acc_phase);
        tc_fence_after();
        uint32_t v[COUT];
#pragma unroll
        for (int h = 0; h < COUT / 32; ++h)
          tmem_ld_x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * COUT + h * 32),
                      *reinterpret_cast<uint32_t(*)[32]>(v + 32 * h));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tail->acc_empty[a]);
        const __nv_bfloat162 zero = __float2bfloat162_rn(0.0f);
        const uint32_t my = (uint32_t)lane * 64;
#pragma unroll
        for (int h = 0; h < COUT / 32; ++h) {                                  // channel half h -> staging tile h
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int c = 32 * h + 8 * q + 2 * j;
              const uint32_t p = pack_bf16x2(__uint_as_float(v[c]) + tail->bias[c], __uint_as_float(v[c + 1]) + tail->bias[c + 1]);
              const __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&p), zero);
              pk[j] = *reinterpret_cast<const uint32_t*>(&m);
            }
            const uint32_t o = (uint32_t)h * 2048 + my;
            *reinterpret_cast<uint4*>(stg + o + ((q ^ (((stg_a + o) >> 7) & 3)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        __syncwarp();
        const int y = ty * S2_ROWS + r;
        const int x0 = tx * S2_COLS + quad * 32;
        uint4 t[4 * (COUT / 32)];
#pragma unroll
        for (int h = 0; h < COUT / 32; ++h)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t o = (uint32_t)h * 2048 + (uint32_t)(t_row + 8 * j) * 64;
            t[4 * h + j] = *reinterpret_cast<const uint4*>(stg + o + ((t_piece ^ (((stg_a + o) >> 7) & 3)) << 4));
          }
        if (y < H) {
          uint8_t* gp = reinterpret_cast<uint8_t*>(out + (((int64_t)b * H + y) * W + x0) * COUT) + t_piece * 16;
#pragma unroll
          for (int h = 0; h < COUT / 32; ++h)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (x0 + t_row + 8 * j < W) *reinterpret_cast<uint4*>(gp + (t_row + 8 * j) * (COUT * 2) + h * 64) = t[4 * h + j];
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(8 * COUT)) : "memory");
  }
}

}  // namespace psw

using namespace psw;

template <int COUT>
static int launch_conv2(const CUtensorMap& map_in, const void* w_taps, const float* bias, void* out, int B, int H, int W,
                        cudaStream_t st) {
  const int tiles = B * ((H + S2_ROWS - 1) / S2_ROWS) * ((W + S2_COLS - 1) / S2_COLS);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  auto kern = stem_conv2_kernel<COUT>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s2_smem<COUT>()));
  kern<<<grid, S2_THREADS, s2_smem<COUT>(), st>>>(map_in, (const bf16*)w_taps, bias, (bf16*)out, B, H, W);
  return launch_status("stem_conv2_kernel");
}

extern "C" PSW_API int psw_stem_conv3x3_c32_relu_fwd(const void* x, const void* w_taps, const float* bias, void* out,
                                                     int B, int H, int W, int cout, void* stream) {
  PSW_REQUIRE(x && w_taps && bias && out, PSW_ERR_BAD_ARG, "psw_stem_conv3x3_c32_relu_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0, PSW_ERR_BAD_ARG, "psw_stem_conv3x3_c32_relu_fwd: B=%d H=%d W=%d", B, H, W);
  PSW_REQUIRE(cout == 32 || cout == 64, PSW_ERR_UNSUPPORTED, "psw_stem_conv3x3_c32_relu_fwd: cout %d (32 or 64)", cout);
  PSW_REQUIRE(aligned16(x) && aligned16(w_taps) && aligned16(out), PSW_ERR_BAD_ARG,
              "psw_stem_conv3x3_c32_relu_fwd: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap map_in;
  const uint64_t dims[4] = {32, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {64, (uint64_t)W * 64, (uint64_t)H * W * 64};
  const uint32_t box[4] = {32, S2_PW, S2_PR, 1};
  int rc = make_tensor_map_nd(&map_in, x, 4, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  if (cout == 64) return launch_conv2<64>(map_in, w_taps, bias, out, B, H, W, st);
  return launch_conv2<32>(map_in, w_taps, bias, out, B, H, W, st);
}
