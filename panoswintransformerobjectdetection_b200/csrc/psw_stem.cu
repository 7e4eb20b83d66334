// Stem, first layer: conv3x3(3 -> COUT = 32 or 64, pad 1) + folded BatchNorm + ReLU as a tcgen05 implicit GEMM that reads the
// fp32 NCHW image directly and writes bf16 NHWC (reference: PatchEmbed.proj[0..2],
// simple_panoswin_transformer.py:743-745, eval mode).  HBM-bound: 12 B read + 64 B written per pixel.
//
// A CTA owns tiles of 8 x 64 output pixels.  The (8+2) x (64+2) x 3 input patch is staged once in shared memory as
// bf16; for each of the tile's four 128-pixel sub-tiles (2 image rows) every thread writes the im2col row of ITS
// pixel (27 taps, zero padded to K = 32) into the 64B-swizzled K-major UMMA layout, one thread issues
// tcgen05.mma M=128 N=32 K=32 (two K=16 instructions) against the resident weight tile, and every thread reads its
// pixel's 32 output channels back from TMEM (lane = pixel), adds the folded bias, applies ReLU and stores 64
// contiguous bytes (a warp stores 2 KB contiguous).  All four im2col tiles are built before ONE proxy fence + barrier
// per tile, the eight MMAs go to four TMEM accumulators, and the next tile's patch is prefetched with cp.async
// (16 B groups) while they run.  COUT = 32 (embed_dim 96): 52 KB smem, 128 TMEM columns, 4 CTAs per SM.  COUT = 64
// serves the other stem widths (PanoSwin-B: 42 channels, zero-padded to 64 by the caller): a second staging tile per
// sub-tile for channels 32..63, two TMA box stores per sub-tile, 2 CTAs per SM.
#include "psw_common.cuh"

namespace psw {

constexpr int ST_TH = 8, ST_TW = 64;             // output tile
constexpr int ST_PH = ST_TH + 2;                 // patch rows  y0-1 .. y0+8
constexpr int ST_PQ = 18;                        // float4 groups per patch row: columns x0-4 .. x0+67
constexpr int ST_PP = 4 * ST_PQ;                 // patch row pitch in floats (72)
constexpr int ST_CIN = 3, ST_K = 32;
constexpr int ST_THREADS = 128;
constexpr int ST_PATCH_FLOATS = ST_CIN * ST_PH * ST_PP;

template <int ST_COUT>
__global__ void __launch_bounds__(ST_THREADS)
stem_conv1_kernel(const float* __restrict__ img, const float* __restrict__ wf, const float* __restrict__ bf,
                  const __grid_constant__ CUtensorMap map_out, int B, int H, int W) {
  constexpr int NSUB = ST_TH / 2;                                  // four 128-pixel sub-tiles (2 image rows each)
  extern __shared__ uint8_t st_smem_raw[];
  uint8_t* st_smem = st_smem_raw + ((1024u - (smem_u32(st_smem_raw) & 1023u)) & 1023u);
  uint8_t (*a_tile)[128 * 64] = reinterpret_cast<uint8_t (*)[128 * 64]>(st_smem);   // [NSUB] im2col tiles, SWIZZLE_64B
  constexpr int NCH = ST_COUT / 32;                                                 // 32-channel groups of the output
  uint8_t (*o_tile)[128 * 64] = a_tile + NSUB;                                      // [NSUB] staging of channels 32..63 (COUT = 64)
  uint8_t* w_tile = st_smem + NCH * NSUB * 128 * 64;                                // weights [COUT out][32 k], SWIZZLE_64B
  float (*patch)[ST_PATCH_FLOATS] = reinterpret_cast<float (*)[ST_PATCH_FLOATS]>(w_tile + ST_COUT * 64);  // [2] fp32 patches
  float* bias_s = reinterpret_cast<float*>(patch + 2);
  uint64_t& bar = *reinterpret_cast<uint64_t*>(bias_s + ST_COUT);
  uint32_t& tmem_slot = *reinterpret_cast<uint32_t*>(bias_s + ST_COUT + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  // weights: element (n, k) -> byte n*64 + ((k/8) ^ ((n>>1)&3))*16 + (k%8)*2, k = c*9 + ky*3 + kx for k < 27;
  // k = 27 / 28 carry the folded bias as a bf16 hi + lo pair (the im2col rows hold 1.0 there), so the tensor core
  // adds the bias; k >= 29 is zero
  for (int i = tid; i < ST_COUT * ST_K; i += ST_THREADS) {
    const int n = i / ST_K, k = i - n * ST_K;
    float v = 0.f;
    if (k < 27) v = wf[n * 27 + k];
    else if (k == 27) v = __bfloat162float(__float2bfloat16_rn(bf[n]));
    else if (k == 28) v = bf[n] - __bfloat162float(__float2bfloat16_rn(bf[n]));
    *reinterpret_cast<bf16*>(w_tile + n * 64 + (((k >> 3) ^ ((n >> 1) & 3)) << 4) + (k & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(4 * ST_COUT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_shared();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int tiles_x = (W + ST_TW - 1) / ST_TW, tiles_y = (H + ST_TH - 1) / ST_TH;
  const int64_t n_tiles = (int64_t)B * tiles_y * tiles_x;
  const int ry = tid >> 6, px = tid & 63;                          // my pixel inside a 2 x 64 sub-tile

  // request the fp32 patch of `tile` into buffer `pb`: 16-byte groups, zero outside the image (= conv zero padding;
  // W is a multiple of 4 and x0 of 64, so a group is either fully inside or fully outside)
  auto issue_patch = [&](int64_t tile, int pb) {
    const int tx = (int)(tile % tiles_x);
    const int ty = (int)((tile / tiles_x) % tiles_y);
    const int b = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int y0 = ty * ST_TH, x0 = tx * ST_TW;
    for (int i = tid; i < ST_CIN * ST_PH * ST_PQ; i += ST_THREADS) {
      const int q = i % ST_PQ;
      const int r = (i / ST_PQ) % ST_PH;
      const int c = i / (ST_PQ * ST_PH);
      const int y = y0 - 1 + r, x = x0 - 4 + 4 * q;
      float* dst = &patch[pb][(c * ST_PH + r) * ST_PP + 4 * q];
      if (y >= 0 && y < H && x >= 0 && x < W) cp_async16(dst, img + (((int64_t)b * ST_CIN + c) * H + y) * W + x);
      else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  if ((int64_t)blockIdx.x < n_tiles) issue_patch(blockIdx.x, 0);
  cp_async_commit();
  uint32_t par = 0;
  int it = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int pb = it & 1;
    const int tx = (int)(tile % tiles_x);
    const int ty = (int)((tile / tiles_x) % tiles_y);
    const int b = (int)(tile / ((int64_t)tiles_x * tiles_y));
    const int y0 = ty * ST_TH, x0 = tx * ST_TW;
    cp_async_wait<0>();                                             // my share of this tile's patch has landed
    if (tid == 0) tma_store_wait_read<0>();                         // the previous tile's TMA stores have left a_tile
    __syncthreads();                                                // ... and everyone else's; previous epilogues done
    const float* pt = patch[pb];
    // ---- im2col rows of my pixel in all four sub-tiles: k = c*9 + ky*3 + kx; patch column = px + kx + 3.
    //      Two sub-tiles at a time share patch rows (5 rows x 3 taps x 3 channels = 45 loads for 54 taps).
#pragma unroll
    for (int half = 0; half < NSUB / 2; ++half) {
      float pv[ST_CIN][5][3];
#pragma unroll
      for (int c = 0; c < ST_CIN; ++c)
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) pv[c][r][kx] = pt[(c * ST_PH + 4 * half + ry + r) * ST_PP + px + kx + 3];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int sub = 2 * half + s2;
        uint32_t kv[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          float lo = 0.f, hi = 0.f;
          {
            const int k = 2 * kk;
            if (k < 27) lo = pv[k / 9][2 * s2 + (k % 9) / 3][k % 3];
          }
          {
            const int k = 2 * kk + 1;
            if (k < 27) hi = pv[k / 9][2 * s2 + (k % 9) / 3][k % 3];
            else if (k == 27) hi = 1.0f;                           // bias (hi part) column
          }
          if (2 * kk == 28) lo = 1.0f;                             // bias (lo part) column
          kv[kk] = pack_bf16x2(lo, hi);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(a_tile[sub] + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) =
              make_uint4(kv[4 * c], kv[4 * c + 1], kv[4 * c + 2], kv[4 * c + 3]);
      }
    }
    fence_async_shared();                                           // one proxy fence per tile
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = umma_idesc_bf16(128, ST_COUT, 0, 0);
      const uint64_t db = umma_smem_desc(smem_u32(w_tile), 16, 512, UMMA_SWIZZLE_64B);
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        const uint64_t da = umma_smem_desc(smem_u32(a_tile[sub]), 16, 512, UMMA_SWIZZLE_64B);
        umma_ss(tmem_base + (uint32_t)(sub * ST_COUT), da, db, idesc, 0);
        umma_ss(tmem_base + (uint32_t)(sub * ST_COUT), da + 2, db + 2, idesc, 1);
      }
      umma_commit(&bar);
    }
    // ---- under the MMAs: request the next tile's patch (the other buffer was last read one tile ago)
    if (tile + gridDim.x < n_tiles) issue_patch(tile + gridDim.x, pb ^ 1);
    cp_async_commit();
    mbar_wait(&bar, par);
    par ^= 1;
    tc_fence_after();
    // ---- epilogue: my pixel of each sub-tile (bias already added by the MMA): ReLU on packed bf16, staged as a
    //      64-byte row of the (now free) im2col tile, then one TMA box store per sub-tile: [2 rows][64 px][32 ch],
    //      full lines, clipped at the image border by the tensor map
    const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub) {
#pragma unroll
      for (int g = 0; g < NCH; ++g) {
        uint32_t acc[32];
        tmem_ld_x32(tmem_base + lane_base + (uint32_t)(sub * ST_COUT + g * 32), acc);
        tmem_ld_wait();
        uint8_t* stg = g == 0 ? a_tile[sub] : o_tile[sub];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __nv_bfloat162 h = __hmax2(__floats2bfloat162_rn(__uint_as_float(acc[8 * c + 2 * e]),
                                                                   __uint_as_float(acc[8 * c + 2 * e + 1])), zero2);
            w4[e] = *reinterpret_cast<const uint32_t*>(&h);
          }
          *reinterpret_cast<uint4*>(stg + tid * 64 + ((c ^ ((tid >> 1) & 3)) << 4)) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
      }
    }
    fence_async_shared();
    tc_fence_before();       // also orders these TMEM reads before the next tile's MMAs (via the barrier)
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int sub = 0; sub < NSUB; ++sub) {
        tma_store_4d(&map_out, a_tile[sub], 0, x0, y0 + 2 * sub, b);
        if (NCH == 2) tma_store_4d(&map_out, o_tile[sub], 32, x0, y0 + 2 * sub, b);
      }
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait<0>();
  cp_async_wait<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(4 * ST_COUT) : "memory");
  }
}

}  // namespace psw

using namespace psw;

template <int COUT>
static int launch_conv1(const float* img, const float* w_folded, const float* bias_folded, void* out, int B, int H, int W,
                        cudaStream_t st) {
  const int64_t tiles = (int64_t)B * ((H + ST_TH - 1) / ST_TH) * ((W + ST_TW - 1) / ST_TW);
  int64_t grid = (int64_t)num_sms() * (COUT == 32 ? 4 : 2);
  if (grid > tiles) grid = tiles;
  const size_t smem = 1024 + (size_t)(COUT / 32) * (ST_TH / 2) * 128 * 64 + COUT * 64 + 2 * ST_PATCH_FLOATS * sizeof(float) +
                      COUT * sizeof(float) + 16;
  auto kern = stem_conv1_kernel<COUT>;
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  // out as a 4-D tensor (channel, x, y, image); box = 32 channels x 64 pixels x 2 rows (one 128-pixel sub-tile)
  CUtensorMap map_out;
  const uint64_t dims[4] = {(uint64_t)COUT, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)COUT * 2, (uint64_t)W * COUT * 2, (uint64_t)H * W * COUT * 2};
  const uint32_t box[4] = {32u, (uint32_t)ST_TW, 2u, 1u};
  int rc = make_tensor_map_nd(&map_out, out, 4, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  kern<<<(unsigned)grid, ST_THREADS, smem, st>>>(img, w_folded, bias_folded, map_out, B, H, W);
  return launch_status("stem_conv1_kernel");
}

extern "C" PSW_API int psw_stem_conv3x3_relu_fwd(const float* img, const float* w_folded, const float* bias_folded,
                                                 void* out, int B, int H, int W, int cin, int cout, void* stream) {
  PSW_REQUIRE(img && w_folded && bias_folded && out, PSW_ERR_BAD_ARG, "psw_stem_conv3x3_relu_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0, PSW_ERR_BAD_ARG, "psw_stem_conv3x3_relu_fwd: bad dims");
  PSW_REQUIRE(cin == ST_CIN && (cout == 32 || cout == 64), PSW_ERR_UNSUPPORTED,
              "psw_stem_conv3x3_relu_fwd: built for 3 -> 32 or 64 channels (zero-pad other widths up to 64); got %d -> %d", cin, cout);
  PSW_REQUIRE(aligned16(out) && aligned16(img), PSW_ERR_BAD_ARG, "psw_stem_conv3x3_relu_fwd: img / out must be 16-byte aligned");
  PSW_REQUIRE(W % 4 == 0, PSW_ERR_UNSUPPORTED, "psw_stem_conv3x3_relu_fwd: W must be a multiple of 4 (the patch-size padding guarantees it)");
  if (cout == 64) return launch_conv1<64>(img, w_folded, bias_folded, out, B, H, W, (cudaStream_t)stream);
  return launch_conv1<32>(img, w_folded, bias_folded, out, B, H, W, (cudaStream_t)stream);
}
