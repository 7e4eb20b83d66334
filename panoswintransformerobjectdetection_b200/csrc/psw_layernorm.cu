// K1 / K4: vectorised bandwidth kernels — LayerNorm (+ optional position add), PatchMerging
// gather + LayerNorm(4C), and the per-stage output LayerNorm + NHWC->NCHW transpose.
//
// One warp owns one row; the row lives in registers (VPL 4-element vectors per lane), so every input
// byte is read from HBM exactly once and statistics are two-pass (mean, then centred variance) in fp32
// like torch.nn.functional.layer_norm.  Algorithmic bytes per row: C*(sizeof(in)+sizeof(out)).
#include "psw_common.cuh"

namespace psw {

constexpr int LN_WARPS = 8;

// Row gather policies -----------------------------------------------------------------------------
// A policy splits the address of 4-vector `vec` of row `r` into a per-vector part (loop invariant for a lane: computed
// once per thread) and a per-row part (computed once per row), so the inner loop is a single add / compare.
struct PlainRows {
  int64_t rows;
  int C;
  struct Vec { int off; };
  struct Row { int64_t base; };
  __device__ __forceinline__ int64_t num_rows() const { return rows; }
  __device__ __forceinline__ Vec vec_info(int vec) const { return {vec * 4}; }
  __device__ __forceinline__ Row row_info(int64_t r) const { return {r * C}; }
  // element offset of the vector, or -1 for an implicit zero vector
  __device__ __forceinline__ int64_t offset(const Row& r, const Vec& v) const { return r.base + v.off; }
};

// PatchMerging (reference :563-573): output row (b, i2, j2) = concat of x[b, 2*i2+dh, 2*j2+dw, :] for
// quadrant q = 0..3 with dh = q & 1, dw = q >> 1; cells beyond an odd H / W are zero.
struct MergeRows {
  int B, H, W, C, H2, W2;
  struct Vec { int dh, dw, off; };                         // quadrant of the vector and its offset inside the token
  struct Row { int64_t base; int h0, w0; };                // token (2*i2, 2*j2) of the merged row
  __device__ __forceinline__ int64_t num_rows() const { return (int64_t)B * H2 * W2; }
  __device__ __forceinline__ Vec vec_info(int vec) const {
    const int vpc = C >> 2;
    const int q = vec / vpc;
    return {q & 1, q >> 1, (vec - q * vpc) * 4};
  }
  __device__ __forceinline__ Row row_info(int64_t r) const {       // merged rows < 2^31 (checked by the host)
    const uint32_t ru = (uint32_t)r;
    const uint32_t t = ru / (uint32_t)W2;
    const int j2 = (int)(ru - t * (uint32_t)W2);
    const uint32_t b = t / (uint32_t)H2;
    const int i2 = (int)(t - b * (uint32_t)H2);
    return {(((int64_t)b * H + 2 * i2) * W + 2 * j2) * C, 2 * i2, 2 * j2};
  }
  __device__ __forceinline__ int64_t offset(const Row& r, const Vec& v) const {
    if (r.h0 + v.dh >= H || r.w0 + v.dw >= W) return -1;
    return r.base + ((int64_t)v.dh * W + v.dw) * C + v.off;
  }
};

// Optional second LayerNorm of the rows just produced (psw_layernorm2_fwd): y2 = LN(y) * gamma2 + beta2 in bf16 while the
// row is still in registers -- the stem's patch_norm (+ position add) followed by the first block's norm1.
struct SecondNorm {
  bf16* y2;
  const float* gamma2;
  const float* beta2;
  float eps2;
};

// LPR lanes cooperate on one row (LPR = 8, 16 or 32), each lane owns VPL 4-element vectors at vector indices
// lane + LPR * k, so a group reads LPR * 16 B contiguous per request and every lane of the warp is busy even for
// C = 96 (24 vectors: 8 lanes x 3).  A warp works on (32 / LPR) * U rows per iteration; all loads are issued before
// the reductions (memory-level parallelism), statistics are two-pass in fp32 from registers.
// POS: the position rows are fetched together with the input rows (before the reductions) instead of at the point of
// use -- they come from L2, and with their latency exposed in the output loop the stem LayerNorm ran at half speed.
template <int LPR, int VPL, int U, typename TI, typename TO, typename Rows, bool POS = false>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_rows_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ pos, int64_t pos_rows,
                      Rows rows, int Cout, float eps, const SecondNorm sn) {
  constexpr int GROUPS = 32 / LPR;                 // rows per warp pass
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane % LPR;                      // lane inside the row group
  const int grp = lane / LPR;
  const int nvec = Cout >> 2;
  const int64_t nrows = rows.num_rows();
  const float inv_c = 1.0f / (float)Cout;
  const int64_t rows_per_iter = (int64_t)GROUPS * U;
  const int64_t stride = (int64_t)gridDim.x * LN_WARPS * rows_per_iter;
  typename Rows::Vec vinfo[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) vinfo[k] = rows.vec_info(sub + LPR * k);
  for (int64_t r0 = ((int64_t)blockIdx.x * LN_WARPS + warp) * rows_per_iter; r0 < nrows; r0 += stride) {
    float v[U][VPL][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * GROUPS + grp;
      const typename Rows::Row rinfo = rows.row_info(r < nrows ? r : 0);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sub + LPR * k;
        v[u][k][0] = v[u][k][1] = v[u][k][2] = v[u][k][3] = 0.f;
        if (vec < nvec && r < nrows) {
          const int64_t off = rows.offset(rinfo, vinfo[k]);
          if (off >= 0) load4(x + off, v[u][k]);
        }
      }
    }
    const float* prow[U];                                  // position row of each of my rows: one division per row, 32-bit
#pragma unroll                                             // where it fits (a 64-bit modulo per vector cost more than the loads)
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u * GROUPS + grp;
      prow[u] = nullptr;
      if (pos) {
        const int64_t pr = ((r | pos_rows) >> 32) == 0 ? (int64_t)((uint32_t)r % (uint32_t)pos_rows) : r % pos_rows;
        prow[u] = pos + pr * Cout;
      }
    }
    float pp[POS ? U : 1][POS ? VPL : 1][4];
    if constexpr (POS) {
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          pp[u][k][0] = pp[u][k][1] = pp[u][k][2] = pp[u][k][3] = 0.f;
          if (sub + LPR * k < nvec && r0 + u * GROUPS + grp < nrows) load4(prow[u] + (sub + LPR * k) * 4, pp[u][k]);
        }
    }
    float mean[U], rstd[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) s += (v[u][k][0] + v[u][k][1]) + (v[u][k][2] + v[u][k][3]);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      mean[u] = s * inv_c;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (sub + LPR * k < nvec) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = v[u][k][e] - mean[u];
            q += d * d;
          }
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      rstd[u] = rsqrtf(q * inv_c + eps);
    }
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vec = sub + LPR * k;
      if (vec < nvec) {
        float g[4], b[4];
        load4(gamma + vec * 4, g);
        load4(beta + vec * 4, b);
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = r0 + u * GROUPS + grp;
          if (r < nrows) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (v[u][k][e] - mean[u]) * rstd[u] * g[e] + b[e];
            if constexpr (POS) {
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] += pp[u][k][e];
            } else if (pos) {
              float pl[4];
              load4(prow[u] + vec * 4, pl);
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] += pl[e];
            }
            store4(y + r * Cout + (int64_t)vec * 4, o);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[u][k][e] = o[e];   // kept for the optional second norm
          }
        }
      }
    }
    if (sn.y2 != nullptr) {                                  // second LayerNorm on the values just written (two-pass again)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k)
          if (sub + LPR * k < nvec) s += (v[u][k][0] + v[u][k][1]) + (v[u][k][2] + v[u][k][3]);
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        mean[u] = s * inv_c;
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          if (sub + LPR * k < nvec) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float d = v[u][k][e] - mean[u];
              q += d * d;
            }
          }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        rstd[u] = rsqrtf(q * inv_c + sn.eps2);
      }
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sub + LPR * k;
        if (vec < nvec) {
          float g[4], b[4];
          load4(sn.gamma2 + vec * 4, g);
          load4(sn.beta2 + vec * 4, b);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t r = r0 + u * GROUPS + grp;
            if (r < nrows) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = (v[u][k][e] - mean[u]) * rstd[u] * g[e] + b[e];
              store4(sn.y2 + r * Cout + (int64_t)vec * 4, o);
            }
          }
        }
      }
    }
  }
}

template <typename TI, typename TO, typename Rows>
static int launch_ln(const TI* x, TO* y, const float* gamma, const float* beta, const float* pos, int64_t pos_rows,
                     Rows rows, int64_t nrows, int Cout, float eps, cudaStream_t st, SecondNorm sn = {nullptr, nullptr, nullptr, 0.f}) {
  const int nvec = Cout / 4;
  // lanes per row: the smallest of 8 / 16 / 32 that keeps at most 4 vectors per lane; wider rows use 32 lanes
  int lpr = 8;
  while (lpr < 32 && (nvec + lpr - 1) / lpr > 4) lpr <<= 1;
  const int vpl = (nvec + lpr - 1) / lpr;
  static const int kInst[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
  int inst = 32;
  for (int k = 9; k >= 0; --k)
    if (kInst[k] >= vpl) inst = kInst[k];
  const int u = inst <= 4 ? 2 : 1;                               // rows in flight per lane group
  const int64_t rows_per_block = (int64_t)LN_WARPS * (32 / lpr) * u;
  int64_t want = (nrows + rows_per_block - 1) / rows_per_block;
  int64_t cap = (int64_t)num_sms() * 16;
  int blocks = (int)(want < cap ? want : cap);
  if (blocks < 1) blocks = 1;
#define PSW_LN_LAUNCH(L, V, UU)                                                                                      \
  do {                                                                                                               \
    if constexpr (V * UU <= 8 && sizeof(TO) == 4) {                                                                  \
      if (pos) {                                                                                                     \
        layernorm_rows_kernel<L, V, UU, TI, TO, Rows, true><<<blocks, LN_WARPS * 32, 0, st>>>(x, y, gamma, beta, pos, \
                                                                                             pos_rows, rows, Cout, eps, sn); \
        break;                                                                                                       \
      }                                                                                                              \
    }                                                                                                                \
    layernorm_rows_kernel<L, V, UU, TI, TO, Rows><<<blocks, LN_WARPS * 32, 0, st>>>(x, y, gamma, beta, pos, pos_rows, \
                                                                                   rows, Cout, eps, sn);            \
  } while (0)
  if (lpr == 8) {
    switch (inst) { case 1: PSW_LN_LAUNCH(8, 1, 2); break; case 2: PSW_LN_LAUNCH(8, 2, 2); break;
                    case 3: PSW_LN_LAUNCH(8, 3, 2); break; default: PSW_LN_LAUNCH(8, 4, 2); break; }
  } else if (lpr == 16) {
    switch (inst) { case 1: PSW_LN_LAUNCH(16, 1, 2); break; case 2: PSW_LN_LAUNCH(16, 2, 2); break;
                    case 3: PSW_LN_LAUNCH(16, 3, 2); break; default: PSW_LN_LAUNCH(16, 4, 2); break; }
  } else {
    switch (inst) {
      case 1: PSW_LN_LAUNCH(32, 1, 2); break; case 2: PSW_LN_LAUNCH(32, 2, 2); break; case 3: PSW_LN_LAUNCH(32, 3, 2); break;
      case 4: PSW_LN_LAUNCH(32, 4, 2); break; case 6: PSW_LN_LAUNCH(32, 6, 1); break; case 8: PSW_LN_LAUNCH(32, 8, 1); break;
      case 12: PSW_LN_LAUNCH(32, 12, 1); break; case 16: PSW_LN_LAUNCH(32, 16, 1); break;
      case 24: PSW_LN_LAUNCH(32, 24, 1); break; default: PSW_LN_LAUNCH(32, 32, 1); break;
    }
  }
#undef PSW_LN_LAUNCH
  (void)u;
  return launch_status("layernorm_rows_kernel");
}

template <typename Rows>
static int dispatch_ln(const void* x, void* y, const float* gamma, const float* beta, const float* pos,
                       int64_t pos_rows, Rows rows, int64_t nrows, int Cout, float eps, int in_dtype, int out_dtype,
                       cudaStream_t st) {
  if (in_dtype == PSW_F32 && out_dtype == PSW_F32)
    return launch_ln((const float*)x, (float*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_F32 && out_dtype == PSW_BF16)
    return launch_ln((const float*)x, (bf16*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_BF16 && out_dtype == PSW_BF16)
    return launch_ln((const bf16*)x, (bf16*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  if (in_dtype == PSW_BF16 && out_dtype == PSW_F32)
    return launch_ln((const bf16*)x, (float*)y, gamma, beta, pos, pos_rows, rows, nrows, Cout, eps, st);
  set_error("layernorm: unknown dtype combination %d -> %d", in_dtype, out_dtype);
  return PSW_ERR_BAD_ARG;
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm + NHWC -> NCHW (fp32 out).  A CTA owns TT consecutive tokens of one image.  Phase 1 is the row kernel
// above (LPR lanes per row, 16-byte loads, two-pass fp32 statistics) writing the normalised row into a token-major
// shared tile [TT][C + 1] (odd pitch: the scalar stores of phase 1 and the column reads of phase 2 are both
// bank-conflict-free).  Phase 2: every warp streams whole channels out, TT * 4 contiguous bytes per channel.
// ---------------------------------------------------------------------------------------------------
// SWZ (C % 32 == 0): the tile is [TT][C] with the 16-byte chunks of row r XOR-swizzled by (r & 7) instead: phase 1
// writes one 16-byte store per vector, phase 2 gives every lane ONE token -- a 16-byte shared load fetches four channels
// of it and each of the four scalar stores of a warp is a 128-byte run of one channel.  A third of the instructions of
// the scalar tile (the kernel was issue-bound: 69% issue-slot utilisation at 3 CTAs per SM).
template <int LPR, int VPL, typename TI, bool SWZ>
__global__ void __launch_bounds__(256)
layernorm_nchw_kernel(const TI* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int64_t HW, int C, int TT, float eps) {
  extern __shared__ __align__(16) float tile[];   // [TT][C + 1], or [TT][C] swizzled
  constexpr int GROUPS = 32 / LPR;
  const int P = SWZ ? C : C + 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane % LPR, grp = lane / LPR;
  const int nvec = C >> 2;
  const int64_t tiles_per_img = (HW + TT - 1) / TT;
  const int b = (int)(blockIdx.x / tiles_per_img);
  const int64_t t0 = (blockIdx.x % tiles_per_img) * TT;
  const float inv_c = 1.0f / (float)C;
  constexpr int U = VPL <= 3 ? 4 : 2;              // rows in flight per lane group (memory-level parallelism)
  for (int rb = warp * GROUPS + grp; rb < TT; rb += 8 * GROUPS * U) {
    float v[U][VPL][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = rb + u * 8 * GROUPS;
      const int64_t t = t0 + rr;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int vec = sub + LPR * k;
        v[u][k][0] = v[u][k][1] = v[u][k][2] = v[u][k][3] = 0.f;
        if (vec < nvec && rr < TT && t < HW) load4(x + ((int64_t)b * HW + t) * C + vec * 4, v[u][k]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int rr = rb + u * 8 * GROUPS;
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) s += (v[u][k][0] + v[u][k][1]) + (v[u][k][2] + v[u][k][3]);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * inv_c;
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        if (sub + LPR * k < nvec) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float d = v[u][k][e] - mean;
            q += d * d;
          }
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
      const float rstd = rsqrtf(q * inv_c + eps);
      if (rr < TT) {
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int vec = sub + LPR * k;
          if (vec < nvec) {
            float g[4], bt[4];
            load4(gamma + vec * 4, g);
            load4(beta + vec * 4, bt);
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = (v[u][k][e] - mean) * rstd * g[e] + bt[e];
            if constexpr (SWZ) {
              *reinterpret_cast<float4*>(tile + rr * P + ((vec ^ (rr & 7)) << 2)) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) tile[rr * P + vec * 4 + e] = o[e];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if constexpr (SWZ) {
    for (int tb = 0; tb < TT; tb += 32) {
      const int row = tb + lane;
      const bool ok = t0 + row < HW;
      const float* trow = tile + row * P;
      float* dst0 = y + (int64_t)b * C * HW + t0 + row;
#pragma unroll 4
      for (int q = warp; q < nvec; q += 8) {
        const float4 val = *reinterpret_cast<const float4*>(trow + ((q ^ (row & 7)) << 2));
        float* dst = dst0 + (int64_t)(4 * q) * HW;
        if (ok) {
          dst[0] = val.x;
          dst[HW] = val.y;
          dst[2 * HW] = val.z;
          dst[3 * HW] = val.w;
        }
      }
    }
    return;
  }
  for (int c = warp; c < C; c += 8) {
    float* dst = y + ((int64_t)b * C + c) * HW + t0;
    for (int seg = lane; seg < TT; seg += 32)
      if (t0 + seg < HW) dst[seg] = tile[seg * P + c];
  }
}

template <typename TI>
static int launch_nchw(const TI* x, float* y, const float* gamma, const float* beta, int B, int64_t HW, int C, float eps,
                       cudaStream_t st) {
  PSW_REQUIRE(C % 4 == 0, PSW_ERR_UNSUPPORTED, "psw_layernorm_nchw_fwd: C=%d must be a multiple of 4", C);
  const int nvec = C / 4;
  int lpr = 8;
  while (lpr < 32 && (nvec + lpr - 1) / lpr > 4) lpr <<= 1;
  const int vpl = (nvec + lpr - 1) / lpr;
  // tokens per tile: as wide as ~56 KiB of shared memory allows, 32..128
  int TT = 128;
  while (TT > 32 && (size_t)TT * (C + 1) * sizeof(float) > 56 * 1024) TT >>= 1;
  const bool swz = C % 32 == 0;
  const size_t smem = (size_t)TT * (swz ? C : C + 1) * sizeof(float);
  PSW_REQUIRE(smem <= 200 * 1024 && vpl <= 8, PSW_ERR_UNSUPPORTED, "psw_layernorm_nchw_fwd: C=%d too large", C);
  const int64_t blocks = (int64_t)B * ((HW + TT - 1) / TT);
  PSW_REQUIRE(blocks < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_layernorm_nchw_fwd: too many tiles");
#define PSW_NCHW_LAUNCH(L, V)                                                                                    \
  do {                                                                                                           \
    auto kern = swz ? layernorm_nchw_kernel<L, V, TI, true> : layernorm_nchw_kernel<L, V, TI, false>;            \
    PSW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
    kern<<<(unsigned)blocks, 256, smem, st>>>(x, y, gamma, beta, HW, C, TT, eps);                               \
  } while (0)
  if (lpr == 8) {
    switch (vpl) { case 1: PSW_NCHW_LAUNCH(8, 1); break; case 2: PSW_NCHW_LAUNCH(8, 2); break;
                   case 3: PSW_NCHW_LAUNCH(8, 3); break; default: PSW_NCHW_LAUNCH(8, 4); break; }
  } else if (lpr == 16) {
    switch (vpl) { case 1: PSW_NCHW_LAUNCH(16, 1); break; case 2: PSW_NCHW_LAUNCH(16, 2); break;
                   case 3: PSW_NCHW_LAUNCH(16, 3); break; default: PSW_NCHW_LAUNCH(16, 4); break; }
  } else {
    switch (vpl) { case 1: PSW_NCHW_LAUNCH(32, 1); break; case 2: PSW_NCHW_LAUNCH(32, 2); break;
                   case 3: PSW_NCHW_LAUNCH(32, 3); break; case 4: PSW_NCHW_LAUNCH(32, 4); break;
                   case 5: case 6: PSW_NCHW_LAUNCH(32, 6); break; default: PSW_NCHW_LAUNCH(32, 8); break; }
  }
#undef PSW_NCHW_LAUNCH
  return launch_status("layernorm_nchw_kernel");
}

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_layernorm_fwd(const void* x, void* y, const float* gamma, const float* beta, const float* pos,
                                 int64_t rows, int C, int64_t pos_rows, float eps, int in_dtype, int out_dtype,
                                 void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: null pointer");
  PSW_REQUIRE(rows > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: rows=%lld C=%d", (long long)rows, C);
  PSW_REQUIRE(C % 4 == 0 && C <= 4096, PSW_ERR_UNSUPPORTED, "psw_layernorm_fwd: C=%d must be a multiple of 4, <= 4096", C);
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta) && aligned16(pos), PSW_ERR_BAD_ARG,
              "psw_layernorm_fwd: pointers must be 16-byte aligned");
  PSW_REQUIRE(pos == nullptr || pos_rows > 0, PSW_ERR_BAD_ARG, "psw_layernorm_fwd: pos given but pos_rows <= 0");
  PlainRows pr{rows, C};
  return dispatch_ln(x, y, gamma, beta, pos, pos_rows > 0 ? pos_rows : 1, pr, rows, C, eps, in_dtype, out_dtype,
                     (cudaStream_t)stream);
}

extern "C" PSW_API int psw_layernorm2_fwd(const void* x, float* y, const float* gamma, const float* beta, const float* pos,
                                          void* y2, const float* gamma2, const float* beta2, int64_t rows, int C,
                                          int64_t pos_rows, float eps, float eps2, int in_dtype, void* stream) {
  PSW_REQUIRE(x && y && gamma && beta && y2 && gamma2 && beta2, PSW_ERR_BAD_ARG, "psw_layernorm2_fwd: null pointer");
  PSW_REQUIRE(rows > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_layernorm2_fwd: rows=%lld C=%d", (long long)rows, C);
  PSW_REQUIRE(C % 4 == 0 && C <= 4096, PSW_ERR_UNSUPPORTED, "psw_layernorm2_fwd: C=%d must be a multiple of 4, <= 4096", C);
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(y2) && aligned16(gamma) && aligned16(beta) && aligned16(gamma2) &&
                  aligned16(beta2) && aligned16(pos), PSW_ERR_BAD_ARG, "psw_layernorm2_fwd: pointers must be 16-byte aligned");
  PSW_REQUIRE(pos == nullptr || pos_rows > 0, PSW_ERR_BAD_ARG, "psw_layernorm2_fwd: pos given but pos_rows <= 0");
  PlainRows pr{rows, C};
  const SecondNorm sn = {(bf16*)y2, gamma2, beta2, eps2};
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == PSW_F32)
    return launch_ln((const float*)x, y, gamma, beta, pos, pos_rows > 0 ? pos_rows : 1, pr, rows, C, eps, st, sn);
  if (in_dtype == PSW_BF16)
    return launch_ln((const bf16*)x, y, gamma, beta, pos, pos_rows > 0 ? pos_rows : 1, pr, rows, C, eps, st, sn);
  PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_layernorm2_fwd: unknown dtype %d", in_dtype);
  return PSW_ERR_BAD_ARG;
}

extern "C" PSW_API int psw_patch_merge_ln_fwd(const void* x, void* y, const float* gamma, const float* beta, int B, int H,
                                      int W, int C, float eps, int in_dtype, int out_dtype, void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_fwd: null pointer");
  PSW_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_patch_merge_ln_fwd: bad dims");
  PSW_REQUIRE(C % 4 == 0 && 4 * C <= 4096, PSW_ERR_UNSUPPORTED, "psw_patch_merge_ln_fwd: C=%d must be a multiple of 4, <= 1024", C);
  PSW_REQUIRE(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), PSW_ERR_BAD_ARG,
              "psw_patch_merge_ln_fwd: pointers must be 16-byte aligned");
  MergeRows mr{B, H, W, C, (H + 1) / 2, (W + 1) / 2};
  int64_t nrows = (int64_t)B * mr.H2 * mr.W2;
  PSW_REQUIRE(nrows < (1ll << 31), PSW_ERR_UNSUPPORTED, "psw_patch_merge_ln_fwd: too many merged rows");
  return dispatch_ln(x, y, gamma, beta, nullptr, 1, mr, nrows, 4 * C, eps, in_dtype, out_dtype, (cudaStream_t)stream);
}

extern "C" PSW_API int psw_layernorm_nchw_fwd(const void* x, float* y, const float* gamma, const float* beta, int B,
                                      int64_t HW, int C, float eps, int in_dtype, void* stream) {
  PSW_REQUIRE(x && y && gamma && beta, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: null pointer");
  PSW_REQUIRE(B > 0 && HW > 0 && C > 0, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: bad dims");
  cudaStream_t st = (cudaStream_t)stream;
  if (in_dtype == PSW_F32) return launch_nchw((const float*)x, y, gamma, beta, B, HW, C, eps, st);
  if (in_dtype == PSW_BF16) return launch_nchw((const bf16*)x, y, gamma, beta, B, HW, C, eps, st);
  PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_layernorm_nchw_fwd: unknown dtype %d", in_dtype);
  return 0;
}
