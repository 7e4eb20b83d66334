// How much HBM bandwidth does the attention kernel's access pattern sustain as a function of the contiguous segment
// size per request?  Memory skeleton only (cp.async gather into double-buffered shared memory, then 16-byte stores of
// the q part), 128 threads, 4 CTAs per SM, ~18.8 KB gathered + 6.3 KB stored per step -- the shape of
// window_attn_bi_kernel -- with the token rows split into 64-byte segments (one head per tile half, two images:
// today's kernel) or 128-byte segments (a head PAIR of one image per tile).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gather_seg_probe gather_seg_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void* d, const void* s) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d)), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int HW = 32768, WTOK = 256, C = 128, HEADS = 4, ROWB = 3 * C * 2;   // qkv row: 768 B
constexpr int NWIN = 36 * 18;                                                  // windows per image (7x7 tokens each)

__device__ __forceinline__ int token_of(int win, int t) {     // window `win`, cell t = r*7+c -> token index in the image
  const int wr = win / 18, wc = win - wr * 18;
  const int r = t / 7, c = t - r * 7;
  return (wr * 7 + r) * WTOK + wc * 7 + c;
}

// SEG = 64 : unit = (window, head, image pair); tile rows [img][49 tokens], each row 3 x 64 B
// SEG = 128: unit = (window, head pair, image);  tile rows [49 tokens],      each row 3 x 128 B
template <int SEG, int STAGES, bool COAL>
__global__ void __launch_bounds__(128, 4) probe(const uint8_t* __restrict__ qkv, uint8_t* __restrict__ out, int B) {
  extern __shared__ uint8_t smem[];
  constexpr int STAGE_BYTES = 3 * 128 * 64;                  // 24 KB
  const int tid = threadIdx.x;
  const int groups = SEG == 64 ? HEADS : HEADS / 2;          // head groups per window
  const int inner = SEG == 64 ? B / 2 : B;                   // steps per (window, head group)
  const int n_items = NWIN * groups;
  const int64_t n_steps = (int64_t)n_items * inner;
  // step index s -> (item, b): items dealt round-robin over CTAs with the head group fastest, 4 image steps per unit
  const int CH = 4;
  const int nch = inner / CH;
  const int64_t n_units = (int64_t)n_items * nch;
  auto issue = [&](int64_t u, int k, int stage) {
    const int g = (int)(u % groups);
    const int64_t wc = u / groups;
    const int chunk = (int)(wc % nch);
    const int win = (int)(wc / nch);
    const int bstep = chunk * CH + k;
    uint8_t* base = smem + stage * STAGE_BYTES;
    if (SEG == 64) {
      // 128 threads: chunk lc of tokens lt0, lt0+32 of both images (4 rows) x 3 parts
      const int lc = tid & 3, lt0 = tid >> 2;
      for (int kk = 0; kk < 4; ++kk) {
        const int t = lt0 + 32 * (kk & 1);
        if (t >= 49) continue;
        const int b = 2 * bstep + (kk >> 1);
        const uint8_t* src = qkv + ((int64_t)b * HW + token_of(win, t)) * ROWB + g * 64 + lc * 16;
        uint8_t* dst = base + ((kk >> 1) * 64 + t) * 64 + lc * 16;
        cp_async16(dst, src);
        cp_async16(dst + 8192, src + C * 2);
        cp_async16(dst + 16384, src + 2 * C * 2);
      }
    } else {
      // chunk lc (0..7) of tokens lt0 (0..15), +16, +32, +48 x 3 parts
      const int lc = tid & 7, lt0 = tid >> 3;
      for (int kk = 0; kk < 4; ++kk) {
        const int t = lt0 + 16 * kk;
        if (t >= 49) continue;
        const uint8_t* src = qkv + ((int64_t)bstep * HW + token_of(win, t)) * ROWB + g * 128 + lc * 16;
        uint8_t* dst = base + t * 128 + lc * 16;
        cp_async16(dst, src);
        cp_async16(dst + 8192, src + C * 2);
        cp_async16(dst + 16384, src + 2 * C * 2);
      }
    }
  };
  auto store = [&](int64_t u, int k, int stage) {
    const int g = (int)(u % groups);
    const int64_t wc = u / groups;
    const int chunk = (int)(wc % nch);
    const int win = (int)(wc / nch);
    const int bstep = chunk * CH + k;
    const uint8_t* base = smem + stage * STAGE_BYTES;
    if (SEG == 64 && COAL) {
      // four lanes per 64-byte row: a warp-wide 16-byte store covers 8 rows (8 lines) instead of 32
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = 0; i < 4; ++i) {
        const int row = warp * 32 + (lane >> 2) + 8 * i;
        const int half = row >> 6, t = row & 63;
        if (t < 49) {
          const int b = 2 * bstep + half;
          uint4* dst = reinterpret_cast<uint4*>(out + ((int64_t)b * HW + token_of(win, t)) * (C * 2) + g * 64) + (lane & 3);
          *dst = *reinterpret_cast<const uint4*>(base + row * 64 + (lane & 3) * 16);
        }
      }
    } else if (SEG == 64) {
      const int half = tid >> 6, t = tid & 63;
      if (t < 49) {
        const int b = 2 * bstep + half;
        uint4* dst = reinterpret_cast<uint4*>(out + ((int64_t)b * HW + token_of(win, t)) * (C * 2) + g * 64);
        const uint4* s4 = reinterpret_cast<const uint4*>(base + (half * 64 + t) * 64);
        for (int c = 0; c < 4; ++c) dst[c] = s4[c];
      }
    } else {
      const int half = tid >> 6, t = tid & 63;                 // thread = (head of the pair, token): 64 B each
      if (t < 49) {
        uint4* dst = reinterpret_cast<uint4*>(out + ((int64_t)bstep * HW + token_of(win, t)) * (C * 2) + g * 128 + half * 64);
        const uint4* s4 = reinterpret_cast<const uint4*>(base + t * 128 + half * 64);
        for (int c = 0; c < 4; ++c) dst[c] = s4[c];
      }
    }
  };
  // flat iteration over (unit, k): this CTA's units are blockIdx.x, + gridDim.x, ...
  int64_t u = blockIdx.x;
  int k = 0, n = 0;
  // prologue: STAGES - 1 steps in flight
  int64_t pu = u; int pk = 0;
  for (int s = 0; s < STAGES - 1; ++s) {
    if (pu < n_units) issue(pu, pk, s);
    cp_commit();
    if (++pk == CH) { pk = 0; pu += gridDim.x; }
  }
  while (u < n_units) {
    if (pu < n_units) issue(pu, pk, (n + STAGES - 1) % STAGES);
    cp_commit();
    if (++pk == CH) { pk = 0; pu += gridDim.x; }
    cp_wait<STAGES - 1>();
    __syncthreads();
    store(u, k, n % STAGES);
    __syncthreads();
    ++n;
    if (++k == CH) { k = 0; u += gridDim.x; }
  }
  (void)n_steps;
}

template <int SEG, int STAGES, bool COAL = false>
static void run(const uint8_t* qkv, uint8_t* out, int B, int ctas_per_sm) {
  const int smem = STAGES * 3 * 128 * 64 + (ctas_per_sm == 4 ? (STAGES == 2 ? 4096 : 0) : 0);
  cudaFuncSetAttribute(probe<SEG, STAGES, COAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<SEG, STAGES, COAL>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<SEG, STAGES, COAL>, 128, smem);
  const int grid = 148 * occ;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) probe<SEG, STAGES, COAL><<<grid, 128, smem>>>(qkv, out, B);
  cudaEventRecord(e0);
  const int iters = 20;
  for (int i = 0; i < iters; ++i) probe<SEG, STAGES, COAL><<<grid, 128, smem>>>(qkv, out, B);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  const double tok = (double)B * NWIN * 49;                         // tokens touched
  const double bytes = tok * C * 2 * 4;                             // q, k, v read + q-sized write, all heads
  printf("SEG %3d stages %d coalesced-stores %d occupancy %d: %8.1f us  %7.0f GB/s  (%s)\n", SEG, STAGES, (int)COAL, occ, ms / iters * 1e3,
         bytes / (ms / iters * 1e-3) / 1e9, cudaGetErrorString(err));
}

int main() {
  const int B = 32;
  uint8_t *qkv, *out;
  cudaMalloc(&qkv, (size_t)B * HW * ROWB);
  cudaMalloc(&out, (size_t)B * HW * C * 2);
  cudaMemset(qkv, 1, (size_t)B * HW * ROWB);
  run<64, 2>(qkv, out, B, 4);
  run<64, 2, true>(qkv, out, B, 4);
  run<128, 2>(qkv, out, B, 4);
  run<64, 3>(qkv, out, B, 3);
  run<128, 3>(qkv, out, B, 3);
  run<64, 4>(qkv, out, B, 2);
  run<128, 4>(qkv, out, B, 2);
  return 0;
}
