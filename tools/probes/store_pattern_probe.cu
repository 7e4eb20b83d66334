// Write bandwidth of 16-byte global stores as a function of how a warp's 32 lanes map onto the rows of a row-major
// [M, N] bf16 matrix (the GEMM epilogues' store pattern): SEG = contiguous bytes per row per warp instruction.
//   SEG  64: 4 lanes per row, 8 rows per instruction   (today's epilogue: one 32-column bf16 chunk)
//   SEG 128: 8 lanes per row, 4 rows per instruction
//   SEG 256: 16 lanes per row, 2 rows
//   SEG 512: the whole warp on one row
// Each warp owns a 32-row x 384-byte panel (N_TILE = 192 columns) and walks over it chunk by chunk like the epilogue
// does; 16 warps per CTA, one CTA per SM, persistent over the row panels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o store_pattern_probe store_pattern_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int SEG>
__global__ void __launch_bounds__(512, 1) probe(uint8_t* __restrict__ y, int64_t M, int N, int tile_n) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int LPR = SEG / 16;                 // lanes per row
  constexpr int RPI = 32 / LPR;                 // rows per instruction
  const int64_t row_bytes = (int64_t)N * 2;
  const int n_tiles = N / tile_n;
  const int64_t panels = (M / 32) * n_tiles;    // 32-row x tile_n panels
  const uint4 v = make_uint4(lane, warp, blockIdx.x, 7);
  for (int64_t p = (int64_t)blockIdx.x * 16 + warp; p < panels; p += (int64_t)gridDim.x * 16) {
    const int64_t m_t = p / n_tiles;
    const int n_t = (int)(p - m_t * n_tiles);
    uint8_t* base = y + m_t * 32 * row_bytes + (int64_t)n_t * tile_n * 2;
    for (int c = 0; c < tile_n * 2; c += SEG) {              // column chunk
#pragma unroll
      for (int j = 0; j < 32 / RPI; ++j) {
        const int row = (lane / LPR) + RPI * j;
        *reinterpret_cast<uint4*>(base + row * row_bytes + c + (lane % LPR) * 16) = v;
      }
    }
  }
}

template <int SEG>
static void run(uint8_t* y, int64_t M, int N, int tile_n) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) probe<SEG><<<148, 512>>>(y, M, N, tile_n);
  cudaEventRecord(e0);
  const int iters = 10;
  for (int i = 0; i < iters; ++i) probe<SEG><<<148, 512>>>(y, M, N, tile_n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("M %lld N %d tile_n %d SEG %3d: %8.1f us  %6.0f GB/s  (%s)\n", (long long)M, N, tile_n, SEG, ms / iters * 1e3,
         (double)M * N * 2 / (ms / iters * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  uint8_t* y;
  const int64_t Mmax = 1 << 20;
  cudaMalloc(&y, (size_t)Mmax * 1536 * 2);
  struct { int64_t M; int N, tn; } cases[] = {{1 << 20, 288, 96}, {1 << 20, 288, 288}, {262144, 576, 192}, {262144, 768, 192}, {65536, 1536, 192}, {65536, 1536, 256}};
  for (auto c : cases) {
    run<64>(y, c.M, c.N, c.tn);
    run<128>(y, c.M, c.N, c.tn);
    if (c.tn * 2 % 256 == 0) run<256>(y, c.M, c.N, c.tn);
    if (c.tn * 2 % 512 == 0) run<512>(y, c.M, c.N, c.tn);
  }
  return 0;
}
