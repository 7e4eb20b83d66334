// Probe: does a tcgen05 SWIZZLE_64B K-major operand still read correctly when its start address is shifted by whole
// 64-byte rows (not a multiple of the 512-byte swizzle pattern)?  Data is laid out the way TMA writes it (16-byte
// chunk index ^= (absolute smem address >> 7) & 3).  Tries base_offset = 0 and base_offset = (start >> 7) & 7.
#include <cstdio>
#include <vector>
#include "../../panoswintransformerobjectdetection_b200/csrc/psw_common.cuh"
using namespace psw;

__global__ void probe(float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* A = smem;              // 160 rows x 64 B
  uint8_t* Bm = smem + 16384;     // 32 rows x 64 B
  __shared__ uint32_t tmem_slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 32; i += blockDim.x) {
    const int row = i / 32, k = i % 32;
    const uint32_t addr_row = smem_u32(A) + row * 64;
    const int chunk = (k / 8) ^ ((addr_row >> 7) & 3);
    reinterpret_cast<__nv_bfloat16*>(A + row * 64 + chunk * 16)[k % 8] = __float2bfloat16((float)((row * 7 + k * 3) % 13 - 6));
  }
  for (int i = tid; i < 32 * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    const uint32_t addr_row = smem_u32(Bm) + n * 64;
    const int chunk = (k / 8) ^ ((addr_row >> 7) & 3);
    reinterpret_cast<__nv_bfloat16*>(Bm + n * 64 + chunk * 16)[k % 8] = __float2bfloat16((float)((n * 5 + k * 11) % 7 - 3));
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_shared();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
  uint32_t phase = 0;
  for (int mode = 0; mode < 2; ++mode) {
    for (int s = 0; s < 10; ++s) {
      if (tid == 0) {
        const uint32_t sa = smem_u32(A) + s * 64;
        uint64_t da = umma_smem_desc(sa, 16, 512, UMMA_SWIZZLE_64B);
        if (mode == 1) da |= (uint64_t)((sa >> 7) & 7) << 49;
        const uint64_t db = umma_smem_desc(smem_u32(Bm), 16, 512, UMMA_SWIZZLE_64B);
        umma_ss(tmem, da, db, idesc, false);
        umma_ss(tmem, da + 2, db + 2, idesc, true);
        umma_commit(&bar);
      }
      mbar_wait(&bar, phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_x32(tmem + ((uint32_t)((tid >> 5) * 32) << 16), r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[((mode * 10 + s) * 128 + tid) * 32 + j] = __uint_as_float(r[j]);
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 20 * 128 * 32 * sizeof(float));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  probe<<<1, 128, 32768>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h(20 * 128 * 32);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  for (int mode = 0; mode < 2; ++mode)
    for (int s = 0; s < 10; ++s) {
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
          float want = 0;
          for (int k = 0; k < 32; ++k) want += (float)(((m + s) * 7 + k * 3) % 13 - 6) * (float)((n * 5 + k * 11) % 7 - 3);
          if (h[((mode * 10 + s) * 128 + m) * 32 + n] != want) ++bad;
        }
      printf("base_offset %s  shift %d rows: %s (%d mismatches)\n", mode ? "(start>>7)&7" : "0", s, bad ? "WRONG" : "ok", bad);
    }
  return 0;
}
