// Probe: cp.async.bulk.tensor.2d.tile::gather4 -- tensor-map box shape it wants, order / swizzle of the 4 gathered rows
// in shared memory, out-of-range row index behaviour.
#include <cstdio>
#include <vector>
#include "../../panoswintransformerobjectdetection_b200/csrc/psw_common.cuh"
using namespace psw;

__global__ void probe(const __grid_constant__ CUtensorMap map, uint16_t* out, int r0, int r1, int r2, int r3, int col) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) smem[i] = 0xEE;
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, 256);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem + 256)), "l"(&map), "r"(smem_u32(&bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < 512; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
}

int main() {
  const int R = 64, C = 96;
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 128 + c);   // raw 16-bit tags
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 1024);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  for (int boxh = 1; boxh <= 4; boxh += 3) {
    CUtensorMap map;
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)R};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {32, (uint32_t)boxh};
    int rc = make_tensor_map_nd(&map, d, 2, dims, strides, box, 2, CU_TENSOR_MAP_SWIZZLE_64B);
    printf("box height %d: encode rc=%d\n", boxh, rc);
    if (rc) continue;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096);
    probe<<<1, 32, 4096>>>(map, o, 5, 17, 2, 70, 32);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  cuda error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<uint16_t> r(512);
    cudaMemcpy(r.data(), o, 1024, cudaMemcpyDeviceToHost);
    for (int row = 0; row < 8; ++row) {
      printf("  smem row %d (64 B):", row);
      for (int ch = 0; ch < 4; ++ch) printf("  [r%d c%d..]", r[row * 32 + ch * 8] >> 7, r[row * 32 + ch * 8] & 127);
      printf("\n");
    }
  }
  return 0;
}
