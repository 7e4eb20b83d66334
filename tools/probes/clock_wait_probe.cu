// Do clock64() differences taken around a suspending mbarrier.try_wait account for the time spent waiting?
// Warp 1 arrives on an mbarrier every ~WORK cycles; thread 0 of warp 0 times each wait with clock64() and sums the
// differences; the sum is compared with the clock64() and globaltimer span of the whole loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o clock_wait_probe clock_wait_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__global__ void probe(long long work, int iters, long long* out) {
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
  }
  __syncthreads();
  if (threadIdx.x == 32) {                       // producer: arrive on bar[0] every `work` cycles, wait for the consumer's ack
    for (int i = 0; i < iters; ++i) {
      const long long c = clock64();
      while (clock64() - c < work) {}
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[0])) : "memory");
      while (!try_wait(&bar[1], i & 1)) {}
    }
  } else if (threadIdx.x == 0) {
    long long sum = 0, t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long cs = clock64();
    for (int i = 0; i < iters; ++i) {
      const long long c0 = clock64();
      while (!try_wait(&bar[0], i & 1)) {}
      const long long c1 = clock64();
      sum += c1 - c0;
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar[1])) : "memory");
    }
    const long long ce = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    out[0] = sum; out[1] = ce - cs; out[2] = t1 - t0;
  }
}

int main() {
  long long* d; cudaMalloc(&d, 3 * sizeof(long long));
  for (long long work : {500LL, 2000LL, 8000LL}) {
    probe<<<1, 64>>>(work, 2000, d);
    long long h[3];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("work %5lld cycles x 2000: sum of timed waits %9lld, loop span %9lld cycles (waits = %.3f of the span), %.1f us -> %.3f GHz\n",
           work, h[0], h[1], (double)h[0] / h[1], h[2] / 1e3, (double)h[1] / h[2]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
