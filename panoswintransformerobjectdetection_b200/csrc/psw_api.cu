// libpanoswin_b200: ABI bookkeeping, error strings, device check, tensor-map encoding, dtype cast.
#include <stdarg.h>
#include <string.h>

#include "psw_common.cuh"

namespace psw {

static thread_local char g_err[512] = "no error";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

EncodeTiledFn get_encode_tiled() {
  // function-local static: initialised exactly once, thread-safe (C++11 magic static)
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return reinterpret_cast<EncodeTiledFn>(p);
    return static_cast<EncodeTiledFn>(nullptr);
  }();
  return fn;
}

// SM count of the CURRENT device (queried per call: a process may drive several devices)
int num_sms() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;
  return n;
}

int make_tensor_map_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                       uint32_t box_cols, int elem_bytes, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_tiled();
  PSW_REQUIRE(enc != nullptr, PSW_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[2] = {cols, rows};                       // innermost first
  cuuint64_t strides[1] = {cols * (uint64_t)elem_bytes};   // bytes, dims 1..rank-1
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PSW_REQUIRE(r == CUDA_SUCCESS, PSW_ERR_DRIVER,
              "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
  return 0;
}

int make_tensor_map_nd(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, int elem_bytes, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_tiled();
  PSW_REQUIRE(enc != nullptr, PSW_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  PSW_REQUIRE(rank >= 1 && rank <= 5, PSW_ERR_BAD_ARG, "tensor map rank %d", rank);
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];      // strides of dims 1..rank-1
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PSW_REQUIRE(r == CUDA_SUCCESS, PSW_ERR_DRIVER, "cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r);
  return 0;
}

template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n4) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float v[4];
    load4(src + i * 4, v);
    store4(dst + i * 4, v);
  }
}
template <typename S, typename D>
__global__ void cast_tail_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t begin, int64_t n) {
  int64_t i = begin + threadIdx.x;
  if (i < n) dst[i] = from_f32<D>(to_f32(src[i]));
}

}  // namespace psw

using namespace psw;

extern "C" PSW_API int psw_abi_version(void) { return PSW_ABI_VERSION; }
extern "C" PSW_API const char* psw_last_error_string(void) { return g_err; }

extern "C" PSW_API int psw_check_device(int dev) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  PSW_REQUIRE(e == cudaSuccess && n > 0, PSW_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
  PSW_REQUIRE(dev >= 0 && dev < n, PSW_ERR_BAD_ARG, "device %d out of range (%d devices)", dev, n);
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  PSW_REQUIRE(major == 10, PSW_ERR_NO_DEVICE, "device %d is sm_%d%d; libpanoswin_b200 is built for sm_100a only",
              dev, major, minor);
  return 0;
}

extern "C" PSW_API int psw_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  PSW_REQUIRE(src && dst && n > 0, PSW_ERR_BAD_ARG, "psw_cast: null pointer or n <= 0");
  PSW_REQUIRE(aligned16(src) && aligned16(dst), PSW_ERR_BAD_ARG, "psw_cast: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n4 = n / 4;
  int threads = 256;
  int blocks = (int)((n4 + threads - 1) / threads);
  if (blocks > num_sms() * 16) blocks = num_sms() * 16;
  if (blocks < 1) blocks = 1;
#define PSW_CAST(S, D)                                                                          \
  do {                                                                                          \
    if (n4 > 0) cast_kernel<S, D><<<blocks, threads, 0, st>>>((const S*)src, (D*)dst, n4);      \
    if (n4 * 4 < n) cast_tail_kernel<S, D><<<1, 4, 0, st>>>((const S*)src, (D*)dst, n4 * 4, n); \
  } while (0)
  if (src_dtype == PSW_F32 && dst_dtype == PSW_BF16) PSW_CAST(float, bf16);
  else if (src_dtype == PSW_BF16 && dst_dtype == PSW_F32) PSW_CAST(bf16, float);
  else if (src_dtype == PSW_F32 && dst_dtype == PSW_F32) PSW_CAST(float, float);
  else if (src_dtype == PSW_BF16 && dst_dtype == PSW_BF16) PSW_CAST(bf16, bf16);
  else PSW_REQUIRE(false, PSW_ERR_BAD_ARG, "psw_cast: unknown dtype %d -> %d", src_dtype, dst_dtype);
#undef PSW_CAST
  return launch_status("cast_kernel");
}
