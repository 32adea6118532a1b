// Error plumbing, version/query entry points and host-side TMA tensor-map encoding.
#include <stdarg.h>
#include <string.h>
#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

int encode_tensor_map_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  EUNET_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  EUNET_REQUIRE(rank >= 2 && rank <= 5, "tensor map rank %d unsupported", rank);
  EUNET_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16-byte aligned", base);
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    EUNET_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map box[%d]=%u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    EUNET_REQUIRE((strides_bytes[i] & 15) == 0, "tensor map stride[%d]=%llu not a multiple of 16 bytes", i,
                  (unsigned long long)strides_bytes[i]);
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else EUNET_REQUIRE(swizzle_bytes == 0, "bad swizzle %d", swizzle_bytes);
  if (swizzle_bytes)
    EUNET_REQUIRE(box[0] * 2 <= (uint32_t)swizzle_bytes, "inner box %u bf16 exceeds swizzle span %d B", box[0], swizzle_bytes);
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EUNET_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace tc
}  // namespace eunet

extern "C" {

const char* eunet_last_error(void) { return eunet::g_err; }

int eunet_abi_version(void) { return EUNET_ABI_VERSION; }

int eunet_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    (void)cudaGetLastError();
    eunet::set_error("no CUDA device");
    return -1;
  }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
  return 0;
}

}  // extern "C"
