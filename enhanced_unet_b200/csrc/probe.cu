// UMMA/TMA probe: a tiny "interpreter" that (1) runs a host-specified list of TMA tile loads into
// shared memory, (2) issues a host-specified list of tcgen05.mma instructions with raw 64-bit
// shared-memory descriptors / 32-bit instruction descriptors, and (3) dumps TMEM and the raw shared
// memory back to global memory.  tests/test_gpu_probe.py uses it to pin the descriptor encodings
// (K-major / MN-major, 128B / 32B swizzle, row-shifted starts) against numpy matmuls, so the conv
// kernels' layouts are verified facts, not guesses.
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct ProbeLoad {
  int32_t map;       // 0: 2D map A, 1: 2D map B, 2: 4D map X
  int32_t c[4];
  uint32_t smem_off;  // bytes from the 1024-aligned smem base
};
struct ProbeMma {
  uint64_t adesc, bdesc;   // start-address field is relative to the smem base
  uint32_t idesc, accumulate, tmem_col, pad;
};
struct ProbeParams {
  int32_t n_loads, n_mma;
  uint32_t tx_bytes;
  int32_t ncols;          // TMEM columns to allocate/dump (power of two >= 32)
  int32_t smem_dump_bytes;
  int32_t pad[3];
  ProbeLoad loads[24];
  ProbeMma mmas[32];
};
static_assert(sizeof(ProbeParams) == 32 + 24 * 24 + 32 * 32, "ProbeParams layout is mirrored in Python");

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
             const __grid_constant__ CUtensorMap mapX, const __grid_constant__ ProbeParams p, float* __restrict__ out_tmem,
             uint8_t* __restrict__ out_smem) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = tc::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // zero the dumped region so untouched bytes are recognisable
  for (int i = threadIdx.x; i < p.smem_dump_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xdeadbeefu;
  if (threadIdx.x == 0) {
    tc::mbar_init(tc::smem_u32(&bars[0]), 1);
    tc::mbar_init(tc::smem_u32(&bars[1]), 1);
    tc::mbar_fence_init();
  }
  if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), (uint32_t)p.ncols);
  // make generic-proxy smem writes visible to the async proxy (TMA) and publish barrier init / tmem base
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    const uint32_t bar_load = tc::smem_u32(&bars[0]);
    tc::mbar_expect_tx(bar_load, p.tx_bytes);
    for (int i = 0; i < p.n_loads; ++i) {
      const ProbeLoad& l = p.loads[i];
      if (l.map == 0) tc::tma_load_2d(sbase + l.smem_off, &mapA, bar_load, l.c[0], l.c[1]);
      else if (l.map == 1) tc::tma_load_2d(sbase + l.smem_off, &mapB, bar_load, l.c[0], l.c[1]);
      else tc::tma_load_4d(sbase + l.smem_off, &mapX, bar_load, l.c[0], l.c[1], l.c[2], l.c[3]);
    }
    tc::mbar_wait(bar_load, 0);
    tc::tc_fence_after();
    for (int i = 0; i < p.n_mma; ++i) {
      const ProbeMma& m = p.mmas[i];
      const uint64_t add = (uint64_t)((sbase >> 4) & 0x3fffu);
      tc::umma_bf16(tmem_base + m.tmem_col, m.adesc + add, m.bdesc + add, m.idesc, m.accumulate);
    }
    tc::umma_commit(tc::smem_u32(&bars[1]));
  }
  __syncwarp();
  tc::mbar_wait(tc::smem_u32(&bars[1]), 0);
  tc::tc_fence_after();

  for (int c0 = 0; c0 < p.ncols; c0 += 16) {
    uint32_t r[16];
    tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tc::tmem_ld_wait();
    float* o = out_tmem + (size_t)(warp * 32 + lane) * p.ncols + c0;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
  }
  for (int i = threadIdx.x; i < p.smem_dump_bytes; i += blockDim.x) out_smem[i] = smem[i];

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_base, (uint32_t)p.ncols);
}

}  // namespace eunet

extern "C" int eunet_probe_umma(const void* a, int a_rows, int a_cols, int a_box_rows, int a_box_cols, int a_swizzle,
                                const void* b, int b_rows, int b_cols, int b_box_rows, int b_box_cols, int b_swizzle,
                                const void* x, const int* x_dims, const int* x_box, int x_swizzle, const void* params_blob,
                                int params_bytes, float* out_tmem, void* out_smem, int smem_bytes, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(params_bytes == (int)sizeof(ProbeParams), "probe params blob is %d bytes, expected %zu", params_bytes,
                sizeof(ProbeParams));
  ProbeParams p;
  memcpy(&p, params_blob, sizeof(p));
  EUNET_REQUIRE(p.n_loads >= 0 && p.n_loads <= 24 && p.n_mma >= 0 && p.n_mma <= 32, "probe: bad counts");
  EUNET_REQUIRE(p.ncols >= 32 && p.ncols <= 512 && (p.ncols & (p.ncols - 1)) == 0, "probe: ncols must be a power of two");
  EUNET_REQUIRE(smem_bytes >= p.smem_dump_bytes && smem_bytes <= 200 * 1024, "probe: bad smem size");
  CUtensorMap ma, mb, mx;
  {
    uint64_t dims[2] = {(uint64_t)a_cols, (uint64_t)a_rows}, str[1] = {(uint64_t)a_cols * 2};
    uint32_t box[2] = {(uint32_t)a_box_cols, (uint32_t)a_box_rows};
    if (tc::encode_tensor_map_bf16(&ma, a, 2, dims, str, box, a_swizzle)) return -1;
  }
  {
    uint64_t dims[2] = {(uint64_t)b_cols, (uint64_t)b_rows}, str[1] = {(uint64_t)b_cols * 2};
    uint32_t box[2] = {(uint32_t)b_box_cols, (uint32_t)b_box_rows};
    if (tc::encode_tensor_map_bf16(&mb, b, 2, dims, str, box, b_swizzle)) return -1;
  }
  {
    // x_dims = {C, W, H, B} of a dense NHWC bf16 tensor
    uint64_t dims[4] = {(uint64_t)x_dims[0], (uint64_t)x_dims[1], (uint64_t)x_dims[2], (uint64_t)x_dims[3]};
    uint64_t str[3] = {dims[0] * 2, dims[0] * dims[1] * 2, dims[0] * dims[1] * dims[2] * 2};
    uint32_t box[4] = {(uint32_t)x_box[0], (uint32_t)x_box[1], (uint32_t)x_box[2], (uint32_t)x_box[3]};
    if (tc::encode_tensor_map_bf16(&mx, x, 4, dims, str, box, x_swizzle)) return -1;
  }
  cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes + 1024);
  EUNET_REQUIRE(e == cudaSuccess, "probe: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  probe_kernel<<<1, 128, smem_bytes + 1024, (cudaStream_t)stream>>>(ma, mb, mx, p, out_tmem, (uint8_t*)out_smem);
  return check_launch("probe_kernel");
}

// ---- cluster-launch probe: a CTA pair launched exactly as conv_halo2.cu launches its kernel ----
namespace eunet {
__global__ void __launch_bounds__(320, 1) probe_cluster_kernel(int* out) {
  extern __shared__ uint8_t sm[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) out[blockIdx.y] = (int)r + (sm[0] & 0);
}
}  // namespace eunet

extern "C" int eunet_probe_cluster(int* out, int ctas, int smem_bytes, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(ctas >= 2 && ctas % 2 == 0 && ctas <= 256, "probe_cluster: bad CTA count");
  cudaError_t e = cudaFuncSetAttribute(probe_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  EUNET_REQUIRE(e == cudaSuccess, "probe_cluster: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, (unsigned)ctas);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 2; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, probe_cluster_kernel, out);
  EUNET_REQUIRE(e == cudaSuccess, "probe_cluster: launch: %s", cudaGetErrorString(e));
  return check_launch("probe_cluster");
}
