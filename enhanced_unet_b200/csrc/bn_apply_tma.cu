// BatchNorm apply + ReLU (reference models.py:220-224 in training: y -> relu(y * scale + shift)) for bf16 mode,
// C % 64 == 0, as a TMA load -> transform -> TMA store pipeline (the layout of tail_dec1_bwd_tma_kernel, which reaches
// 0.94 of the HBM roofline): a CTA owns one 64-channel block (its 16 constants per thread in registers), warp 0 streams
// 128-pixel tiles of the raw fp16 conv output through a 3-stage TMA ring, compute warp w transforms channel chunk w into a
// double-buffered, 128B-swizzled staging tile that ONE thread writes with a tensor store (row stride of the
// destination arbitrary: the skip slices of the concat buffers).  Two CTAs per SM.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct BnApplyTmaParams {
  const float* scale;
  const float* shift;
  int tiles, splits;
};

constexpr int kBaStages = 3;
constexpr int kBaTile = 128;
constexpr int kBaBytes = kBaTile * 128;       // 16 KB

template <bool F16>
__global__ void __launch_bounds__(288, 2)
bn_apply_relu_tma_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmOut, const BnApplyTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kBaStages], empty[kBaStages];
  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = sbase + kBaStages * kBaBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb = blockIdx.x;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kBaStages; ++s) { tc::mbar_init(tc::smem_u32(&full[s]), 1); tc::mbar_init(tc::smem_u32(&empty[s]), 8); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmY);
    tc::tma_prefetch_desc(&tmOut);
  }
  __syncthreads();

  if (warp == 0) {
    if (tc::elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.y; t < p.tiles; t += p.splits, ++it) {
        const uint32_t s = it % kBaStages;
        tc::mbar_wait(tc::smem_u32(&empty[s]), ((it / kBaStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full[s]);
        tc::mbar_expect_tx(fb, kBaBytes);
        tc::tma_load_2d(sbase + s * kBaBytes, &tmY, fb, cb * 64, t * kBaTile);
      }
    }
  } else {
    const int w = warp - 1, ctid = threadIdx.x - 32;
    float sc[8], sh[8];
    {
      const F8 a = load8(p.scale + cb * 64 + w * 8), b = load8(p.shift + cb * 64 + w * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] = a.v[e]; sh[e] = b.v[e]; }
    }
    uint32_t it = 0;
    for (int t = blockIdx.y; t < p.tiles; t += p.splits, ++it) {
      const uint32_t s = it % kBaStages;
      const uint32_t tile = sbase + s * kBaBytes, stg = out_base + (it & 1u) * kBaBytes;
      tc::mbar_wait(tc::smem_u32(&full[s]), (it / kBaStages) & 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = lane + 32 * k;
        const uint32_t off = (uint32_t)(px * 128) + ((uint32_t)(w ^ (px & 7)) << 4);
        uint32_t h0, h1, h2, h3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(tile + off));
        const uint32_t hw[4] = {h0, h1, h2, h3};
        uint32_t o[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e2]));
          o[e2] = pack16x2<F16>(fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.f), fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.f));
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + off), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
      }
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[s]));
      if (ctid == 0) tc::tma_store_wait_read<0>();      // the store of tile it-1 has left the buffer tile it+1 will use
      tc::named_bar_sync(1, 256);
      if (ctid == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)),
                     "r"(stg), "r"(cb * 64), "r"(t * kBaTile)
                     : "memory");
        tc::tma_store_commit();
      }
    }
    if (ctid == 0) tc::tma_store_wait<0>();
  }
}

// returns 0 = launched, 1 = not applicable (caller uses the ring kernel), < 0 = error
int bn_apply_relu_tma(const void* y, int ldy, void* out, int ldo, long long M, int C, const float* scale, const float* shift,
                      bool out_f16, cudaStream_t st) {
  if (C % 64 != 0 || M < 8 * kBaTile || M > 0x7fffffffLL) return 1;
  BnApplyTmaParams p;
  p.scale = scale; p.shift = shift;
  p.tiles = (int)((M + kBaTile - 1) / kBaTile);
  const int blocks = C / 64;
  int splits = (2 * kNumSMs) / blocks;
  if (splits < 1) splits = 1;
  if (splits > p.tiles) splits = p.tiles;
  p.splits = splits;
  CUtensorMap tmY, tmOut;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[1] = {(uint64_t)ldy * 2};
    uint32_t box[2] = {64u, (uint32_t)kBaTile};
    if (tc::encode_tensor_map_bf16(&tmY, y, 2, dims, str, box, 128)) return -1;          // fp16 bits, 2-byte elements
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M}, str[1] = {(uint64_t)ldo * 2};
    uint32_t box[2] = {64u, (uint32_t)kBaTile};
    if (tc::encode_tensor_map_bf16(&tmOut, out, 2, dims, str, box, 128)) return -1;
  }
  constexpr int SMEM = 1024 + kBaStages * kBaBytes + 2 * kBaBytes;
  auto kern = out_f16 ? bn_apply_relu_tma_kernel<true> : bn_apply_relu_tma_kernel<false>;
  // set on every launch: the attribute is per device, and a process may drive more than one GPU
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  EUNET_REQUIRE(e == cudaSuccess, "bn_apply_relu(tma): cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  dim3 grid((unsigned)blocks, (unsigned)splits);
  kern<<<grid, 288, SMEM, st>>>(tmY, tmOut, p);
  return check_launch("bn_apply_relu(tma)");
}

}  // namespace eunet
