// v2 weight gradient of the 3x3 convolution: halo-tile implicit GEMM on tcgen05, all nine taps per CTA pass.
//
// Work item = (16x8 pixel tile of one image, 64-channel input chunk, 64-channel output tile).  Per tile ONE TMA box
// {64, 10, 18, 1} brings the X halo tile and ONE box {64, 8, 16, 1} the dense dY tile; both are consumed MN-major
// exactly as the NHWC boxes land in shared memory (K = pixels):
//     D_pair[(h, ci), co] += sum_p X[p + tap_h, ci] * dY[p, co]        h = 0, 1
// The A' operand stacks TWO taps in M = 128: its second 64-channel "block" is the same halo tile at the other tap's
// row shift (LBO = byte distance between the two tap origins; pinned by tests/test_gpu_probe.py).  Taps are paired
// (0,1) (2,3) (4,5) (6,7) (8,8) -> five 128x64 fp32 accumulators = 320 TMEM columns, so every X / dY byte is fetched
// once per item (30 B per MMA cycle instead of 104 B in the per-tap kernel).  Pixel tiles are split across CTAs and
// the 9x64x64 result is added to the packed fp32 gradient with atomics.
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"

namespace eunet {

struct WgradHaloParams {
  float* dw;
  int B, H, W, Cin, Cout;
  int blocks_x, blocks_y, tiles, tiles_per_split, ci_chunks;
  uint32_t fmt16;     // operand format of x and dy: 1 = bf16, 0 = fp16
};

template <int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                          const WgradHaloParams p) {
  constexpr int X_BYTES = 180 * 128, X_SLOT = 23 * 1024, D_BYTES = 128 * 128, STAGE_BYTES = X_SLOT + D_BYTES;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;

  // blockIdx.x = (ci chunk, co tile) column, blockIdx.y = pixel split: CTAs that share a pixel range are adjacent in
  // launch order, so the X / dY tiles they all read are fetched from HBM once and hit in L2 afterwards
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(p.tiles, t_begin + p.tiles_per_split);
  const int kiters = t_end - t_begin;
  if (kiters <= 0) return;

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci0 = (blockIdx.x % p.ci_chunks) * 64, co0 = (blockIdx.x / p.ci_chunks) * 64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(tc::smem_u32(&full_bar[s]), 1); tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1); }
    tc::mbar_init(tc::smem_u32(&accum_bar), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmDY);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), (((uint32_t)(it / STAGES)) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full_bar[s]);
        tc::mbar_expect_tx(fb, X_BYTES + D_BYTES);
        const int t = t_begin + it;
        const int bx = t % p.blocks_x, by = (t / p.blocks_x) % p.blocks_y, b = t / (p.blocks_x * p.blocks_y);
        const uint32_t xs = sbase + s * STAGE_BYTES;
        tc::tma_load_4d(xs, &tmX, fb, ci0, bx * 8 - 1, by * 16 - 1, b);
        tc::tma_load_4d(xs + X_SLOT, &tmDY, fb, co0, bx * 8, by * 16, b);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(128, 64, 1, 1, p.fmt16);   // both operands MN-major
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), ((uint32_t)(it / STAGES)) & 1u);
        tc::tc_fence_after();
        const uint32_t xs = sbase + s * STAGE_BYTES, ds = xs + X_SLOT;
#pragma unroll
        for (int pr = 0; pr < 5; ++pr) {
          const int t1 = 2 * pr, t2 = pr < 4 ? 2 * pr + 1 : 8;
          const uint32_t off1 = (uint32_t)(((t1 / 3) * 10 + (t1 % 3)) * 128), off2 = (uint32_t)(((t2 / 3) * 10 + (t2 % 3)) * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // A': K group (8 pixels of one output row) stride = 10 halo rows; K = 16 -> output rows 2j, 2j+1
            const uint64_t adesc = tc::make_smem_desc(xs + off1 + j * (2 * 10 * 128), off2 - off1, 10 * 128, tc::kSwizzle128);
            const uint64_t bdesc = tc::make_smem_desc(ds + j * 2048, 0, 1024, tc::kSwizzle128);
            tc::umma_bf16(tmem_base + pr * 64, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      tc::umma_commit(tc::smem_u32(&accum_bar));
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int h = r >> 6, ci = ci0 + (r & 63);
    tc::mbar_wait(tc::smem_u32(&accum_bar), 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int pr = 0; pr < 5; ++pr) {
      const int tap = pr < 4 ? 2 * pr + h : 8;
      const bool live = !(pr == 4 && h == 1);     // the duplicated half of the last pair
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t raw[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pr * 64 + c0), raw);
        tc::tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int co = co0 + c0 + i;
            if (co < p.Cout) atomicAdd(p.dw + ((long long)co * 9 + tap) * p.Cin + ci, __uint_as_float(raw[i]));
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// 16-channel inputs (the 3-channel layers enc1.0 / enhance.0, padded to 16): the layer is a pure stream over dY
// (128 B per pixel against 32 B of X), so the point is to touch every dY byte once and keep many tiles in flight.
// Roles are swapped against the kernel above: A = dY with M = 64 output channels (one 64-channel box; an M = 128
// instruction with a zero-filled second box spends its shared-memory read cycles on zeros: 57 instead of ~30 clocks
// per MMA, measured), B = the X halo tile, whose three taps of a filter row are three 16-channel
// MN-major blocks one halo row (32 B) apart (LBO = 32), 8-pixel K groups 10 halo rows apart:
//     D_dy[co, (dx, ci)] += sum_p dY[p, co] * X[p + (dy, dx), ci]        three accumulators of N = 48
// One X box + two dY boxes per 16x8-pixel tile instead of nine tap boxes per 64 pixels (per-tap kernel).
// ------------------------------------------------------------------------------------------------------------
template <int STAGES>
__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad16_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                            const WgradHaloParams p) {
  constexpr int X_BYTES = 180 * 32, X_SLOT = 6 * 1024, D_BYTES = 128 * 128, STAGE_BYTES = X_SLOT + D_BYTES;
  constexpr uint32_t TMEM_COLS = 256;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(p.tiles, t_begin + p.tiles_per_split);
  const int kiters = t_end - t_begin;
  if (kiters <= 0) return;

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(tc::smem_u32(&full_bar[s]), 1); tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1); }
    tc::mbar_init(tc::smem_u32(&accum_bar), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmDY);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), (((uint32_t)(it / STAGES)) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full_bar[s]);
        tc::mbar_expect_tx(fb, X_BYTES + D_BYTES);
        const int t = t_begin + it;
        const int bx = t % p.blocks_x, by = (t / p.blocks_x) % p.blocks_y, b = t / (p.blocks_x * p.blocks_y);
        const uint32_t xs = sbase + s * STAGE_BYTES;
        tc::tma_load_4d(xs, &tmX, fb, 0, bx * 8 - 1, by * 16 - 1, b);
        tc::tma_load_4d(xs + X_SLOT, &tmDY, fb, co0, bx * 8, by * 16, b);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(64, 48, 1, 1, p.fmt16);   // both operands MN-major
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), ((uint32_t)(it / STAGES)) & 1u);
        tc::tc_fence_after();
        const uint32_t xs = sbase + s * STAGE_BYTES, ds = xs + X_SLOT;
#pragma unroll
        for (int j = 0; j < 8; ++j) {       // K = 16 pixels = output rows 2j, 2j+1 of the tile
          const uint64_t adesc = tc::make_smem_desc(ds + j * 2048, 0, 1024, tc::kSwizzle128);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t bdesc = tc::make_smem_desc(xs + (uint32_t)((dy * 10 + j * 20) * 32), 32, 10 * 32, tc::kSwizzle32);
            tc::umma_bf16(tmem_base + dy * 48, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
          }
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      tc::umma_commit(tc::smem_u32(&accum_bar));
    }
  } else {
    // M = 64 accumulator layout (cta_group::1): row 16 q + l lives in TMEM lane 32 q + l, l < 16
    const int q = warp & 3;
    const int co = co0 + q * 16 + lane;
    tc::mbar_wait(tc::smem_u32(&accum_bar), 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {      // TMEM column of (dy, dx, ci) = dy * 48 + dx * 16 + ci = tap * 16 + ci
      uint32_t raw[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tap * 16), raw);
      tc::tmem_ld_wait();
      if (lane < 16 && co < p.Cout) {
        float* dst = p.dw + ((long long)co * 9 + tap) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(dst + i, __uint_as_float(raw[i]));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

static int conv3x3_wgrad16_halo(const void* x, int ldx, const void* dy, int lddy, float* dw, int B, int H, int W, int Cout,
                                int f16, cudaStream_t st) {
  constexpr int STAGES = 8;
  constexpr int SMEM = 1024 + STAGES * (6 * 1024 + 128 * 128);
  WgradHaloParams p;
  p.dw = dw; p.B = B; p.H = H; p.W = W; p.Cin = 16; p.Cout = Cout; p.fmt16 = f16 ? 0u : 1u;
  p.blocks_x = (W + 7) / 8;
  p.blocks_y = (H + 15) / 16;
  const long long tiles = (long long)p.blocks_x * p.blocks_y * B;
  if (tiles > 0x7fffffffLL) return 1;
  p.tiles = (int)tiles;
  p.ci_chunks = 1;
  const int co_tiles = (Cout + 63) / 64;
  int splits = kNumSMs / co_tiles;          // one CTA per SM (177 KB of stages each)
  if (splits > p.tiles) splits = p.tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles + splits - 1) / splits;
  splits = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  CUtensorMap tmX, tmDY;
  {
    uint64_t dims[4] = {16ull, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    uint32_t box[4] = {16u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmX, x, 4, dims, str, box, 32)) return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)lddy * 2, (uint64_t)lddy * 2 * W, (uint64_t)lddy * 2 * W * H};
    uint32_t box[4] = {64u, 8u, 16u, 1u};
    if (tc::encode_tensor_map_bf16(&tmDY, dy, 4, dims, str, box, 128)) return -1;
  }
  auto kern = conv3x3_wgrad16_halo_kernel<STAGES>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_wgrad16_halo: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  dim3 grid((unsigned)co_tiles, (unsigned)splits);
  kern<<<grid, 192, SMEM, st>>>(tmX, tmDY, p);
  return check_launch("conv3x3_wgrad16_halo");
}

// returns 0 = launched, 1 = shape not covered, < 0 = error
int conv3x3_wgrad_halo_bf16(const void* x, int ldx, const void* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                            int f16, cudaStream_t st) {
  if (H < 8 || W < 8) return 1;
  if (Cin == 16) return conv3x3_wgrad16_halo(x, ldx, dy, lddy, dw, B, H, W, Cout, f16, st);
  if (Cin % 64 != 0) return 1;
  // Cout a multiple of 128: the per-tap kernel (M = 128 output channels, N = 192 per MMA, two CTAs per SM, row-halo X
  // boxes) is not shared-memory-port bound and wins (measured 1.0-1.25 PF against 0.85-1.02 PF here); this kernel keeps
  // the Cout = 64 layers, where the per-tap kernel would waste half of its M = 128 rows
  if (g_opt_conv_halo < 2 && Cout % 128 == 0) return 1;
  constexpr int STAGES = 4;
  constexpr int SMEM = 1024 + STAGES * (23 * 1024 + 128 * 128);
  WgradHaloParams p;
  p.dw = dw; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.fmt16 = f16 ? 0u : 1u;
  p.blocks_x = (W + 7) / 8;
  p.blocks_y = (H + 15) / 16;
  const long long tiles = (long long)p.blocks_x * p.blocks_y * B;
  if (tiles > 0x7fffffffLL) return 1;
  p.tiles = (int)tiles;
  p.ci_chunks = Cin / 64;
  const int co_tiles = (Cout + 63) / 64;
  const int cols = p.ci_chunks * co_tiles;
  int splits = (2 * kNumSMs) / cols;   // cols * splits <= 2 CTAs per SM: never spill into a third, nearly empty wave
  if (splits > p.tiles) splits = p.tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles + splits - 1) / splits;
  splits = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  CUtensorMap tmX, tmDY;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    uint32_t box[4] = {64u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmX, x, 4, dims, str, box, 128)) return -1;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)lddy * 2, (uint64_t)lddy * 2 * W, (uint64_t)lddy * 2 * W * H};
    uint32_t box[4] = {64u, 8u, 16u, 1u};
    if (tc::encode_tensor_map_bf16(&tmDY, dy, 4, dims, str, box, 128)) return -1;
  }
  auto kern = conv3x3_wgrad_halo_kernel<STAGES>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_wgrad_halo: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  dim3 grid((unsigned)cols, (unsigned)splits);
  kern<<<grid, 192, SMEM, st>>>(tmX, tmDY, p);
  return check_launch("conv3x3_wgrad_halo");
}

}  // namespace eunet
