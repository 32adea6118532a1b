// Data gradient of a 3x3 convolution whose INPUT has only a few (<= 3) real channels: enhance.0 (3 -> 64 at 2Hx2W,
// reference models.py:309).  The generic dgrad (conv3x3_halo_kernel<64,16,4>) pads the 3 channels to N = 16 and
// re-reads the 64-channel dY tile from shared memory once per tap: 9 x 16 KB of operand reads per 128 pixels make it
// shared-memory-port bound at 2.4 TB/s of HBM traffic.  Here the roles are transposed so that every dY row is read
// ONCE:
//     U[(i, tap), p] = sum_c Wt[i][tap][c] * dY[p][c]          A = packed flipped filters (M = 64 rows, 27 real)
//                                                             B = the dY halo tile, all 180 rows (N = 184), K = 64
//     dX[q][i]       = sum_tap U[(i, tap), q + tap - (1,1)]    "col2im" gather of 9 values per output channel
// One tcgen05.mma set of four K = 16 instructions (M = 64, N = 184) per 16x8-pixel tile; U^T goes TMEM -> shared
// memory (27 rows x 180 columns fp32) and the 128 interior pixels gather their 27 values from there.  Out-of-image
// halo rows are zero-filled by TMA, which is exactly the transposed convolution's boundary condition.
// Output: fp32 [pixels][4] (3 gradients + 0), consumed by tail_up_bwd.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue (the four warps
// on TMEM lane quarters 0 / 1 drain the accumulator into shared memory, the other four gather from it).
#include <cuda_bf16.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct DgradFewParams {
  float* dx4;
  int B, H, W;
  int blocks_x, blocks_y, items;
  uint32_t fmt16;     // operand format of dy and the flipped filters: 1 = bf16, 0 = fp16
};

constexpr int kDfRows = 180;                 // halo tile rows: 18 x 10 pixels
constexpr int kDfN = 184;                    // MMA N (multiple of 8 >= 180); the 4 extra rows are slot padding
constexpr int kDfSlot = kDfN * 128;          // 23552 B = 23 KB: one stage of the dY ring (1024-byte aligned)
constexpr int kDfStages = 5;
constexpr int kDfSPitch = 188;               // floats per U^T row in shared memory
constexpr int kDfSBytes = 27 * kDfSPitch * 4;

__global__ void __launch_bounds__(320, 1)
conv3x3_dgrad_few_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmW,
                         const DgradFewParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kDfStages], a_empty[kDfStages], acc_full[2], acc_empty[2], s_full[2], s_empty[2], w_full;
  __shared__ uint32_t tmem_base_s;

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = sbase, a_base = sbase + 8192, s_base = a_base + kDfStages * kDfSlot;
  float* const s_gen = reinterpret_cast<float*>(smem_raw + (s_base - tc::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kDfStages; ++s) { tc::mbar_init(tc::smem_u32(&a_full[s]), 1); tc::mbar_init(tc::smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tc::smem_u32(&acc_full[s]), 1); tc::mbar_init(tc::smem_u32(&acc_empty[s]), 4);
      tc::mbar_init(tc::smem_u32(&s_full[s]), 4); tc::mbar_init(tc::smem_u32(&s_empty[s]), 4);
    }
    tc::mbar_init(tc::smem_u32(&w_full), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmDY);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t wb = tc::smem_u32(&w_full);
      tc::mbar_expect_tx(wb, 8192);
      tc::tma_load_2d(w_base, &tmW, wb, 0, 0);       // rows (i * 9 + tap) of the packed filters, 64 channels each
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
        const uint32_t s = it % kDfStages;
        tc::mbar_wait(tc::smem_u32(&a_empty[s]), ((it / kDfStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&a_full[s]);
        tc::mbar_expect_tx(fb, kDfRows * 128);
        tc::tma_load_4d(a_base + s * kDfSlot, &tmDY, fb, 0, bx * 8 - 1, by * 16 - 1, b);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(64, kDfN, 0, 0, p.fmt16);
      tc::mbar_wait(tc::smem_u32(&w_full), 0);
      tc::tc_fence_after();
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t s = it % kDfStages, ab = it & 1u;
        tc::mbar_wait(tc::smem_u32(&acc_empty[ab]), ((it >> 1) & 1u) ^ 1u);
        tc::mbar_wait(tc::smem_u32(&a_full[s]), (it / kDfStages) & 1u);
        tc::tc_fence_after();
        const uint32_t t_addr = a_base + s * kDfSlot;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t adesc = tc::make_smem_desc(w_base + j * 32, 16, 1024, tc::kSwizzle128);
          const uint64_t bdesc = tc::make_smem_desc(t_addr + j * 32, 16, 1024, tc::kSwizzle128);
          tc::umma_bf16(tmem_base + ab * 256, adesc, bdesc, idesc, j != 0 ? 1u : 0u);
        }
        tc::umma_commit(tc::smem_u32(&a_empty[s]));
        tc::umma_commit(tc::smem_u32(&acc_full[ab]));
      }
    }
  } else {
    // M = 64 accumulator layout (cta_group::1): row 16 q + l lives in TMEM lane 32 q + l, l < 16.  Rows 0..26 are real,
    // so only lane quarters 0 and 1 hold data: the four warps on those quarters (two per quarter, splitting the 192
    // columns) move U^T to shared memory; the four warps on quarters 2 / 3 gather (one interior pixel per thread).
    // The two groups hand the double-buffered U^T over through s_full / s_empty, so draining item i+1 overlaps
    // gathering item i.
    const int e = warp - 2, q = warp & 3;
    if (q < 2) {
      const int chalf = e >> 2;                            // 0: columns [0, 96), 1: [96, 192)
      const int row = q * 16 + lane;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t ab = it & 1u, ph = (it >> 1) & 1u;
        float* S = s_gen + ab * (kDfSBytes / 4);
        tc::mbar_wait(tc::smem_u32(&acc_full[ab]), ph);
        tc::tc_fence_after();
        uint32_t raw[3][32];
#pragma unroll
        for (int c = 0; c < 3; ++c)
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 256 + chalf * 96 + c * 32), raw[c]);
        tc::tmem_ld_wait();
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[ab]));      // TMEM drained: the next MMA set may start
        tc::mbar_wait(tc::smem_u32(&s_empty[ab]), ph ^ 1u);
        if (lane < 16 && row < 27) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int col0 = chalf * 96 + c * 32;
            float4* dst = reinterpret_cast<float4*>(S + row * kDfSPitch + col0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (col0 + 4 * i < kDfSPitch)
                dst[i] = make_float4(__uint_as_float(raw[c][4 * i]), __uint_as_float(raw[c][4 * i + 1]),
                                     __uint_as_float(raw[c][4 * i + 2]), __uint_as_float(raw[c][4 * i + 3]));
            }
          }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&s_full[ab]));
      }
    } else {
      const int tid = ((e >> 2) * 2 + (q - 2)) * 32 + lane;   // 0..127: pixel (ty, tx) = (tid / 8, tid % 8)
      const int ty = tid >> 3, tx = tid & 7;
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
        const uint32_t ab = it & 1u, ph = (it >> 1) & 1u;
        const float* S = s_gen + ab * (kDfSBytes / 4);
        const int gx = bx * 8 + tx, gy = by * 16 + ty;
        tc::mbar_wait(tc::smem_u32(&s_full[ab]), ph);
        float o[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          float acc = 0.f;
#pragma unroll
          for (int t = 0; t < 9; ++t) acc += S[(i * 9 + t) * kDfSPitch + (ty + t / 3) * 10 + tx + t % 3];
          o[i] = acc;
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&s_empty[ab]));
        if (gx < p.W && gy < p.H)
          reinterpret_cast<float4*>(p.dx4)[((long long)b * p.H + gy) * p.W + gx] = make_float4(o[0], o[1], o[2], 0.f);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace eunet

using namespace eunet;

extern "C" int eunet_conv3x3_dgrad_few(const void* dy, int lddy, const void* w_packed_flip, float* dx4, int dtype, int B, int H,
                                       int W, int Cout, int cin_pad, void* stream) {
  EUNET_REQUIRE(dtype == EUNET_BF16 || dtype == EUNET_F16, "conv3x3_dgrad_few: tensor-core path only (dtype %d)", dtype);
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "conv3x3_dgrad_few: bad shape");
  EUNET_REQUIRE(Cout == 64 && cin_pad >= 8 && cin_pad % 8 == 0, "conv3x3_dgrad_few: needs 64 gradient channels and a packed "
                "filter of >= 8 input-channel rows (got Cout=%d, cin_pad=%d)", Cout, cin_pad);
  EUNET_REQUIRE(lddy % 8 == 0 && lddy >= 64, "conv3x3_dgrad_few: bad lddy %d", lddy);
  DgradFewParams p;
  p.dx4 = dx4; p.B = B; p.H = H; p.W = W; p.fmt16 = dtype == EUNET_F16 ? 0u : 1u;
  p.blocks_x = (W + 7) / 8;
  p.blocks_y = (H + 15) / 16;
  const long long items = (long long)p.blocks_x * p.blocks_y * B;
  EUNET_REQUIRE(items <= 0x7fffffffLL, "conv3x3_dgrad_few: too many tiles");
  p.items = (int)items;
  CUtensorMap tmDY, tmW;
  {
    uint64_t dims[4] = {64ull, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)lddy * 2, (uint64_t)lddy * 2 * W, (uint64_t)lddy * 2 * W * H};
    uint32_t box[4] = {64u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmDY, dy, 4, dims, str, box, 128)) return -1;
  }
  {
    // packed flipped filters [cin_pad][9][64]: row i * 9 + tap; the first 64 rows cover input channels 0..6 (>= the 3 real)
    uint64_t dims[2] = {64ull, (uint64_t)cin_pad * 9}, str[1] = {128ull};
    uint32_t box[2] = {64u, 64u};
    if (tc::encode_tensor_map_bf16(&tmW, w_packed_flip, 2, dims, str, box, 128)) return -1;
  }
  constexpr int SMEM = 1024 + 8192 + kDfStages * kDfSlot + 2 * kDfSBytes;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(conv3x3_dgrad_few_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_dgrad_few: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  conv3x3_dgrad_few_kernel<<<grid, 320, SMEM, (cudaStream_t)stream>>>(tmDY, tmW, p);
  return check_launch("conv3x3_dgrad_few");
}
