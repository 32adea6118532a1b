// 3x3 convolution (stride 1, zero pad 1) as NHWC bf16 implicit GEMM on the Blackwell tensor cores.
//
//   forward / dgrad:  D[128 pixels, BN channels] (fp32, TMEM) += A[128 px, KC ch] . B[BN, KC]^T  per (tap, channel chunk)
//     A = activation tile fetched by ONE 4-D TMA box {KC, bw, bh, bb} at spatial offset (dx-1, dy-1): the
//         box rows are the 128 GEMM rows, and TMA's out-of-bounds zero fill IS the conv zero padding.
//     B = packed filter slice [Cout][9*Cin] (K-major), 2-D TMA box {KC, BN}.
//     dgrad is the same kernel run over dY with the flipped/transposed filter pack.
//   wgrad:  D[128 co, TAPS*KC] += dY^T[128 co, 64 px] . X_tap[64 px, KC ci]  (both operands MN-major,
//     i.e. the NHWC tiles are consumed as-is), split over pixel ranges with fp32 atomics at the end.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread tcgen05.mma
// issuer, warps 2..5 = epilogue (TMEM -> registers -> BN statistics / affine / ReLU -> global).
// smem ring of STAGES {A,B} tiles with full/empty mbarriers; tcgen05.commit releases ring slots.
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

int g_opt_conv_halo = 1;
int g_opt_cta_pair = 0;   // measured slower than the single-CTA kernels (see conv_halo2.cu): kept as an option
int g_opt_tma_store = 1;
int g_opt_bn192 = 0;      // measured slower than three N = 64 tiles with the TMA-store epilogue (1.37 vs 1.16 ms, dgrad of dec2.0)
int g_opt_a_ahead = 1;
extern int g_opt_tail_dbg;       // conv_tail_bwd.cu: timing experiments only
extern int g_opt_tail_out_tma;   // tail.cu: 1 (default) = TMA-pipelined tail_out_fwd
extern int g_opt_bn_tma;         // elementwise.cu: 1 (default) = TMA-pipelined bn_apply_relu

PixelTile choose_pixel_tile(int B, int H, int W, int pixels) {
  PixelTile best{};
  long long best_tiles = -1;
  for (int bw = 1; bw <= pixels && bw <= 256; bw <<= 1)
    for (int bh = 1; bw * bh <= pixels && bh <= 256; bh <<= 1) {
      const int bb = pixels / (bw * bh);
      if (bb > 256) continue;
      PixelTile t{bw, bh, bb, (W + bw - 1) / bw, (H + bh - 1) / bh, (B + bb - 1) / bb};
      const long long n = t.tiles();
      // fewer tiles first; then wider rows (longer contiguous TMA runs), then taller
      if (best_tiles < 0 || n < best_tiles || (n == best_tiles && (bw > best.bw || (bw == best.bw && bh > best.bh)))) {
        best = t;
        best_tiles = n;
      }
    }
  return best;
}

struct ConvFwdParams {
  __nv_bfloat16* y;
  int ldy;
  int B, H, W, Cin, Cout;
  int bw, bh, bb, tiles_x, tiles_y;
  double* stats;
  const float* scale;
  const float* shift;
  int relu;
  int out_f16;   // 1: store fp16 (raw pre-BN tensor, or any tensor in fp16 mode), 0: store bf16 (activation / gradient)
  uint32_t fmt16;   // tensor-core operand format of x and w: 1 = bf16, 0 = fp16
  float* amax;   // fp16 outputs: atomicMax of |y| when it exceeds the fp16 range (may be NULL)
};

struct ConvWgradParams {
  float* dw;
  int B, H, W, Cin, Cout;
  int bw, bh, bb, tiles_x, tiles_y;
  int tiles, tiles_per_split;
  uint32_t fmt16;
};

// sum over the 32 lanes of a warp of 32 per-lane values; lane l ends with the total of v[l]
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? v[i + s] : v[i];
      const float send = upper ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int KC, int BN, int STAGES>
__global__ void __launch_bounds__(192)
conv3x3_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvFwdParams p) {
  constexpr int A_BYTES = 128 * KC * 2, B_BYTES = BN * KC * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t LAYOUT = KC == 64 ? tc::kSwizzle128 : tc::kSwizzle32;
  constexpr uint32_t SBO = 8 * KC * 2;   // 8 rows of the K-major tile
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int CH = BN >= 32 ? 32 : 16;  // accumulator columns per tcgen05.ld
  static_assert(KC == 64 || KC == 16, "channel chunk");
  static_assert(BN == 16 || BN == 64 || BN == 128 || BN == 256, "N tile");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[2][4][BN];

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int m = blockIdx.x;
  const int tix = m % p.tiles_x, tiy = (m / p.tiles_x) % p.tiles_y, tib = m / (p.tiles_x * p.tiles_y);
  const int x0 = tix * p.bw, y0 = tiy * p.bh, b0 = tib * p.bb;
  const int n0 = blockIdx.y * BN;
  const int cchunks = p.Cin / KC;
  const int kiters = 9 * cchunks;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
    tc::mbar_init(tc::smem_u32(&accum_bar), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (tc::elect_one()) {
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = tc::smem_u32(&full_bar[s]);
        tc::mbar_expect_tx(fb, STAGE_BYTES);
        const int tap = it / cchunks, cc = it - tap * cchunks;
        const uint32_t a_dst = sbase + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
        tc::tma_load_4d(a_dst, &tmX, fb, cc * KC, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0);
        tc::tma_load_2d(b_dst, &tmW, fb, tap * p.Cin + cc * KC, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(128, BN, 0, 0, p.fmt16);
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
        tc::tc_fence_after();
        const uint32_t a_addr = sbase + s * STAGE_BYTES, b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int j = 0; j < KC / 16; ++j) {
          const uint64_t adesc = tc::make_smem_desc(a_addr + j * 32, 16, SBO, LAYOUT);
          const uint64_t bdesc = tc::make_smem_desc(b_addr + j * 32, 16, SBO, LAYOUT);
          tc::umma_bf16(tmem_base, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));   // frees the ring slot when these MMAs retire
      }
      tc::umma_commit(tc::smem_u32(&accum_bar));
    }
  } else {
    // ===== epilogue: 4 warps, warp q = warp % 4 owns TMEM lanes [32q, 32q+32) =====
    const int q = warp & 3;
    const int r = q * 32 + lane;                      // GEMM row == pixel of the tile
    const int xx = r % p.bw, yy = (r / p.bw) % p.bh, bb = r / (p.bw * p.bh);
    const int gx = x0 + xx, gy = y0 + yy, gb = b0 + bb;
    const bool valid = gx < p.W && gy < p.H && gb < p.B;
    __nv_bfloat16* yrow = p.y + (((long long)gb * p.H + gy) * p.W + gx) * p.ldy + n0;
    tc::mbar_wait(tc::smem_u32(&accum_bar), 0);
    tc::tc_fence_after();
    float amax = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (CH == 32) tc::tmem_ld32(taddr, raw);
      else {
        tc::tmem_ld16(taddr, raw);
#pragma unroll
        for (int i = 16; i < 32; ++i) raw[i] = 0u;
      }
      tc::tmem_ld_wait();
      if (p.stats != nullptr) {
        float s1[32], s2[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float f = valid ? __uint_as_float(raw[i]) : 0.f;
          s1[i] = f;
          s2[i] = f * f;
        }
        const float t1 = warp_transpose_sum32(s1, lane);
        const float t2 = warp_transpose_sum32(s2, lane);
        if (lane < CH) {
          red[0][q][c0 + lane] = t1;
          red[1][q][c0 + lane] = t2;
        }
      }
      if (valid) {
#pragma unroll
        for (int g8 = 0; g8 < CH / 8; ++g8) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float f = __uint_as_float(raw[g8 * 8 + e]);
            if (p.scale != nullptr) f = fmaf(f, __ldg(p.scale + n0 + c0 + g8 * 8 + e), __ldg(p.shift + n0 + c0 + g8 * 8 + e));
            if (p.relu) f = fmaxf(f, 0.f);
            o[e] = f;
            amax = fmaxf(amax, fabsf(f));
          }
          uint4 u;
          if (p.out_f16) {
            u.x = pack_f16x2(o[0], o[1]);
            u.y = pack_f16x2(o[2], o[3]);
            u.z = pack_f16x2(o[4], o[5]);
            u.w = pack_f16x2(o[6], o[7]);
          } else {
            u.x = pack_bf16x2(o[0], o[1]);
            u.y = pack_bf16x2(o[2], o[3]);
            u.z = pack_bf16x2(o[4], o[5]);
            u.w = pack_bf16x2(o[6], o[7]);
          }
          *reinterpret_cast<uint4*>(yrow + c0 + g8 * 8) = u;
        }
      }
    }
    if (p.amax != nullptr && p.out_f16 && amax > 65504.f) atomicMax(reinterpret_cast<unsigned int*>(p.amax), __float_as_uint(amax));
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps only
      for (int c = threadIdx.x - 64; c < BN; c += 128) {
        const float t1 = red[0][0][c] + red[0][1][c] + red[0][2][c] + red[0][3][c];
        const float t2 = red[1][0][c] + red[1][1][c] + red[1][2][c] + red[1][3][c];
        atomicAdd(p.stats + n0 + c, (double)t1);
        atomicAdd(p.stats + p.Cout + n0 + c, (double)t2);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------------------
// RH ("row halo", 8x8-pixel tiles, TAPS = 3): the three taps of a filter row come from ONE X box of 8 rows x 10 pixels
// (tap dx = +dx rows inside the box: LBO = one row, 8-pixel K groups 10 rows apart) instead of three shifted 8x8 boxes:
// 26 KB instead of 40 KB through L2 per stage (the wide layers pull ~28 TB/s L2->SM with three boxes).
template <int KC, int TAPS, int STAGES, bool RH>
__global__ void __launch_bounds__(192, 2)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                        const ConvWgradParams p) {
  constexpr int KP = 64;                                    // pixels (GEMM K) per stage
  constexpr int DYB = KP * 128;                             // one 64-channel dY block
  constexpr int A_BYTES = 2 * DYB;                          // 128 output channels
  constexpr int XB = KP * KC * 2;                           // one tap tile of X
  constexpr int B_BYTES = RH ? 80 * KC * 2 : TAPS * XB, STAGE_BYTES = A_BYTES + B_BYTES;
  static_assert(!RH || TAPS == 3, "row-halo X boxes serve the three taps of one filter row");
  constexpr int N = TAPS * KC;
  constexpr int TAPGROUPS = 9 / TAPS;
  constexpr uint32_t TMEM_COLS = 256;
  constexpr int CH = KC == 64 ? 32 : 16;
  static_assert(N <= 256 && N % 16 == 0 && 9 % TAPS == 0, "tap grouping");

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_s;

  const int t_begin = blockIdx.x * p.tiles_per_split;
  const int t_end = min(p.tiles, t_begin + p.tiles_per_split);
  const int kiters = t_end - t_begin;
  if (kiters <= 0) return;   // uniform for the whole CTA

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * 128;
  const int ci0 = (blockIdx.z / TAPGROUPS) * KC, tap0 = (blockIdx.z % TAPGROUPS) * TAPS;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&full_bar[s]), 1);
      tc::mbar_init(tc::smem_u32(&empty_bar[s]), 1);
    }
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
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        tc::mbar_wait(tc::smem_u32(&empty_bar[s]), ph ^ 1u);
        const uint32_t fb = tc::smem_u32(&full_bar[s]);
        tc::mbar_expect_tx(fb, STAGE_BYTES);
        const int t = t_begin + it;
        const int tix = t % p.tiles_x, tiy = (t / p.tiles_x) % p.tiles_y, tib = t / (p.tiles_x * p.tiles_y);
        const int x0 = tix * p.bw, y0 = tiy * p.bh, b0 = tib * p.bb;
        const uint32_t a_dst = sbase + s * STAGE_BYTES, b_dst = a_dst + A_BYTES;
        tc::tma_load_4d(a_dst, &tmDY, fb, co0, x0, y0, b0);
        tc::tma_load_4d(a_dst + DYB, &tmDY, fb, co0 + 64, x0, y0, b0);   // beyond Cout -> zero filled
        if (RH) {
          tc::tma_load_4d(b_dst, &tmX, fb, ci0, x0 - 1, y0 + tap0 / 3 - 1, b0);
        } else {
#pragma unroll
          for (int tt = 0; tt < TAPS; ++tt) {
            const int tap = tap0 + tt;
            tc::tma_load_4d(b_dst + tt * XB, &tmX, fb, ci0, x0 + tap % 3 - 1, y0 + tap / 3 - 1, b0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(128, N, 1, 1, p.fmt16);   // both operands MN-major
      constexpr uint32_t B_LAYOUT = KC == 64 ? tc::kSwizzle128 : tc::kSwizzle32;
      constexpr uint32_t B_ROW = KC * 2;            // bytes per pixel row of an X tile
      for (int it = 0; it < kiters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        tc::mbar_wait(tc::smem_u32(&full_bar[s]), ph);
        tc::tc_fence_after();
        const uint32_t a_addr = sbase + s * STAGE_BYTES, b_addr = a_addr + A_BYTES;
#pragma unroll
        for (int j = 0; j < KP / 16; ++j) {
          // MN-major: LBO = stride between 64-/16-channel blocks, SBO = stride between 8-pixel groups
          const uint64_t adesc = tc::make_smem_desc(a_addr + j * 16 * 128, DYB, 8 * 128, tc::kSwizzle128);
          const uint64_t bdesc = RH ? tc::make_smem_desc(b_addr + j * 20 * B_ROW, B_ROW, 10 * B_ROW, B_LAYOUT)
                                    : tc::make_smem_desc(b_addr + j * 16 * B_ROW, XB, 8 * B_ROW, B_LAYOUT);
          tc::umma_bf16(tmem_base, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
        }
        tc::umma_commit(tc::smem_u32(&empty_bar[s]));
      }
      tc::umma_commit(tc::smem_u32(&accum_bar));
    }
  } else {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    tc::mbar_wait(tc::smem_u32(&accum_bar), 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += CH) {
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      if (CH == 32) tc::tmem_ld32(taddr, raw);
      else tc::tmem_ld16(taddr, raw);
      tc::tmem_ld_wait();
      if (co < p.Cout) {
        const int tt = c0 / KC, ci = c0 % KC;   // a CH-column chunk never straddles a tap (KC % CH == 0)
        // 16-byte vector reductions: this thread's CH values are contiguous (one (co, tap) row of the packed gradient),
        // but the rows of neighbouring lanes are 9*Cin floats apart, so every warp instruction touches 32 lines
        float* dst = p.dw + ((long long)co * 9 + tap0 + tt) * p.Cin + ci0 + ci;
#pragma unroll
        for (int i = 0; i < CH; i += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "r"(raw[i]), "r"(raw[i + 1]), "r"(raw[i + 2]),
                       "r"(raw[i + 3])
                       : "memory");
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int make_act_map(CUtensorMap* m, const void* base, int ld, int C, int B, int H, int W, int box_c, const PixelTile& t,
                        int swizzle) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * W, (uint64_t)ld * 2 * W * H};
  uint32_t box[4] = {(uint32_t)box_c, (uint32_t)t.bw, (uint32_t)t.bh, (uint32_t)t.bb};
  return tc::encode_tensor_map_bf16(m, base, 4, dims, str, box, swizzle);
}

template <int KC, int BN, int STAGES>
static int launch_fwd(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvFwdParams& p, long long mtiles, int ntiles,
                      cudaStream_t st) {
  constexpr int SMEM = STAGES * (128 * KC * 2 + BN * KC * 2) + 1024;
  auto kern = conv3x3_fwd_tc_kernel<KC, BN, STAGES>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_fwd: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  dim3 grid((unsigned)mtiles, (unsigned)ntiles);
  kern<<<grid, 192, SMEM, st>>>(tmX, tmW, p);
  return check_launch("conv3x3_fwd_tc");
}

static int conv3x3_fwd_bf16(const void* x, int ldx, const void* w, void* y, int ldy, int B, int H, int W, int Cin, int Cout,
                            double* stats, const float* scale, const float* shift, int relu, int out_raw, int f16, float* amax,
                            cudaStream_t st) {
  EUNET_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "conv3x3_fwd(bf16): Cin=%d and Cout=%d must be multiples of 16", Cin, Cout);
  EUNET_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= Cin && ldy >= Cout, "conv3x3_fwd(bf16): bad ld (%d, %d)", ldx, ldy);
  EUNET_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0, "conv3x3_fwd(bf16): y not 16-byte aligned");
  if (g_opt_conv_halo && g_opt_cta_pair && Cout % 256 != 0 && (Cout % 128 == 0 || (g_opt_cta_pair >= 2 && Cout % 64 == 0))) {
    const int rc = conv3x3_fwd_halo2(x, ldx, w, y, ldy, B, H, W, Cin, Cout, stats, scale, shift, relu, out_raw, f16, amax, st);
    if (rc <= 0) return rc;
  }
  if (g_opt_conv_halo) {
    const int rc = conv3x3_fwd_halo_bf16(x, ldx, w, y, ldy, B, H, W, Cin, Cout, stats, scale, shift, relu, out_raw, f16, amax, st);
    if (rc <= 0) return rc;   // launched (0) or failed (< 0); 1 = not covered, fall through to the per-tap kernel
  }
  EUNET_REQUIRE(y != nullptr, "conv3x3_fwd(bf16): y == NULL (statistics only) is implemented for 16 -> 64 channels on >= 8x8 images");
  const int KC = (Cin % 64 == 0) ? 64 : 16;
  const int BN = (Cout % 256 == 0) ? 256 : (Cout % 128 == 0) ? 128 : (Cout % 64 == 0) ? 64 : 16;
  const PixelTile t = choose_pixel_tile(B, H, W, 128);
  CUtensorMap tmX, tmW;
  if (make_act_map(&tmX, x, ldx, Cin, B, H, W, KC, t, KC == 64 ? 128 : 32)) return -1;
  {
    uint64_t dims[2] = {(uint64_t)9 * Cin, (uint64_t)Cout}, str[1] = {(uint64_t)9 * Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (tc::encode_tensor_map_bf16(&tmW, w, 2, dims, str, box, KC == 64 ? 128 : 32)) return -1;
  }
  ConvFwdParams p;
  p.y = (__nv_bfloat16*)y; p.ldy = ldy;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bw = t.bw; p.bh = t.bh; p.bb = t.bb; p.tiles_x = t.tiles_x; p.tiles_y = t.tiles_y;
  p.stats = stats; p.scale = scale; p.shift = shift; p.relu = relu; p.out_f16 = (out_raw || f16) ? 1 : 0;
  p.fmt16 = f16 ? 0u : 1u; p.amax = amax;
  const long long mt = t.tiles();
  const int nt = Cout / BN;
  EUNET_REQUIRE(mt <= 0x7fffffffLL, "conv3x3_fwd(bf16): too many tiles");
#define EUNET_FWD_CASE(kc, bn, stages) \
  if (KC == kc && BN == bn) return launch_fwd<kc, bn, stages>(tmX, tmW, p, mt, nt, st)
  EUNET_FWD_CASE(64, 256, 4);
  EUNET_FWD_CASE(64, 128, 3);
  EUNET_FWD_CASE(64, 64, 4);
  EUNET_FWD_CASE(64, 16, 4);
  EUNET_FWD_CASE(16, 256, 8);
  EUNET_FWD_CASE(16, 128, 8);
  EUNET_FWD_CASE(16, 64, 8);
  EUNET_FWD_CASE(16, 16, 8);
#undef EUNET_FWD_CASE
  set_error("conv3x3_fwd(bf16): no kernel for KC=%d BN=%d", KC, BN);
  return -1;
}

template <int KC, int TAPS, int STAGES, bool RH = false>
static int launch_wgrad(const CUtensorMap& tmX, const CUtensorMap& tmDY, ConvWgradParams p, cudaStream_t st) {
  constexpr int SMEM = STAGES * (2 * 64 * 128 + (RH ? 80 * KC * 2 : TAPS * 64 * KC * 2)) + 1024;
  auto kern = conv3x3_wgrad_tc_kernel<KC, TAPS, STAGES, RH>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_wgrad: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int co_tiles = (p.Cout + 127) / 128;
  const int zdim = (p.Cin / KC) * (9 / TAPS);
  const int cols = co_tiles * zdim;
  int splits = (2 * kNumSMs) / cols;                     // <= 2 CTAs of work per SM (no third, nearly empty wave)
  if (splits > p.tiles) splits = p.tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.tiles + splits - 1) / splits;
  splits = (p.tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  dim3 grid((unsigned)splits, (unsigned)co_tiles, (unsigned)zdim);
  kern<<<grid, 192, SMEM, st>>>(tmX, tmDY, p);
  return check_launch("conv3x3_wgrad_tc");
}

static int conv3x3_wgrad_bf16(const void* x, int ldx, const void* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                              int f16, cudaStream_t st) {
  EUNET_REQUIRE(Cin % 16 == 0 && Cout % 64 == 0, "conv3x3_wgrad(bf16): Cin=%d must be a multiple of 16 and Cout=%d of 64", Cin, Cout);
  EUNET_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && ldx >= Cin && lddy >= Cout, "conv3x3_wgrad(bf16): bad ld (%d, %d)", ldx, lddy);
  if (g_opt_conv_halo) {
    const int rc = conv3x3_wgrad_halo_bf16(x, ldx, dy, lddy, dw, B, H, W, Cin, Cout, f16, st);
    if (rc <= 0) return rc;
  }
  const int KC = (Cin % 64 == 0) ? 64 : 16;
  PixelTile t = choose_pixel_tile(B, H, W, 64);
  CUtensorMap tmX, tmDY;
  const bool row_halo = KC == 64 && H >= 8 && W >= 8 && g_opt_conv_halo != 5;
  if (row_halo) t = PixelTile{8, 8, 1, (W + 7) / 8, (H + 7) / 8, B};   // 8x8-pixel tiles: the row-halo X box is 8 rows x 10 pixels
  if (row_halo) {
    PixelTile tx = t;
    tx.bw = 10;                 // 8 rows x (8 + 2) pixels: the three dx taps are row shifts inside the box
    if (make_act_map(&tmX, x, ldx, Cin, B, H, W, KC, tx, 128)) return -1;
  } else if (make_act_map(&tmX, x, ldx, Cin, B, H, W, KC, t, KC == 64 ? 128 : 32)) {
    return -1;
  }
  if (make_act_map(&tmDY, dy, lddy, Cout, B, H, W, 64, t, 128)) return -1;
  ConvWgradParams p;
  p.dw = dw; p.fmt16 = f16 ? 0u : 1u;
  p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.bw = t.bw; p.bh = t.bh; p.bb = t.bb; p.tiles_x = t.tiles_x; p.tiles_y = t.tiles_y;
  EUNET_REQUIRE(t.tiles() <= 0x7fffffffLL, "conv3x3_wgrad(bf16): too many tiles");
  p.tiles = (int)t.tiles();
  p.tiles_per_split = 0;
  // two stages of 40 KB: TWO CTAs per SM (256 TMEM columns each), so one CTA's reduction epilogue runs under the other's MMAs
  if (KC == 64 && row_halo) return launch_wgrad<64, 3, 4, true>(tmX, tmDY, p, st);   // 4 x 26 KB: still two CTAs per SM
  if (KC == 64) return launch_wgrad<64, 3, 2>(tmX, tmDY, p, st);
  return launch_wgrad<16, 9, 4>(tmX, tmDY, p, st);
}

}  // namespace eunet

using namespace eunet;

extern "C" int eunet_set_option(const char* name, int value) {
  if (strcmp(name, "conv_halo") == 0) {
    g_opt_conv_halo = value;
    return 0;
  }
  if (strcmp(name, "tail_dbg") == 0) {
    g_opt_tail_dbg = value;
    return 0;
  }
  if (strcmp(name, "cta_pair") == 0) {
    g_opt_cta_pair = value;
    return 0;
  }
  if (strcmp(name, "bn_tma") == 0) {
    g_opt_bn_tma = value;
    return 0;
  }
  if (strcmp(name, "tail_out_tma") == 0) {
    g_opt_tail_out_tma = value;
    return 0;
  }
  if (strcmp(name, "a_ahead") == 0) {
    g_opt_a_ahead = value;
    return 0;
  }
  if (strcmp(name, "bn192") == 0) {
    g_opt_bn192 = value;
    return 0;
  }
  if (strcmp(name, "tma_store") == 0) {
    g_opt_tma_store = value;
    return 0;
  }
  set_error("unknown option '%s'", name);
  return -1;
}

extern "C" int eunet_conv3x3_fwd(const void* x, int ldx, const void* w_packed, void* y, int ldy, int dtype, int B, int H, int W,
                                 int Cin, int Cout, double* stats, const float* scale, const float* shift, int relu,
                                 int out_raw, float* amax, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_fwd: bad shape");
  EUNET_REQUIRE((scale == nullptr) == (shift == nullptr), "conv3x3_fwd: scale and shift must be given together");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EUNET_BF16 || dtype == EUNET_F16)
    return conv3x3_fwd_bf16(x, ldx, w_packed, y, ldy, B, H, W, Cin, Cout, stats, scale, shift, relu, out_raw, dtype == EUNET_F16, amax,
                            st);
  if (dtype == EUNET_F32) {
    EUNET_REQUIRE(Cin % 16 == 0 && Cout % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0, "conv3x3_fwd(f32): Cin%%16, Cout%%4, ld%%4");
    return conv3x3_fwd_f32((const float*)x, ldx, (const float*)w_packed, (float*)y, ldy, B, H, W, Cin, Cout, stats, scale, shift,
                           relu, st);
  }
  set_error("conv3x3_fwd: unknown dtype %d", dtype);
  return -1;
}

extern "C" int eunet_conv3x3_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_packed, int dtype, int B, int H,
                                   int W, int Cin, int Cout, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv3x3_wgrad: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == EUNET_BF16 || dtype == EUNET_F16)
    return conv3x3_wgrad_bf16(x, ldx, dy, lddy, dw_packed, B, H, W, Cin, Cout, dtype == EUNET_F16, st);
  if (dtype == EUNET_F32) {
    EUNET_REQUIRE(Cin % 4 == 0 && Cout % 4 == 0 && ldx % 4 == 0 && lddy % 4 == 0, "conv3x3_wgrad(f32): channels/ld %% 4");
    return conv3x3_wgrad_f32((const float*)x, ldx, (const float*)dy, lddy, dw_packed, B, H, W, Cin, Cout, st);
  }
  set_error("conv3x3_wgrad: unknown dtype %d", dtype);
  return -1;
}
