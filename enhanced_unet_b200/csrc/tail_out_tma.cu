// Training-mode forward of the 2Hx2W tail after the enhance.0 convolution (reference models.py:310-313 + 337):
//     out[b,k,p] = d1[p][k] + b3[k] + sum_c w3[k][c] * relu(mid[p][c] * scale[c] + shift[c])
// over the RAW fp16 conv output `mid` (2.15 GB at batch 16 / 512^2): a pure stream, 128 B in and 12 B out per pixel.
// (In inference this pass does not exist: the conv epilogue produces `out` directly, conv_halo.cu EPI = 1.)
//
// The thread-per-pixel kernel in tail.cu spends 88 shared-memory loads per pixel on the per-channel constants at 16 %
// occupancy (ncu: short_scoreboard) and reaches 0.5 of the HBM roofline.  Here the roles are turned around:
//   * two persistent CTAs per SM; warp 0 streams 128-pixel tiles of `mid` (16 KB, 128B-swizzled rows) and of the residual
//     d14 (2 KB) through a 3-stage TMA ring;
//   * compute warp w owns CHANNEL CHUNK w (8 channels): its 40 constants live in registers for the whole kernel, lane l
//     walks pixels l, l+32, l+64, l+96 of the tile - one conflict-free 16-byte shared-memory load per (pixel, chunk);
//   * the eight partial class sums of a pixel meet in a double-buffered shared-memory array; after ONE 256-thread named
//     barrier per tile the first four warps add them, the residual and the bias and store the three logit planes with
//     fully coalesced 128-byte rows.
#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

constexpr int kToStages = 3;                       // 3 x 18 KB + 32 KB of partials = 87 KB: TWO CTAs per SM
constexpr int kToTile = 128;                       // pixels per tile
constexpr int kToMid = kToTile * 128;              // 16 KB
constexpr int kToRes = kToTile * 16;               // 2 KB
constexpr int kToStage = kToMid + kToRes;          // 18 KB (multiple of 1024)
constexpr int kToPart = kToTile * 8 * 16;          // partial sums [pixel][chunk] float4: 16 KB per buffer

struct TailOutParams {
  const float* scale;
  const float* shift;
  const float* w3;
  const float* b3;
  float* out;
  long long M, HW;      // pixels in total / per image (2H * 2W)
  int tiles;
};

// DEC1 = true: the same streaming layout for z = dec1(d2) (1x1, 64 -> 3, models.py:212, evaluated at HxW): bf16 input rows
// (any row stride), no affine / ReLU / residual, output fp32 [pixels][4] (z4) instead of NCHW planes.
template <bool DEC1, bool F16>
__global__ void __launch_bounds__(288, 2)
tail_out_tma_kernel(const __grid_constant__ CUtensorMap tmMid, const __grid_constant__ CUtensorMap tmRes, const TailOutParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kToStages], empty[kToStages];
  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen0 = smem_raw + (sbase - tc::smem_u32(smem_raw));
  const uint32_t part_base = sbase + kToStages * kToStage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kToStages; ++s) { tc::mbar_init(tc::smem_u32(&full[s]), 1); tc::mbar_init(tc::smem_u32(&empty[s]), 8); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmMid);
    tc::tma_prefetch_desc(&tmRes);
  }
  __syncthreads();

  if (warp == 0) {
    if (tc::elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % kToStages;
        tc::mbar_wait(tc::smem_u32(&empty[s]), ((it / kToStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full[s]);
        tc::mbar_expect_tx(fb, DEC1 ? kToMid : kToStage);   // out-of-range rows of the last tile are zero-filled, bytes still count
        tc::tma_load_2d(sbase + s * kToStage, &tmMid, fb, 0, t * kToTile);
        if (!DEC1) tc::tma_load_2d(sbase + s * kToStage + kToMid, &tmRes, fb, 0, t * kToTile);
      }
    }
  } else {
    const int w = warp - 1;                           // channel chunk 0..7
    const int ctid = threadIdx.x - 32;                // 0..255
    float sc[8], sh[8], w0[8], w1[8], w2[8];
    {
      const F8 x0 = load8(p.w3 + w * 8), x1 = load8(p.w3 + 64 + w * 8), x2 = load8(p.w3 + 128 + w * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) { sc[e] = 1.f; sh[e] = 0.f; w0[e] = x0.v[e]; w1[e] = x1.v[e]; w2[e] = x2.v[e]; }
      if (!DEC1) {
        const F8 a = load8(p.scale + w * 8), b = load8(p.shift + w * 8);
#pragma unroll
        for (int e = 0; e < 8; ++e) { sc[e] = a.v[e]; sh[e] = b.v[e]; }
      }
    }
    const float bb0 = p.b3[0], bb1 = p.b3[1], bb2 = p.b3[2];
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const uint32_t s = it % kToStages;
      const uint32_t tile = sbase + s * kToStage;
      const uint32_t part = part_base + (it & 1u) * kToPart;
      tc::mbar_wait(tc::smem_u32(&full[s]), (it / kToStages) & 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = lane + 32 * k;
        uint32_t h0, h1, h2, h3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3)
                     : "r"(tile + (uint32_t)(px * 128) + ((uint32_t)(w ^ (px & 7)) << 4)));
        const uint32_t hw[4] = {h0, h1, h2, h3};
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          float a0, a1;
          if (DEC1) {      // bf16 / fp16 activations, plain dot products
            unpack16x2<F16>(hw[e2], a0, a1);
          } else {         // raw fp16 conv outputs: BN affine + ReLU first
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e2]));
            a0 = fmaxf(fmaf(f.x, sc[2 * e2], sh[2 * e2]), 0.f);
            a1 = fmaxf(fmaf(f.y, sc[2 * e2 + 1], sh[2 * e2 + 1]), 0.f);
          }
          s0 = fmaf(a1, w0[2 * e2 + 1], fmaf(a0, w0[2 * e2], s0));
          s1 = fmaf(a1, w1[2 * e2 + 1], fmaf(a0, w1[2 * e2], s1));
          s2 = fmaf(a1, w2[2 * e2 + 1], fmaf(a0, w2[2 * e2], s2));
        }
        // partial[px][chunk]: 16-byte slot, chunk XOR-swizzled by the pixel so that the 8 lanes of a quarter-warp store
        // (and the reducer later loads) conflict-free
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(part + (uint32_t)(px * 128) + ((uint32_t)(w ^ (px & 7)) << 4)),
                     "f"(s0), "f"(s1), "f"(s2), "f"(0.f)
                     : "memory");
      }
      tc::named_bar_sync(1, 256);                     // all partials of this tile are written; the tile itself is consumed
      if (ctid < kToTile) {
        const int px = ctid;
        float s0 = bb0, s1 = bb1, s2 = bb2;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a, b, d, pad;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(a), "=f"(b), "=f"(d), "=f"(pad)
                       : "r"(part + (uint32_t)(px * 128) + ((uint32_t)(c ^ (px & 7)) << 4)));
          s0 += a; s1 += b; s2 += d;
        }
        const long long pix = (long long)t * kToTile + px;
        if (DEC1) {
          if (pix < p.M) reinterpret_cast<float4*>(p.out)[pix] = make_float4(s0, s1, s2, 0.f);
        } else {
          float r0, r1, r2, rpad;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r0), "=f"(r1), "=f"(r2), "=f"(rpad) : "r"(tile + kToMid + (uint32_t)(px * 16)));
          if (pix < p.M) {
            const long long b = pix / p.HW, hw = pix - b * p.HW;
            float* o = p.out + b * 3 * p.HW + hw;
            o[0] = s0 + r0;
            o[p.HW] = s1 + r1;
            o[2 * p.HW] = s2 + r2;
          }
        }
      }
      // the stage may be refilled once every compute warp is past its reads (the reducers read the residual last)
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[s]));
    }
  }
  (void)gen0;
}

// returns 0 = launched, 1 = not applicable, < 0 = error
int tail_out_fwd_tma(const float* d14, const void* mid, const float* scale, const float* shift, const float* w3, const float* b3,
                     float* out, int B, int H, int W, cudaStream_t st) {
  TailOutParams p;
  p.scale = scale; p.shift = shift; p.w3 = w3; p.b3 = b3; p.out = out;
  p.HW = 4LL * H * W;
  p.M = p.HW * B;
  if (p.M < 4 * kToTile) return 1;
  const long long tiles = (p.M + kToTile - 1) / kToTile;
  if (tiles > 0x7fffffffLL) return 1;
  p.tiles = (int)tiles;
  CUtensorMap tmMid, tmRes;
  {
    uint64_t dims[2] = {64ull, (uint64_t)p.M}, str[1] = {128ull};
    uint32_t box[2] = {64u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmMid, mid, 2, dims, str, box, 128)) return -1;        // fp16 bits, 2-byte elements
  }
  {
    uint64_t dims[2] = {8ull, (uint64_t)p.M}, str[1] = {16ull};                           // fp32 x 4 per pixel = 8 two-byte units
    uint32_t box[2] = {8u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmRes, d14, 2, dims, str, box, 0)) return -1;
  }
  constexpr int SMEM = 1024 + kToStages * kToStage + 2 * kToPart;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(tail_out_tma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "tail_out_fwd: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.tiles < 2 * kNumSMs ? p.tiles : 2 * kNumSMs;   // two co-resident CTAs: one reduces / stores while the other streams
  tail_out_tma_kernel<false, false><<<grid, 288, SMEM, st>>>(tmMid, tmRes, p);
  return check_launch("tail_out_fwd(tma)");
}

// z4 = dec1(d2): returns 0 = launched, 1 = not applicable, < 0 = error
int tail_dec1_fwd_tma(const void* d2, int ldd2, const float* w1, const float* b1, float* z4, long long M, bool f16,
                      cudaStream_t st) {
  TailOutParams p;
  p.scale = nullptr; p.shift = nullptr; p.w3 = w1; p.b3 = b1; p.out = z4;
  p.HW = 1;
  p.M = M;
  if (M < 4 * kToTile) return 1;
  const long long tiles = (M + kToTile - 1) / kToTile;
  if (tiles > 0x7fffffffLL) return 1;
  p.tiles = (int)tiles;
  CUtensorMap tmIn;
  {
    uint64_t dims[2] = {64ull, (uint64_t)M}, str[1] = {(uint64_t)ldd2 * 2};
    uint32_t box[2] = {64u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmIn, d2, 2, dims, str, box, 128)) return -1;
  }
  constexpr int SMEM = 1024 + kToStages * kToStage + 2 * kToPart;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(f16 ? tail_out_tma_kernel<true, true> : tail_out_tma_kernel<true, false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "tail_dec1_fwd: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.tiles < 2 * kNumSMs ? p.tiles : 2 * kNumSMs;
  if (f16) tail_out_tma_kernel<true, true><<<grid, 288, SMEM, st>>>(tmIn, tmIn, p);
  else tail_out_tma_kernel<true, false><<<grid, 288, SMEM, st>>>(tmIn, tmIn, p);
  return check_launch("tail_dec1_fwd(tma)");
}

// ------------------------------------------------------------------------------------------------------------
// Backward pass 1 over (dout4, mid) in the same style: BN sums + enhance.3 weight / bias gradients
// (acc layout as tail_bwd_reduce_kernel: [0,64) sum g', [64,128) sum g' xhat, [128,320) dW3[k][c], [320,323) db3[k]).
// Channel chunk per warp: the masked sums P_k = sum m g_k, Q_k = sum m g_k xc (see tail.cu) accumulate in registers over
// all pixels a lane sees - no per-tile synchronisation at all; one shuffle tree + fp64 atomics per warp at the end.
// ------------------------------------------------------------------------------------------------------------
struct TailReduceParams {
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* w3;
  double* acc;
  long long M;
  int tiles;
};

constexpr int kTrStages = 4;      // 4 x 18 KB = 72 KB: two CTAs per SM

__global__ void __launch_bounds__(288, 2)
tail_reduce_tma_kernel(const __grid_constant__ CUtensorMap tmMid, const __grid_constant__ CUtensorMap tmG, const TailReduceParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kTrStages], empty[kTrStages];
  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kTrStages; ++s) { tc::mbar_init(tc::smem_u32(&full[s]), 1); tc::mbar_init(tc::smem_u32(&empty[s]), 8); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmMid);
    tc::tma_prefetch_desc(&tmG);
  }
  __syncthreads();

  if (warp == 0) {
    if (tc::elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % kTrStages;
        tc::mbar_wait(tc::smem_u32(&empty[s]), ((it / kTrStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full[s]);
        tc::mbar_expect_tx(fb, kToStage);
        tc::tma_load_2d(sbase + s * kToStage, &tmMid, fb, 0, t * kToTile);
        tc::tma_load_2d(sbase + s * kToStage + kToMid, &tmG, fb, 0, t * kToTile);   // rows past M are zero-filled: g = 0
      }
    }
  } else {
    const int w = warp - 1;                           // channel chunk 0..7
    float sc[8], mu[8], beta[8], P[3][8], Q[3][8], db0 = 0.f, db1 = 0.f, db2 = 0.f;
    {
      const F8 a = load8(p.scale + w * 8), b = load8(p.shift + w * 8), m = load8(p.mean + w * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sc[e] = a.v[e]; mu[e] = m.v[e];
        beta[e] = fmaf(a.v[e], m.v[e], b.v[e]);
        P[0][e] = P[1][e] = P[2][e] = Q[0][e] = Q[1][e] = Q[2][e] = 0.f;
      }
    }
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const uint32_t s = it % kTrStages;
      const uint32_t tile = sbase + s * kToStage;
      tc::mbar_wait(tc::smem_u32(&full[s]), (it / kTrStages) & 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = lane + 32 * k;
        uint32_t h0, h1, h2, h3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3)
                     : "r"(tile + (uint32_t)(px * 128) + ((uint32_t)(w ^ (px & 7)) << 4)));
        float g0, g1, g2, gpad;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(gpad) : "r"(tile + kToMid + (uint32_t)(px * 16)));
        const uint32_t hw[4] = {h0, h1, h2, h3};
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e2]));
          const float v[2] = {f.x, f.y};
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int e = 2 * e2 + h;
            const float xc = v[h] - mu[e];
            if (fmaf(xc, sc[e], beta[e]) > 0.f) {
              P[0][e] += g0; P[1][e] += g1; P[2][e] += g2;
              Q[0][e] = fmaf(g0, xc, Q[0][e]); Q[1][e] = fmaf(g1, xc, Q[1][e]); Q[2][e] = fmaf(g2, xc, Q[2][e]);
            }
          }
        }
        if (w == 0) { db0 += g0; db1 += g1; db2 += g2; }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[s]));
    }
    // per-channel totals over the warp's 32 lanes, then the five reductions of this chunk
    const F8 is = load8(p.invstd + w * 8), x0 = load8(p.w3 + w * 8), x1 = load8(p.w3 + 64 + w * 8), x2 = load8(p.w3 + 128 + w * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
#pragma unroll
      for (int k = 0; k < 3; ++k) { P[k][e] = warp_sum(P[k][e]); Q[k][e] = warp_sum(Q[k][e]); }
    }
    if (w == 0) { db0 = warp_sum(db0); db1 = warp_sum(db1); db2 = warp_sum(db2); }
    if (lane < 8) {
      const int e = lane, c = w * 8 + e;
      // select this lane's channel without dynamic register indexing
      float p0 = 0.f, p1 = 0.f, p2 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, sce = 0.f, be = 0.f, ise = 0.f, wa = 0.f, wb = 0.f, wc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i == e) {
          p0 = P[0][i]; p1 = P[1][i]; p2 = P[2][i]; q0 = Q[0][i]; q1 = Q[1][i]; q2 = Q[2][i];
          sce = sc[i]; be = beta[i]; ise = is.v[i]; wa = x0.v[i]; wb = x1.v[i]; wc = x2.v[i];
        }
      atomicAdd(p.acc + c, (double)(wa * p0 + wb * p1 + wc * p2));
      atomicAdd(p.acc + 64 + c, (double)(ise * (wa * q0 + wb * q1 + wc * q2)));
      atomicAdd(p.acc + 128 + c, (double)fmaf(sce, q0, be * p0));
      atomicAdd(p.acc + 192 + c, (double)fmaf(sce, q1, be * p1));
      atomicAdd(p.acc + 256 + c, (double)fmaf(sce, q2, be * p2));
    }
    if (w == 0 && lane == 0) {
      atomicAdd(p.acc + 320, (double)db0);
      atomicAdd(p.acc + 321, (double)db1);
      atomicAdd(p.acc + 322, (double)db2);
    }
  }
}

// returns 0 = launched, 1 = not applicable, < 0 = error
int tail_bwd_reduce_tma(const float* dout4, const void* mid, const float* scale, const float* shift, const float* mean,
                        const float* invstd, const float* w3, double* acc, int B, int H, int W, cudaStream_t st) {
  TailReduceParams p;
  p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd; p.w3 = w3; p.acc = acc;
  p.M = 4LL * H * W * B;
  if (p.M < 4 * kToTile) return 1;
  const long long tiles = (p.M + kToTile - 1) / kToTile;
  if (tiles > 0x7fffffffLL) return 1;
  p.tiles = (int)tiles;
  CUtensorMap tmMid, tmG;
  {
    uint64_t dims[2] = {64ull, (uint64_t)p.M}, str[1] = {128ull};
    uint32_t box[2] = {64u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmMid, mid, 2, dims, str, box, 128)) return -1;
  }
  {
    uint64_t dims[2] = {8ull, (uint64_t)p.M}, str[1] = {16ull};
    uint32_t box[2] = {8u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmG, dout4, 2, dims, str, box, 0)) return -1;
  }
  constexpr int SMEM = 1024 + kTrStages * kToStage;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(tail_reduce_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "tail_bwd_reduce: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.tiles < 2 * kNumSMs ? p.tiles : 2 * kNumSMs;
  tail_reduce_tma_kernel<<<grid, 288, SMEM, st>>>(tmMid, tmG, p);
  return check_launch("tail_bwd_reduce(tma)");
}

// ------------------------------------------------------------------------------------------------------------
// dec1 backward in the same style: dd2[p][c] = sum_k dz[p][k] w1[k][c] (bf16, any row stride, written with a TMA tensor
// store from a swizzled staging tile), dW1[k][c] += dz[p][k] d2[p][c] and db1[k] += dz[p][k] in registers
// (acc layout as tail_dec1_bwd_kernel: [0,192) dW1[k][c], [192,195) db1).
// ------------------------------------------------------------------------------------------------------------
struct Dec1BwdParams {
  const float* w1;
  double* acc;
  int tiles;
};

template <bool F16>
__global__ void __launch_bounds__(288, 2)
tail_dec1_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmD2, const __grid_constant__ CUtensorMap tmDz,
                         const __grid_constant__ CUtensorMap tmOut, const Dec1BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[kToStages], empty[kToStages];
  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t out_base = sbase + kToStages * kToStage;           // 2 x 16 KB staging, 1024-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kToStages; ++s) { tc::mbar_init(tc::smem_u32(&full[s]), 1); tc::mbar_init(tc::smem_u32(&empty[s]), 8); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmD2);
    tc::tma_prefetch_desc(&tmDz);
    tc::tma_prefetch_desc(&tmOut);
  }
  __syncthreads();

  if (warp == 0) {
    if (tc::elect_one()) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
        const uint32_t s = it % kToStages;
        tc::mbar_wait(tc::smem_u32(&empty[s]), ((it / kToStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&full[s]);
        tc::mbar_expect_tx(fb, kToStage);
        tc::tma_load_2d(sbase + s * kToStage, &tmD2, fb, 0, t * kToTile);
        tc::tma_load_2d(sbase + s * kToStage + kToMid, &tmDz, fb, 0, t * kToTile);      // rows past M: zero fill (dz = 0)
      }
    }
  } else {
    const int w = warp - 1;
    const int ctid = threadIdx.x - 32;
    float w0[8], w1[8], w2[8], dw0[8], dw1[8], dw2[8], db0 = 0.f, db1 = 0.f, db2 = 0.f;
    {
      const F8 x0 = load8(p.w1 + w * 8), x1 = load8(p.w1 + 64 + w * 8), x2 = load8(p.w1 + 128 + w * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) { w0[e] = x0.v[e]; w1[e] = x1.v[e]; w2[e] = x2.v[e]; dw0[e] = dw1[e] = dw2[e] = 0.f; }
    }
    uint32_t it = 0;
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x, ++it) {
      const uint32_t s = it % kToStages;
      const uint32_t tile = sbase + s * kToStage;
      const uint32_t stg = out_base + (it & 1u) * kToMid;
      tc::mbar_wait(tc::smem_u32(&full[s]), (it / kToStages) & 1u);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = lane + 32 * k;
        const uint32_t off = (uint32_t)(px * 128) + ((uint32_t)(w ^ (px & 7)) << 4);
        uint32_t h0, h1, h2, h3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(tile + off));
        float z0, z1, z2, zpad;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(z0), "=f"(z1), "=f"(z2), "=f"(zpad) : "r"(tile + kToMid + (uint32_t)(px * 16)));
        const uint32_t hw[4] = {h0, h1, h2, h3};
        float o[8];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          float v[2];
          unpack16x2<F16>(hw[e2], v[0], v[1]);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int e = 2 * e2 + h;
            o[e] = z0 * w0[e] + z1 * w1[e] + z2 * w2[e];
            dw0[e] = fmaf(z0, v[h], dw0[e]);
            dw1[e] = fmaf(z1, v[h], dw1[e]);
            dw2[e] = fmaf(z2, v[h], dw2[e]);
          }
        }
        if (w == 0) { db0 += z0; db1 += z1; db2 += z2; }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + off), "r"(pack16x2<F16>(o[0], o[1])), "r"(pack16x2<F16>(o[2], o[3])),
                     "r"(pack16x2<F16>(o[4], o[5])), "r"(pack16x2<F16>(o[6], o[7]))
                     : "memory");
      }
      // input stage consumed; the staged gradient tile goes out with one tensor store (rows past M are clipped).  The
      // issuer first waits until the store of tile it-1 has finished reading its buffer, which tile it+1 overwrites.
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&empty[s]));
      if (ctid == 0) tc::tma_store_wait_read<0>();
      tc::named_bar_sync(1, 256);
      if (ctid == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tmOut)),
                     "r"(stg), "r"(0), "r"(t * kToTile)
                     : "memory");
        tc::tma_store_commit();
      }
    }
    if (ctid == 0) tc::tma_store_wait<0>();
#pragma unroll
    for (int e = 0; e < 8; ++e) { dw0[e] = warp_sum(dw0[e]); dw1[e] = warp_sum(dw1[e]); dw2[e] = warp_sum(dw2[e]); }
    if (w == 0) { db0 = warp_sum(db0); db1 = warp_sum(db1); db2 = warp_sum(db2); }
    if (lane < 8) {
      float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (e == lane) { a = dw0[e]; b = dw1[e]; c = dw2[e]; }
      atomicAdd(p.acc + w * 8 + lane, (double)a);
      atomicAdd(p.acc + 64 + w * 8 + lane, (double)b);
      atomicAdd(p.acc + 128 + w * 8 + lane, (double)c);
    }
    if (w == 0 && lane == 0) {
      atomicAdd(p.acc + 192, (double)db0);
      atomicAdd(p.acc + 193, (double)db1);
      atomicAdd(p.acc + 194, (double)db2);
    }
  }
}

// returns 0 = launched, 1 = not applicable, < 0 = error
int tail_dec1_bwd_tma(const float* dz4, const void* d2, int ldd2, void* dd2, int lddd2, const float* w1, double* acc, long long M,
                      bool f16, cudaStream_t st) {
  if (M < 4 * kToTile) return 1;
  const long long tiles = (M + kToTile - 1) / kToTile;
  if (tiles > 0x7fffffffLL) return 1;
  Dec1BwdParams p;
  p.w1 = w1; p.acc = acc; p.tiles = (int)tiles;
  CUtensorMap tmD2, tmDz, tmOut;
  {
    uint64_t dims[2] = {64ull, (uint64_t)M}, str[1] = {(uint64_t)ldd2 * 2};
    uint32_t box[2] = {64u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmD2, d2, 2, dims, str, box, 128)) return -1;
  }
  {
    uint64_t dims[2] = {8ull, (uint64_t)M}, str[1] = {16ull};
    uint32_t box[2] = {8u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmDz, dz4, 2, dims, str, box, 0)) return -1;
  }
  {
    uint64_t dims[2] = {64ull, (uint64_t)M}, str[1] = {(uint64_t)lddd2 * 2};
    uint32_t box[2] = {64u, (uint32_t)kToTile};
    if (tc::encode_tensor_map_bf16(&tmOut, dd2, 2, dims, str, box, 128)) return -1;
  }
  constexpr int SMEM = 1024 + kToStages * kToStage + 2 * kToMid;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(f16 ? tail_dec1_bwd_tma_kernel<true> : tail_dec1_bwd_tma_kernel<false>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "tail_dec1_bwd: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.tiles < 2 * kNumSMs ? p.tiles : 2 * kNumSMs;
  if (f16) tail_dec1_bwd_tma_kernel<true><<<grid, 288, SMEM, st>>>(tmD2, tmDz, tmOut, p);
  else tail_dec1_bwd_tma_kernel<false><<<grid, 288, SMEM, st>>>(tmD2, tmDz, tmOut, p);
  return check_launch("tail_dec1_bwd(tma)");
}

}  // namespace eunet
