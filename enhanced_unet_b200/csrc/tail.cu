// The 2Hx2W tail of EnhancedUNet (models.py:212 dec1, 236 final upsample, 308-313 enhance head,
// 337 residual) - bandwidth kernels around the generic 3x3 convolution:
//   z   = dec1(d2)                       1x1, computed at HxW (commutes with bilinear interpolation)
//   d1  = up(z)                          3 channels @ 2Hx2W, stored padded to 16 ch for the conv kernels
//   mid = conv3x3(d1; enhance.0)         generic conv kernel (+ BN batch statistics)
//   out = d1 + b3 + W3 . relu(bn(mid))   [B,3,2H,2W] fp32 NCHW (the model's output tensor)
// and the matching backward pieces.  64-channel pixels are processed by 8 lanes x 8 channels with
// a 3-step shuffle reduction; per-channel gradient sums go through shared memory + fp64 atomics.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

#define DISPATCH_DTYPE EUNET_DISPATCH_DTYPE

__device__ __forceinline__ float group8_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}

// bilinear x2 sample of the 3-channel z4 map at output pixel (oy, ox) of the 2Hx2W grid
__device__ __forceinline__ void up_sample3(const float* __restrict__ z4, int b, int oy, int ox, int H, int W, float d[3]) {
  const int k = oy >> 1, j = ox >> 1;
  int k0, k1, j0, j1;
  float wy0, wy1, wx0, wx1;
  if (oy & 1) { k0 = k; k1 = k < H - 1 ? k + 1 : k; wy0 = k < H - 1 ? 0.75f : 1.f; wy1 = k < H - 1 ? 0.25f : 0.f; }
  else        { k0 = k > 0 ? k - 1 : 0; k1 = k; wy0 = k > 0 ? 0.25f : 0.f; wy1 = k > 0 ? 0.75f : 1.f; }
  if (ox & 1) { j0 = j; j1 = j < W - 1 ? j + 1 : j; wx0 = j < W - 1 ? 0.75f : 1.f; wx1 = j < W - 1 ? 0.25f : 0.f; }
  else        { j0 = j > 0 ? j - 1 : 0; j1 = j; wx0 = j > 0 ? 0.25f : 0.f; wx1 = j > 0 ? 0.75f : 1.f; }
  const float4* zb = reinterpret_cast<const float4*>(z4) + (long long)b * H * W;
  const float4 a = __ldg(zb + (long long)k0 * W + j0), c = __ldg(zb + (long long)k0 * W + j1);
  const float4 e = __ldg(zb + (long long)k1 * W + j0), f = __ldg(zb + (long long)k1 * W + j1);
  d[0] = wy0 * (wx0 * a.x + wx1 * c.x) + wy1 * (wx0 * e.x + wx1 * f.x);
  d[1] = wy0 * (wx0 * a.y + wx1 * c.y) + wy1 * (wx0 * e.y + wx1 * f.y);
  d[2] = wy0 * (wx0 * a.z + wx1 * c.z) + wy1 * (wx0 * e.z + wx1 * f.z);
}

// ---- z = dec1(d2): 8 lanes per pixel ----
template <typename T>
__global__ void __launch_bounds__(256)
tail_dec1_fwd_kernel(const T* __restrict__ d2, int ld, const float* __restrict__ w1, const float* __restrict__ b1,
                     float* __restrict__ z4, long long M) {
  const int cg = threadIdx.x & 7;
  float w[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F8 t = load8(w1 + k * 64 + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) w[k][e] = t.v[e];
  }
  const float bb0 = b1[0], bb1 = b1[1], bb2 = b1[2];
  const long long stride = (long long)gridDim.x * 32;
  const long long Mr = (M + stride - 1) / stride * stride;   // keep whole warps in the loop (shuffles)
  for (long long p = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); p < Mr; p += stride) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (p < M) {
      const F8 v = load8(d2 + p * ld + cg * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s0 = fmaf(v.v[e], w[0][e], s0);
        s1 = fmaf(v.v[e], w[1][e], s1);
        s2 = fmaf(v.v[e], w[2][e], s2);
      }
    }
    s0 = group8_sum(s0); s1 = group8_sum(s1); s2 = group8_sum(s2);
    if (cg == 0 && p < M) reinterpret_cast<float4*>(z4)[p] = make_float4(s0 + bb0, s1 + bb1, s2 + bb2, 0.f);
  }
}

// ---- d1p = up(z) padded to 16 channels: one thread per output pixel ----
template <typename T>
__global__ void tail_up_fwd_kernel(const float* __restrict__ z4, T* __restrict__ d1p, float* __restrict__ d14, int B, int H, int W) {
  const int Wo = 2 * W, Ho = 2 * H;
  for (int row = blockIdx.x; row < B * Ho; row += gridDim.x)      // one output row per block iteration (32-bit indexing)
  for (int ox = threadIdx.x; ox < Wo; ox += blockDim.x) {
    const int oy = row % Ho, b = row / Ho;
    const long long i = (long long)row * Wo + ox;
    float d[3];
    up_sample3(z4, b, oy, ox, H, W, d);
    F8 lo, hi;
#pragma unroll
    for (int e = 0; e < 8; ++e) lo.v[e] = hi.v[e] = 0.f;
    lo.v[0] = d[0]; lo.v[1] = d[1]; lo.v[2] = d[2];
    store8(d1p + i * 16, lo);
    store8(d1p + i * 16 + 8, hi);
    if (d14 != nullptr) reinterpret_cast<float4*>(d14)[i] = make_float4(d[0], d[1], d[2], 0.f);   // fp32 copy for the residual
  }
}

// NCHW fp32 [B,3,Ho,Wo] -> pixel-major float4 (3 channels + pad): one 16-byte load per pixel for the tail kernels
__global__ void tail_pack3_kernel(const float* __restrict__ src, float* __restrict__ dst4, int B, long long HW,
                                  const float* __restrict__ gscale) {
  const float S = gscale_fwd(gscale);       // fp16 mode: power-of-two gradient scale (exact)
  for (int b = blockIdx.y; b < B; b += gridDim.y)
    for (long long hw = (long long)blockIdx.x * blockDim.x + threadIdx.x; hw < HW; hw += (long long)gridDim.x * blockDim.x) {
      const float* p = src + (long long)b * 3 * HW + hw;
      reinterpret_cast<float4*>(dst4)[(long long)b * HW + hw] = make_float4(S * __ldg(p), S * __ldg(p + HW), S * __ldg(p + 2 * HW), 0.f);
    }
}

constexpr int kTailStages = 4;   // 64-channel pixels (16 B per lane) in flight per thread

// ---- out = d1 + b3 + W3 . relu(mid*scale+shift) ----
// One thread per output pixel (no shuffles, coalesced NCHW stores); the 64-channel rows are staged block-wide
// through shared memory with cp.async (coalesced 16-byte chunks, 3 tiles of 128 pixels in flight) and read back
// with a padded pitch so every thread reads its own row conflict-free.  ~50 instructions per 16 bytes moved,
// i.e. below the ~80 at which a B200 SM stops being HBM-bound.
template <typename TY>
__global__ void __launch_bounds__(128)
tail_out_fwd_kernel(const float* __restrict__ d14, const TY* __restrict__ mid, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ w3, const float* __restrict__ b3,
                    float* __restrict__ out, int B, int H, int W) {
  constexpr int STAGES = 3, TP = 128;
  constexpr int ROWB = 64 * (int)sizeof(TY), PITCH = ROWB + 16, CPR = ROWB / 16;   // 16-byte chunks per row
  constexpr int STAGE_BYTES = TP * PITCH;
  extern __shared__ __align__(16) char dyn_smem[];
  __shared__ __align__(16) float cst[5][64];          // scale, shift, w3[0..2]
  __shared__ __align__(16) float4 dres[STAGES][TP];   // the residual d1 pixel of every thread
  const int tid = threadIdx.x;
  const int Wo = 2 * W, Ho = 2 * H;
  const long long HWo = (long long)Ho * Wo, M = (long long)B * HWo;
  const long long ntiles = (M + TP - 1) / TP;
  if (tid < 64) {
    cst[0][tid] = scale[tid]; cst[1][tid] = shift[tid];
    cst[2][tid] = w3[tid]; cst[3][tid] = w3[64 + tid]; cst[4][tid] = w3[128 + tid];
  }
  const float bb0 = b3[0], bb1 = b3[1], bb2 = b3[2];
  auto issue = [&](int stage, long long tile) {
    char* dst = dyn_smem + stage * STAGE_BYTES;
    const long long pbase = tile * TP;
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
      const int id = k * TP + tid, px = id / CPR, c = id % CPR;
      if (pbase + px < M) cp_async16(dst + px * PITCH + c * 16, reinterpret_cast<const char*>(mid + (pbase + px) * 64) + c * 16);
    }
    if (pbase + tid < M) cp_async16(&dres[stage][tid], d14 + (pbase + tid) * 4);
  };
  long long t = blockIdx.x;
#pragma unroll
  for (int i = 0; i < STAGES - 1; ++i) {
    if (t + (long long)i * gridDim.x < ntiles) issue(i, t + (long long)i * gridDim.x);
    cp_async_commit();
  }
  int it = 0;
  for (; t < ntiles; t += gridDim.x, ++it) {
    const long long tn = t + (long long)(STAGES - 1) * gridDim.x;
    if (tn < ntiles) issue((it + STAGES - 1) % STAGES, tn);
    cp_async_commit();
    cp_async_wait<STAGES - 1>();
    __syncthreads();                                  // tile `it` has landed for every thread (also publishes cst)
    const int stage = it % STAGES;
    const long long p = t * TP + tid;
    if (p < M) {
      const char* row = dyn_smem + stage * STAGE_BYTES + tid * PITCH;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const F8 v = load8(reinterpret_cast<const TY*>(row) + g * 8);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 sc = *reinterpret_cast<const float4*>(&cst[0][g * 8 + h * 4]), sh = *reinterpret_cast<const float4*>(&cst[1][g * 8 + h * 4]);
          const float4 w0 = *reinterpret_cast<const float4*>(&cst[2][g * 8 + h * 4]), w1 = *reinterpret_cast<const float4*>(&cst[3][g * 8 + h * 4]);
          const float4 w2 = *reinterpret_cast<const float4*>(&cst[4][g * 8 + h * 4]);
          const float a0 = fmaxf(fmaf(v.v[h * 4 + 0], sc.x, sh.x), 0.f), a1 = fmaxf(fmaf(v.v[h * 4 + 1], sc.y, sh.y), 0.f);
          const float a2 = fmaxf(fmaf(v.v[h * 4 + 2], sc.z, sh.z), 0.f), a3 = fmaxf(fmaf(v.v[h * 4 + 3], sc.w, sh.w), 0.f);
          s0 = fmaf(a0, w0.x, s0); s0 = fmaf(a1, w0.y, s0); s0 = fmaf(a2, w0.z, s0); s0 = fmaf(a3, w0.w, s0);
          s1 = fmaf(a0, w1.x, s1); s1 = fmaf(a1, w1.y, s1); s1 = fmaf(a2, w1.z, s1); s1 = fmaf(a3, w1.w, s1);
          s2 = fmaf(a0, w2.x, s2); s2 = fmaf(a1, w2.y, s2); s2 = fmaf(a2, w2.z, s2); s2 = fmaf(a3, w2.w, s2);
        }
      }
      const float4 d = dres[stage][tid];
      const long long b = p / HWo, hw = p - b * HWo;
      float* o = out + b * 3 * HWo + hw;
      o[0] = d.x + bb0 + s0;
      o[HWo] = d.y + bb1 + s1;
      o[2 * HWo] = d.z + bb2 + s2;
    }
    __syncthreads();                                  // everyone is done with this stage before it is refilled
  }
  cp_async_wait<0>();
}

// Shared helper: reduce per-thread arrays (thread = 8 channels of group cg, 32 pixel rows per block)
// over the block's pixel rows and atomically add to fp64 accumulators acc[base + cg*8 + e].
__device__ __forceinline__ void block_reduce64(const float v[8], float (*red)[64], int cg, int row, double* acc_base) {
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 8; ++e) red[row][cg * 8 + e] = v[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(acc_base + threadIdx.x, (double)s);
  }
}

// ---- backward pass 1 over (dout, mid): BN sums + enhance.3 weight/bias gradients ----
// acc layout: [0,64) sum_g, [64,128) sum_g*xhat, [128,320) dW3[k][c], [320,323) db3[k]
template <typename TY>
__global__ void __launch_bounds__(256, 2)
tail_bwd_reduce_kernel(const float* __restrict__ dout, const TY* __restrict__ mid, const float* __restrict__ scale,
                       const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                       const float* __restrict__ w3, double* __restrict__ acc, int B, int H, int W) {
  __shared__ float red[32][64];
  const int cg = threadIdx.x & 7, row = threadIdx.x >> 3;
  const int Wo = 2 * W, Ho = 2 * H;
  const long long HWo = (long long)Ho * Wo, M = (long long)B * HWo;
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8), mu = load8(mean + cg * 8), is = load8(invstd + cg * 8);
  float w[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F8 t = load8(w3 + k * 64 + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) w[k][e] = t.v[e];
  }
  // Issue-bound kernel (ncu: issue_active 80 %, DRAM 41 %), so the inner loop keeps only what depends on the pixel:
  // with xc = v - mean, pre = sc * xc + beta and m = [pre > 0], the masked sums
  //     P_k[c] = sum_p m * g_k          Q_k[c] = sum_p m * g_k * xc
  // (predicated adds / fmas: 9 instructions per element instead of 14) give all five reductions afterwards:
  //     sum g' = sum_k w_k P_k,   sum g' xhat = invstd * sum_k w_k Q_k,   dW3[k] = sum_p g_k relu(pre) = sc Q_k + beta P_k
  // (centring on the batch mean keeps sc Q_k + beta P_k free of cancellation).
  float beta[8], P[3][8], Q[3][8], db[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    beta[e] = fmaf(sc.v[e], mu.v[e], sh.v[e]);
    P[0][e] = P[1][e] = P[2][e] = Q[0][e] = Q[1][e] = Q[2][e] = 0.f;
  }
  extern __shared__ __align__(16) char dyn_smem[];
  Stream8<TY, kTailStages> sm(dyn_smem, 256);
  Stream4f<kTailStages> sdo(dyn_smem + Stream8<TY, kTailStages>::bytes(256), 256);
  const long long p0 = (long long)blockIdx.x * 32 + row, step = (long long)gridDim.x * 32;
  const long long n = p0 < M ? (M - p0 + step - 1) / step : 0;
#pragma unroll
  for (int i = 0; i < kTailStages - 1; ++i) {
    if (i < n) {
      sm.issue(i, mid + (p0 + i * step) * 64 + cg * 8);
      sdo.issue(i, dout + (p0 + i * step) * 4);
    }
    cp_async_commit();
  }
  for (long long i = 0; i < n; ++i) {
    const long long j = i + kTailStages - 1;
    if (j < n) {
      sm.issue((int)(j % kTailStages), mid + (p0 + j * step) * 64 + cg * 8);
      sdo.issue((int)(j % kTailStages), dout + (p0 + j * step) * 4);
    }
    cp_async_commit();
    cp_async_wait<kTailStages - 1>();
    const float4 gq = sdo.get((int)(i % kTailStages));
    const float g0 = gq.x, g1 = gq.y, g2 = gq.z;
    const F8 v = sm.get((int)(i % kTailStages));
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xc = v.v[e] - mu.v[e];
      if (fmaf(xc, sc.v[e], beta[e]) > 0.f) {
        P[0][e] += g0; P[1][e] += g1; P[2][e] += g2;
        Q[0][e] = fmaf(g0, xc, Q[0][e]); Q[1][e] = fmaf(g1, xc, Q[1][e]); Q[2][e] = fmaf(g2, xc, Q[2][e]);
      }
    }
    if (cg == 0) { db[0] += g0; db[1] += g1; db[2] += g2; }
  }
  float sg[8], sgx[8], dw[3][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[e] = w[0][e] * P[0][e] + w[1][e] * P[1][e] + w[2][e] * P[2][e];
    sgx[e] = is.v[e] * (w[0][e] * Q[0][e] + w[1][e] * Q[1][e] + w[2][e] * Q[2][e]);
#pragma unroll
    for (int k = 0; k < 3; ++k) dw[k][e] = fmaf(sc.v[e], Q[k][e], beta[e] * P[k][e]);
  }
  cp_async_wait<0>();
  block_reduce64(sg, red, cg, row, acc);
  block_reduce64(sgx, red, cg, row, acc + 64);
  block_reduce64(dw[0], red, cg, row, acc + 128);
  block_reduce64(dw[1], red, cg, row, acc + 192);
  block_reduce64(dw[2], red, cg, row, acc + 256);
  // bias gradient: rows' lane-0 partials
  __syncthreads();
  if (cg == 0) { red[row][0] = db[0]; red[row][1] = db[1]; red[row][2] = db[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(acc + 320 + threadIdx.x, (double)s);
  }
}

// ---- backward pass 2: dmid = scale * (g - mean(g) - xhat * mean(g*xhat)) ----
template <typename T, typename TY>
__global__ void __launch_bounds__(256)
tail_bwd_dmid_kernel(const float* __restrict__ dout, const TY* __restrict__ mid, T* __restrict__ dmid,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ w3, const double* __restrict__ acc, int B,
                     int H, int W) {
  const int cg = threadIdx.x & 7, row = threadIdx.x >> 3;
  const int Wo = 2 * W, Ho = 2 * H;
  const long long HWo = (long long)Ho * Wo, M = (long long)B * HWo;
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8), mu = load8(mean + cg * 8), is = load8(invstd + cg * 8);
  // dmid = sc * (g' - k1 - xhat * k2) with g' = [pre > 0] * sum_k g_k w_k, xhat = (v - mean) * invstd, rewritten as
  //     dmid = A + Bc v + [sc v + sh > 0] * sum_k g_k * (sc w_k),   Bc = -sc k2 invstd, A = -sc k1 - Bc mean
  // (7 instructions per element; the same expression, in the same order, as the transform stage of
  // tail_bwd_fused_kernel, so both paths produce bit-identical bf16 gradients)
  float ws[3][8], A[8], Bc[8];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F8 t = load8(w3 + k * 64 + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) ws[k][e] = t.v[e] * sc.v[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float k1 = (float)(acc[cg * 8 + e] / (double)M), k2 = (float)(acc[64 + cg * 8 + e] / (double)M);
    Bc[e] = -sc.v[e] * k2 * is.v[e];
    A[e] = -sc.v[e] * k1 - Bc[e] * mu.v[e];
  }
  extern __shared__ __align__(16) char dyn_smem[];
  Stream8<TY, kTailStages> sm(dyn_smem, 256);
  Stream4f<kTailStages> sdo(dyn_smem + Stream8<TY, kTailStages>::bytes(256), 256);
  const long long p0 = (long long)blockIdx.x * 32 + row, step = (long long)gridDim.x * 32;
  const long long n = p0 < M ? (M - p0 + step - 1) / step : 0;
#pragma unroll
  for (int i = 0; i < kTailStages - 1; ++i) {
    if (i < n) {
      sm.issue(i, mid + (p0 + i * step) * 64 + cg * 8);
      sdo.issue(i, dout + (p0 + i * step) * 4);
    }
    cp_async_commit();
  }
  for (long long i = 0; i < n; ++i) {
    const long long j = i + kTailStages - 1;
    if (j < n) {
      sm.issue((int)(j % kTailStages), mid + (p0 + j * step) * 64 + cg * 8);
      sdo.issue((int)(j % kTailStages), dout + (p0 + j * step) * 4);
    }
    cp_async_commit();
    const long long p = p0 + i * step;
    cp_async_wait<kTailStages - 1>();
    const float4 gq = sdo.get((int)(i % kTailStages));
    const float g0 = gq.x, g1 = gq.y, g2 = gq.z;
    const F8 v = sm.get((int)(i % kTailStages));
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float r = fmaf(Bc[e], v.v[e], A[e]);
      if (fmaf(v.v[e], sc.v[e], sh.v[e]) > 0.f) r = fmaf(g2, ws[2][e], fmaf(g1, ws[1][e], fmaf(g0, ws[0][e], r)));
      o.v[e] = r;
    }
    store8(dmid + p * 64 + cg * 8, o);
  }
  cp_async_wait<0>();
}

// ---- dz = up^T(dd1[:3] + dout): one thread per HxW pixel gathers the 4x4 output neighbourhood ----
template <typename T>
__device__ __forceinline__ void load3(const T* p, float v[3]);
template <>
__device__ __forceinline__ void load3<float>(const float* p, float v[3]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z;
}
template <>
__device__ __forceinline__ void load3<__nv_bfloat16>(const __nv_bfloat16* p, float v[3]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16);
  v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16);
}

template <>
__device__ __forceinline__ void load3<__half>(const __half* p, float v[3]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x;
}

template <typename T>
__global__ void tail_up_bwd_kernel(const T* __restrict__ dd1p, int stride, const float* __restrict__ dout,
                                   float* __restrict__ dz4, int B, int H, int W, const float* __restrict__ gscale) {
  const float S = gscale_fwd(gscale);       // dd1p carries the gradient scale already, dout (the caller's tensor) does not
  const int Wo = 2 * W, Ho = 2 * H;
  const long long HWo = (long long)Ho * Wo;
  for (int row = blockIdx.x; row < B * H; row += gridDim.x)
  for (int j = threadIdx.x; j < W; j += blockDim.x) {
    const int k = row % H, b = row / H;
    const long long i = (long long)row * W + j;
    const float wy[4] = {k > 0 ? 0.25f : 0.f, k > 0 ? 0.75f : 1.f, k < H - 1 ? 0.75f : 1.f, k < H - 1 ? 0.25f : 0.f};
    const float wx[4] = {j > 0 ? 0.25f : 0.f, j > 0 ? 0.75f : 1.f, j < W - 1 ? 0.75f : 1.f, j < W - 1 ? 0.25f : 0.f};
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (wy[r] == 0.f) continue;
      const int oy = 2 * k - 1 + r;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        if (wx[s] == 0.f) continue;
        const int ox = 2 * j - 1 + s;
        const long long op = (long long)b * HWo + (long long)oy * Wo + ox;
        float v[3];
        load3<T>(dd1p + op * stride, v);
        const float* dp = dout + (long long)b * 3 * HWo + (long long)oy * Wo + ox;
        const float wgt = wy[r] * wx[s];
        a0 += wgt * fmaf(S, __ldg(dp), v[0]);
        a1 += wgt * fmaf(S, __ldg(dp + HWo), v[1]);
        a2 += wgt * fmaf(S, __ldg(dp + 2 * HWo), v[2]);
      }
    }
    reinterpret_cast<float4*>(dz4)[i] = make_float4(a0, a1, a2, 0.f);
  }
}

// ---- dec1 backward: dd2 = dz . W1 ; dW1 += dz^T d2 ; db1 += dz.  acc: [0,192) dW1[k][c], [192,195) db1 ----
template <typename T>
__global__ void __launch_bounds__(256)
tail_dec1_bwd_kernel(const float* __restrict__ dz4, const T* __restrict__ d2, int ldd2, T* __restrict__ dd2, int lddd2,
                     const float* __restrict__ w1, double* __restrict__ acc, long long M) {
  __shared__ float red[32][64];
  const int cg = threadIdx.x & 7, row = threadIdx.x >> 3;
  float w[3][8], dw[3][8], db[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const F8 t = load8(w1 + k * 64 + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) { w[k][e] = t.v[e]; dw[k][e] = 0.f; }
  }
  for (long long p = (long long)blockIdx.x * 32 + row; p < M; p += (long long)gridDim.x * 32) {
    const float4 dz = __ldg(reinterpret_cast<const float4*>(dz4) + p);
    const F8 v = load8(d2 + p * ldd2 + cg * 8);
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      o.v[e] = dz.x * w[0][e] + dz.y * w[1][e] + dz.z * w[2][e];
      dw[0][e] = fmaf(dz.x, v.v[e], dw[0][e]);
      dw[1][e] = fmaf(dz.y, v.v[e], dw[1][e]);
      dw[2][e] = fmaf(dz.z, v.v[e], dw[2][e]);
    }
    store8(dd2 + p * lddd2 + cg * 8, o);
    if (cg == 0) { db[0] += dz.x; db[1] += dz.y; db[2] += dz.z; }
  }
  block_reduce64(dw[0], red, cg, row, acc);
  block_reduce64(dw[1], red, cg, row, acc + 64);
  block_reduce64(dw[2], red, cg, row, acc + 128);
  __syncthreads();
  if (cg == 0) { red[row][0] = db[0]; red[row][1] = db[1]; red[row][2] = db[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(acc + 192 + threadIdx.x, (double)s);
  }
}

static void tail_ring_attr(const void* kernel, int bytes) {
  if (bytes > 32 * 1024) (void)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
static inline int rows_grid(long long M) { return clamp_grid((M + 31) / 32, 8); }
static inline int rows_grid4(long long M) { return clamp_grid((M + 127) / 128, 8); }
static inline int ew_grid(long long items) { return clamp_grid((items + 255) / 256, 16); }

}  // namespace eunet

namespace eunet {
int g_opt_tail_out_tma = 1;
int tail_out_fwd_tma(const float* d14, const void* mid, const float* scale, const float* shift, const float* w3, const float* b3,
                     float* out, int B, int H, int W, cudaStream_t st);
int tail_dec1_fwd_tma(const void* d2, int ldd2, const float* w1, const float* b1, float* z4, long long M, bool f16,
                      cudaStream_t st);
int tail_dec1_bwd_tma(const float* dz4, const void* d2, int ldd2, void* dd2, int lddd2, const float* w1, double* acc, long long M,
                      bool f16, cudaStream_t st);
int tail_bwd_reduce_tma(const float* dout4, const void* mid, const float* scale, const float* shift, const float* mean,
                        const float* invstd, const float* w3, double* acc, int B, int H, int W, cudaStream_t st);
}  // namespace eunet

using namespace eunet;

extern "C" {

int eunet_tail_dec1_fwd(const void* d2, int ldd2, int dtype, const float* w1, const float* b1, float* z4, long long M,
                        void* stream) {
  EUNET_REQUIRE(M > 0 && ldd2 >= 64 && (ldd2 & 7) == 0, "tail_dec1_fwd: bad shape M=%lld ld=%d", M, ldd2);
  if ((dtype == EUNET_BF16 || dtype == EUNET_F16) && g_opt_tail_out_tma) {
    const int rc = tail_dec1_fwd_tma(d2, ldd2, w1, b1, z4, M, dtype == EUNET_F16, (cudaStream_t)stream);
    if (rc <= 0) return rc;      // launched or failed; 1 = too small, 8-lanes-per-pixel kernel below
  }
  DISPATCH_DTYPE(dtype, tail_dec1_fwd_kernel<T><<<rows_grid(M), 256, 0, (cudaStream_t)stream>>>((const T*)d2, ldd2, w1, b1, z4, M));
  return check_launch("tail_dec1_fwd");
}

int eunet_tail_pack3(const float* src, float* dst4, int B, int H, int W, const float* gscale, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_pack3: bad shape");
  {
    const long long HW = (long long)H * W;
    long long bx = (HW + 255) / 256, cap = ((long long)kNumSMs * 8 + B - 1) / B;
    dim3 grid((unsigned)(bx < cap ? bx : cap), (unsigned)(B < 65535 ? B : 65535));
    tail_pack3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst4, B, HW, gscale);
  }
  return check_launch("tail_pack3");
}

int eunet_tail_up_fwd(const float* z4, void* d1p, float* d14, int dtype, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_up_fwd: bad shape");
  DISPATCH_DTYPE(dtype, tail_up_fwd_kernel<T><<<clamp_grid(2LL * B * H, 16), 256, 0, (cudaStream_t)stream>>>(z4, (T*)d1p, d14, B, H, W));
  return check_launch("tail_up_fwd");
}

int eunet_tail_out_fwd(const float* d14, const void* mid, int dtype, const float* scale, const float* shift, const float* w3,
                       const float* b3, float* out, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_out_fwd: bad shape");
  if ((dtype == EUNET_BF16 || dtype == EUNET_F16) && g_opt_tail_out_tma) {
    const int rc = tail_out_fwd_tma(d14, mid, scale, shift, w3, b3, out, B, H, W, (cudaStream_t)stream);
    if (rc <= 0) return rc;      // launched or failed; 1 = too small, use the thread-per-pixel kernel
  }
  DISPATCH_DTYPE(dtype, tail_ring_attr((const void*)tail_out_fwd_kernel<TY>, 3 * 128 * (64 * (int)sizeof(TY) + 16));
                 tail_out_fwd_kernel<TY><<<clamp_grid((4LL * B * H * W + 127) / 128, 4), 128, 3 * 128 * (64 * (int)sizeof(TY) + 16), (cudaStream_t)stream>>>(
                            d14, (const TY*)mid, scale, shift, w3, b3, out, B, H, W));
  return check_launch("tail_out_fwd");
}

int eunet_tail_bwd_reduce(const float* dout, const void* mid, int dtype, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const float* w3, double* acc, int B, int H, int W,
                          void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_bwd_reduce: bad shape");
  if ((dtype == EUNET_BF16 || dtype == EUNET_F16) && g_opt_tail_out_tma) {
    const int rc = tail_bwd_reduce_tma(dout, mid, scale, shift, mean, invstd, w3, acc, B, H, W, (cudaStream_t)stream);
    if (rc <= 0) return rc;      // launched or failed; 1 = too small, use the ring kernel
  }
  DISPATCH_DTYPE(dtype, tail_ring_attr((const void*)tail_bwd_reduce_kernel<TY>, Stream8<TY, kTailStages>::bytes(256) + Stream4f<kTailStages>::bytes(256));
                 tail_bwd_reduce_kernel<TY><<<rows_grid(4LL * B * H * W), 256, Stream8<TY, kTailStages>::bytes(256) + Stream4f<kTailStages>::bytes(256), (cudaStream_t)stream>>>(
                            dout, (const TY*)mid, scale, shift, mean, invstd, w3, acc, B, H, W));
  return check_launch("tail_bwd_reduce");
}

int eunet_tail_bwd_dmid(const float* dout, const void* mid, void* dmid, int dtype, const float* scale, const float* shift,
                        const float* mean, const float* invstd, const float* w3, const double* acc, int B, int H, int W,
                        void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_bwd_dmid: bad shape");
  DISPATCH_DTYPE(dtype, tail_ring_attr((const void*)tail_bwd_dmid_kernel<T, TY>, Stream8<TY, kTailStages>::bytes(256) + Stream4f<kTailStages>::bytes(256));
                 tail_bwd_dmid_kernel<T, TY><<<rows_grid(4LL * B * H * W), 256, Stream8<TY, kTailStages>::bytes(256) + Stream4f<kTailStages>::bytes(256), (cudaStream_t)stream>>>(
                            dout, (const TY*)mid, (T*)dmid, scale, shift, mean, invstd, w3, acc, B, H, W));
  return check_launch("tail_bwd_dmid");
}

int eunet_tail_up_bwd(const void* dd1p, int dtype, int dd1_stride, const float* dout, float* dz4, int B, int H, int W,
                      const float* gscale, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "tail_up_bwd: bad shape");
  EUNET_REQUIRE(dd1_stride >= 4 && dd1_stride % 4 == 0, "tail_up_bwd: dd1_stride %d must be a multiple of 4 (>= 4)", dd1_stride);
  DISPATCH_DTYPE(dtype, tail_up_bwd_kernel<T><<<clamp_grid((long long)B * H, 16), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)dd1p, dd1_stride, dout, dz4, B, H, W, gscale));
  return check_launch("tail_up_bwd");
}

int eunet_tail_dec1_bwd(const float* dz4, const void* d2, int ldd2, void* dd2, int lddd2, int dtype, const float* w1,
                        double* acc, long long M, void* stream) {
  EUNET_REQUIRE(M > 0 && ldd2 >= 64 && lddd2 >= 64, "tail_dec1_bwd: bad shape");
  if ((dtype == EUNET_BF16 || dtype == EUNET_F16) && g_opt_tail_out_tma && (ldd2 & 7) == 0 && (lddd2 & 7) == 0) {
    const int rc = tail_dec1_bwd_tma(dz4, d2, ldd2, dd2, lddd2, w1, acc, M, dtype == EUNET_F16, (cudaStream_t)stream);
    if (rc <= 0) return rc;      // launched or failed; 1 = too small, 8-lanes-per-pixel kernel below
  }
  DISPATCH_DTYPE(dtype, tail_dec1_bwd_kernel<T><<<rows_grid(M), 256, 0, (cudaStream_t)stream>>>(dz4, (const T*)d2, ldd2, (T*)dd2,
                                                                                               lddd2, w1, acc, M));
  return check_launch("tail_dec1_bwd");
}

}  // extern "C"
