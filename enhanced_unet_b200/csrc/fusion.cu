// In-file fusion blocks of the reference's smp body (models.py:276-302, 320-328), eval mode:
//   f = cat[out_main, out_aux] (6 ch);  f *= sigmoid(bn(conv1x1(gelu(bn(conv3x3(f))))))      (attention gate)
//   out = fusion_head(f) + fusion_residual(f)
// The 3x3 convolutions of the head (6->256->128->64) run on the generic tensor-core kernels; this file holds the
// bandwidth-bound pieces: the gate (162+18 MAC/pixel, CUDA cores) fused with the concat, the gating multiply,
// the residual 1x1 and the NHWC16 packing - and the final "head 1x1 + residual" store to NCHW fp32.
#include <string.h>
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct FusionGateParams {
  float w0[3 * 6 * 9];   // attention_gate.0.weight [3][6][3][3]
  float s1[3], h1[3];    // attention_gate.1 folded (eval) scale / shift
  float w3[6 * 3];       // attention_gate.3.weight [6][3]
  float s4[6], h4[6];    // attention_gate.4 folded scale / shift
  float wr[3 * 6];       // fusion_residual.weight [3][6]
  float br[3];           // fusion_residual.bias
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <typename T>
__global__ void __launch_bounds__(256)
fusion_gate_kernel(const float* __restrict__ main_, const float* __restrict__ aux, const __grid_constant__ FusionGateParams p,
                   T* __restrict__ fg16, float* __restrict__ res4, int B, int H, int W) {
  const long long HW = (long long)H * W, M = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H);
    const long long b = i / HW;
    const float* pm = main_ + b * 3 * HW;
    const float* pa = aux + b * 3 * HW;
    float a3[3] = {0.f, 0.f, 0.f};
    float f[6];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
        const long long o = (long long)yy * W + xx;
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const float v = in ? (c < 3 ? __ldg(pm + c * HW + o) : __ldg(pa + (c - 3) * HW + o)) : 0.f;
          if (ky == 1 && kx == 1) f[c] = v;
#pragma unroll
          for (int k = 0; k < 3; ++k) a3[k] = fmaf(v, p.w0[(k * 6 + c) * 9 + ky * 3 + kx], a3[k]);
        }
      }
    float g3[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g3[k] = gelu_erf(fmaf(a3[k], p.s1[k], p.h1[k]));
    F8 lo, hi;
#pragma unroll
    for (int e = 0; e < 8; ++e) lo.v[e] = hi.v[e] = 0.f;
    float fgv[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) t = fmaf(g3[k], p.w3[c * 3 + k], t);
      t = fmaf(t, p.s4[c], p.h4[c]);
      const float gate = 1.f / (1.f + expf(-t));
      fgv[c] = f[c] * gate;
      lo.v[c] = fgv[c];
    }
    store8(fg16 + i * 16, lo);
    store8(fg16 + i * 16 + 8, hi);
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float t = p.br[k];
#pragma unroll
      for (int c = 0; c < 6; ++c) t = fmaf(fgv[c], p.wr[k * 6 + c], t);
      r[k] = t;
    }
    reinterpret_cast<float4*>(res4)[i] = make_float4(r[0], r[1], r[2], 0.f);
  }
}

__global__ void fusion_out_kernel(const float* __restrict__ z4, const float* __restrict__ res4, float* __restrict__ out, int B,
                                  long long HW) {
  const long long M = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(z4)[i], r = reinterpret_cast<const float4*>(res4)[i];
    const long long b = i / HW, hw = i % HW;
    float* o = out + b * 3 * HW + hw;
    o[0] = a.x + r.x;
    o[HW] = a.y + r.y;
    o[2 * HW] = a.z + r.z;
  }
}


// ================================================================================================================
// TRAINING mode of the attention gate (batch-statistics BatchNorm) and its backward.  Three forward passes (each BN needs
// the statistics of the whole batch before it can be applied) and four backward passes; everything is fp32 CUDA-core
// work on 3 / 6 channels per pixel.  Layouts: a1 [M][4] (3 real), a2 [M][8] (6 real), dz1 [M][4], dz2 [M][8] fp32.
// ================================================================================================================
struct GateW {
  float w0[3 * 6 * 9];   // attention_gate.0.weight [3][6][3][3]
  float w3[6 * 3];       // attention_gate.3.weight [6][3]
  float wr[3 * 6];       // fusion_residual.weight [3][6]
  float br[3];           // fusion_residual.bias
};

__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
  return cdf + x * 0.39894228040143267794f * expf(-0.5f * x * x);
}

__device__ __forceinline__ void load_f6(const float* __restrict__ pm, const float* __restrict__ pa, long long HW, long long o, float f[6]) {
#pragma unroll
  for (int c = 0; c < 3; ++c) { f[c] = __ldg(pm + c * HW + o); f[3 + c] = __ldg(pa + c * HW + o); }
}

// block-wide sum of N per-thread values into fp64 accumulators (one atomic per value and block)
template <int N>
__device__ __forceinline__ void block_accumulate(float (&v)[N], double* __restrict__ acc) {
  __shared__ float red[8][N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w][i];
    atomicAdd(acc + i, (double)s);
  }
  __syncthreads();
}

// G1: a1 = conv3x3(cat[main, aux]; w0) (no bias) + batch statistics {sum[3], sumsq[3]}
__global__ void __launch_bounds__(256)
gate_conv_fwd_kernel(const float* __restrict__ main_, const float* __restrict__ aux, const __grid_constant__ GateW p,
                     float* __restrict__ a1, double* __restrict__ stats, int B, int H, int W) {
  const long long HW = (long long)H * W, M = (long long)B * HW;
  float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const long long nthr = (long long)gridDim.x * blockDim.x, Mr = (M + nthr - 1) / nthr * nthr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Mr; i += nthr) {
    if (i < M) {
      const int x = (int)(i % W), y = (int)((i / W) % H);
      const long long b = i / HW;
      const float* pm = main_ + b * 3 * HW;
      const float* pa = aux + b * 3 * HW;
      float a3[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = y + ky - 1, xx = x + kx - 1;
          if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
          float f[6];
          load_f6(pm, pa, HW, (long long)yy * W + xx, f);
#pragma unroll
          for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) a3[k] = fmaf(f[c], p.w0[(k * 6 + c) * 9 + ky * 3 + kx], a3[k]);
        }
      reinterpret_cast<float4*>(a1)[i] = make_float4(a3[0], a3[1], a3[2], 0.f);
#pragma unroll
      for (int k = 0; k < 3; ++k) { st[k] += a3[k]; st[3 + k] = fmaf(a3[k], a3[k], st[3 + k]); }
    }
  }
  block_accumulate<6>(st, stats);
}

// G2: a2 = conv1x1(gelu(bn1(a1)); w3) (no bias) + batch statistics {sum[6], sumsq[6]}
__global__ void __launch_bounds__(256)
gate_mid_fwd_kernel(const float* __restrict__ a1, const float* __restrict__ sc1, const float* __restrict__ sh1,
                    const __grid_constant__ GateW p, float* __restrict__ a2, double* __restrict__ stats, long long M) {
  float st[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) st[i] = 0.f;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += nthr) {
    const float4 v = reinterpret_cast<const float4*>(a1)[i];
    const float a[3] = {v.x, v.y, v.z};
    float g[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = gelu_erf(fmaf(a[k], sc1[k], sh1[k]));
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) t = fmaf(g[k], p.w3[c * 3 + k], t);
      o[c] = t;
      st[c] += t;
      st[6 + c] = fmaf(t, t, st[6 + c]);
    }
    reinterpret_cast<float4*>(a2)[2 * i] = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(a2)[2 * i + 1] = make_float4(o[4], o[5], 0.f, 0.f);
  }
  block_accumulate<12>(st, stats);
}

// G3: gate = sigmoid(bn2(a2)); fg = f * gate -> fg16 (activation dtype, 16 channels) and res4 = fusion_residual(fg)
template <typename T>
__global__ void __launch_bounds__(256)
gate_apply_fwd_kernel(const float* __restrict__ main_, const float* __restrict__ aux, const float* __restrict__ a2,
                      const float* __restrict__ sc2, const float* __restrict__ sh2, const __grid_constant__ GateW p,
                      T* __restrict__ fg16, float* __restrict__ res4, int B, long long HW) {
  const long long M = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, o = i % HW;
    float f[6];
    load_f6(main_ + b * 3 * HW, aux + b * 3 * HW, HW, o, f);
    const float4 q0 = reinterpret_cast<const float4*>(a2)[2 * i], q1 = reinterpret_cast<const float4*>(a2)[2 * i + 1];
    const float av[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
    F8 lo, hi;
#pragma unroll
    for (int e = 0; e < 8; ++e) lo.v[e] = hi.v[e] = 0.f;
    float fg[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const float gate = 1.f / (1.f + expf(-fmaf(av[c], sc2[c], sh2[c])));
      fg[c] = f[c] * gate;
      lo.v[c] = fg[c];
    }
    store8(fg16 + i * 16, lo);
    store8(fg16 + i * 16 + 8, hi);
    float r[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float t = p.br[k];
#pragma unroll
      for (int c = 0; c < 6; ++c) t = fmaf(fg[c], p.wr[k * 6 + c], t);
      r[k] = t;
    }
    reinterpret_cast<float4*>(res4)[i] = make_float4(r[0], r[1], r[2], 0.f);
  }
}

// B1: dfg = dfg16[:6] + wr^T dout;  dz2 = dfg * f * gate (1 - gate)  -> dz2 [M][8];
//     acc: [0,6) sum dz2, [6,12) sum dz2 * xhat2, [12,30) dWr[k][c] = sum dout[k] fg[c], [30,33) dbr[k] = sum dout[k]
template <typename T>
__global__ void __launch_bounds__(256)
gate_bwd1_kernel(const float* __restrict__ main_, const float* __restrict__ aux, const float* __restrict__ a2,
                 const float* __restrict__ sc2, const float* __restrict__ sh2, const float* __restrict__ mean2,
                 const float* __restrict__ invstd2, const T* __restrict__ dfg16, const float* __restrict__ dout4,
                 const __grid_constant__ GateW p, float* __restrict__ dz2, double* __restrict__ acc, int B, long long HW) {
  const long long M = (long long)B * HW;
  float st[33];
#pragma unroll
  for (int i = 0; i < 33; ++i) st[i] = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, o = i % HW;
    float f[6];
    load_f6(main_ + b * 3 * HW, aux + b * 3 * HW, HW, o, f);
    const float4 q0 = reinterpret_cast<const float4*>(a2)[2 * i], q1 = reinterpret_cast<const float4*>(a2)[2 * i + 1];
    const float av[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
    const F8 dv = load8(dfg16 + i * 16);
    const float4 dq = reinterpret_cast<const float4*>(dout4)[i];
    const float d3[3] = {dq.x, dq.y, dq.z};
    float dz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const float gate = 1.f / (1.f + expf(-fmaf(av[c], sc2[c], sh2[c])));
      float dfg = dv.v[c];
#pragma unroll
      for (int k = 0; k < 3; ++k) dfg = fmaf(p.wr[k * 6 + c], d3[k], dfg);
      dz[c] = dfg * f[c] * gate * (1.f - gate);
      st[c] += dz[c];
      st[6 + c] = fmaf(dz[c], (av[c] - mean2[c]) * invstd2[c], st[6 + c]);
      const float fg = f[c] * gate;
#pragma unroll
      for (int k = 0; k < 3; ++k) st[12 + k * 6 + c] = fmaf(d3[k], fg, st[12 + k * 6 + c]);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) st[30 + k] += d3[k];
    reinterpret_cast<float4*>(dz2)[2 * i] = make_float4(dz[0], dz[1], dz[2], dz[3]);
    reinterpret_cast<float4*>(dz2)[2 * i + 1] = make_float4(dz[4], dz[5], 0.f, 0.f);
  }
  block_accumulate<33>(st, acc);
}

// B2: da2 = sc2 (dz2 - k1 - xhat2 k2);  dW3[c][k] += da2[c] g1[k];  dz1[k] = gelu'(z1[k]) sum_c w3[c][k] da2[c] -> dz1 [M][4];
//     acc: [0,3) sum dz1, [3,6) sum dz1 * xhat1, [6,24) dW3[c][k]
__global__ void __launch_bounds__(256)
gate_bwd2_kernel(const float* __restrict__ a1, const float* __restrict__ a2, const float* __restrict__ dz2,
                 const float* __restrict__ sc1, const float* __restrict__ sh1, const float* __restrict__ mean1,
                 const float* __restrict__ invstd1, const float* __restrict__ sc2, const float* __restrict__ mean2,
                 const float* __restrict__ invstd2, const double* __restrict__ acc2, const __grid_constant__ GateW p,
                 float* __restrict__ dz1, double* __restrict__ acc, long long M) {
  float st[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) st[i] = 0.f;
  float k1[6], k2[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) { k1[c] = (float)(acc2[c] / (double)M); k2[c] = (float)(acc2[6 + c] / (double)M); }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(a1)[i];
    const float a[3] = {v.x, v.y, v.z};
    const float4 q0 = reinterpret_cast<const float4*>(a2)[2 * i], q1 = reinterpret_cast<const float4*>(a2)[2 * i + 1];
    const float av[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
    const float4 z0 = reinterpret_cast<const float4*>(dz2)[2 * i], z1q = reinterpret_cast<const float4*>(dz2)[2 * i + 1];
    const float dz[6] = {z0.x, z0.y, z0.z, z0.w, z1q.x, z1q.y};
    float z1[3], g1[3], dg1[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; ++k) { z1[k] = fmaf(a[k], sc1[k], sh1[k]); g1[k] = gelu_erf(z1[k]); }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const float da2 = sc2[c] * (dz[c] - k1[c] - (av[c] - mean2[c]) * invstd2[c] * k2[c]);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        st[6 + c * 3 + k] = fmaf(da2, g1[k], st[6 + c * 3 + k]);
        dg1[k] = fmaf(p.w3[c * 3 + k], da2, dg1[k]);
      }
    }
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k] = dg1[k] * gelu_erf_grad(z1[k]);
      st[k] += o[k];
      st[3 + k] = fmaf(o[k], (a[k] - mean1[k]) * invstd1[k], st[3 + k]);
    }
    reinterpret_cast<float4*>(dz1)[i] = make_float4(o[0], o[1], o[2], 0.f);
  }
  block_accumulate<24>(st, acc);
}

// B3: da1 = sc1 (dz1 - k1 - xhat1 k2) -> overwrites dz1 in place
__global__ void __launch_bounds__(256)
gate_bwd3_kernel(const float* __restrict__ a1, float* __restrict__ dz1, const float* __restrict__ sc1, const float* __restrict__ mean1,
                 const float* __restrict__ invstd1, const double* __restrict__ acc1, long long M) {
  float k1[3], k2[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { k1[k] = (float)(acc1[k] / (double)M); k2[k] = (float)(acc1[3 + k] / (double)M); }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(a1)[i], d = reinterpret_cast<const float4*>(dz1)[i];
    const float a[3] = {v.x, v.y, v.z}, dz[3] = {d.x, d.y, d.z};
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = sc1[k] * (dz[k] - k1[k] - (a[k] - mean1[k]) * invstd1[k] * k2[k]);
    reinterpret_cast<float4*>(dz1)[i] = make_float4(o[0], o[1], o[2], 0.f);
  }
}

// B4: df = dfg * gate (direct path) + conv3x3^T(da1; w0) (through the gate) -> dmain / daux (NCHW fp32, times inv_scale);
//     dW0[k][c][tap] += da1[p][k] f[p + tap][c]  -> acc[162]
template <typename T>
__global__ void __launch_bounds__(256)
gate_bwd4_kernel(const float* __restrict__ main_, const float* __restrict__ aux, const float* __restrict__ a2,
                 const float* __restrict__ sc2, const float* __restrict__ sh2, const T* __restrict__ dfg16,
                 const float* __restrict__ dout4, const float* __restrict__ da1, const __grid_constant__ GateW p,
                 float* __restrict__ dmain, float* __restrict__ daux, double* __restrict__ acc, const float* __restrict__ gscale,
                 int B, int H, int W) {
  const long long HW = (long long)H * W, M = (long long)B * HW;
  const float inv = gscale_inv(gscale);
  __shared__ float sacc[162];
  for (int i = threadIdx.x; i < 162; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const long long nthr = (long long)gridDim.x * blockDim.x, Mr = (M + nthr - 1) / nthr * nthr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < Mr; i += nthr) {
    const bool live = i < M;
    const long long ii = live ? i : 0;
    const int x = (int)(ii % W), y = (int)((ii / W) % H);
    const long long b = ii / HW, o = ii % HW;
    const float* pm = main_ + b * 3 * HW;
    const float* pa = aux + b * 3 * HW;
    float df[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float4 dc = reinterpret_cast<const float4*>(da1)[ii];
    const float dcen[3] = {live ? dc.x : 0.f, live ? dc.y : 0.f, live ? dc.z : 0.f};
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        // forward: a1[p] += f[p + (ky-1, kx-1)] w0[.][.][ky][kx]  =>  df[q] += da1[q - (ky-1, kx-1)] w0[.][.][ky][kx]
        const int yy = y - (ky - 1), xx = x - (kx - 1);
        if (live && yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float4 d = reinterpret_cast<const float4*>(da1)[b * HW + (long long)yy * W + xx];
          const float dn[3] = {d.x, d.y, d.z};
#pragma unroll
          for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int k = 0; k < 3; ++k) df[c] = fmaf(dn[k], p.w0[(k * 6 + c) * 9 + ky * 3 + kx], df[c]);
        }
        // weight gradient of tap (ky, kx): da1[p][k] * f[p + (ky-1, kx-1)][c], reduced over the warp, then the block
        const int fy = y + ky - 1, fx = x + kx - 1;
        float fn[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (live && fy >= 0 && fy < H && fx >= 0 && fx < W) load_f6(pm, pa, HW, (long long)fy * W + fx, fn);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const float s = warp_sum(dcen[k] * fn[c]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&sacc[(k * 6 + c) * 9 + ky * 3 + kx], s);
          }
      }
    if (live) {
      float f[6];
      load_f6(pm, pa, HW, o, f);
      const float4 q0 = reinterpret_cast<const float4*>(a2)[2 * i], q1 = reinterpret_cast<const float4*>(a2)[2 * i + 1];
      const float av[6] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y};
      const F8 dv = load8(dfg16 + i * 16);
      const float4 dq = reinterpret_cast<const float4*>(dout4)[i];
      const float d3[3] = {dq.x, dq.y, dq.z};
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const float gate = 1.f / (1.f + expf(-fmaf(av[c], sc2[c], sh2[c])));
        float dfg = dv.v[c];
#pragma unroll
        for (int k = 0; k < 3; ++k) dfg = fmaf(p.wr[k * 6 + c], d3[k], dfg);
        const float g = (df[c] + dfg * gate) * inv;
        if (c < 3) dmain[b * 3 * HW + c * HW + o] = g;
        else daux[b * 3 * HW + (c - 3) * HW + o] = g;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 162; i += blockDim.x) atomicAdd(acc + i, (double)sacc[i]);
}

// Dropout2d as the reference applies it in training (models.py:290, 294): x[b, :, :, c] *= scale[b][c]
// (scale = keep / (1 - p), drawn or supplied by the caller); also its backward (the same multiply on the gradient).
template <typename T>
__global__ void __launch_bounds__(256)
channel_scale_kernel(T* __restrict__ x, int ld, const float* __restrict__ scale, long long HW, long long M, int C) {
  const int G = C >> 3;
  const long long items = M * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    const long long px = i / G, b = px / HW;
    F8 v = load8(x + px * ld + cg * 8);
    const F8 s = load8(scale + b * C + cg * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) v.v[e] *= s.v[e];
    store8(x + px * ld + cg * 8, v);
  }
}

}  // namespace eunet

using namespace eunet;

static int load_gate_w(GateW& g, const float* params) {
  memcpy(&g, params, sizeof(GateW));
  return 0;
}

extern "C" int eunet_fusion_gate_fwd(const float* out_main, const float* out_aux, const float* params /*host, 219 floats*/,
                                     void* fg16, int dtype, float* res4, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "fusion_gate_fwd: bad shape");
  static_assert(sizeof(FusionGateParams) == 219 * sizeof(float), "FusionGateParams layout is mirrored in Python");
  FusionGateParams p;
  memcpy(&p, params, sizeof(p));
  const long long M = (long long)B * H * W;
  const int grid = clamp_grid((M + 255) / 256, 8);
  if (dtype == EUNET_BF16)
    fusion_gate_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(out_main, out_aux, p, (__nv_bfloat16*)fg16, res4, B, H, W);
  else if (dtype == EUNET_F16)
    fusion_gate_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(out_main, out_aux, p, (__half*)fg16, res4, B, H, W);
  else if (dtype == EUNET_F32)
    fusion_gate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(out_main, out_aux, p, (float*)fg16, res4, B, H, W);
  else {
    set_error("fusion_gate_fwd: unknown dtype %d", dtype);
    return -1;
  }
  return check_launch("fusion_gate_fwd");
}

extern "C" int eunet_fusion_out_fwd(const float* z4, const float* res4, float* out, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "fusion_out_fwd: bad shape");
  const long long M = (long long)B * H * W;
  fusion_out_kernel<<<clamp_grid((M + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(z4, res4, out, B, (long long)H * W);
  return check_launch("fusion_out_fwd");
}

/* ---- training mode of the fusion blocks: see include/eunet.h.  `gate_w`: HOST pointer to 201 floats
 * {w0[3][6][9], w3[6][3], wr[3][6], br[3]}. ---- */
static inline int gate_grid(long long M) { return clamp_grid((M + 255) / 256, 8); }

extern "C" int eunet_fusion_gate_conv_fwd(const float* out_main, const float* out_aux, const float* gate_w, float* a1, double* stats,
                                          int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && out_main && out_aux && gate_w && a1 && stats, "fusion_gate_conv_fwd: bad arguments");
  static_assert(sizeof(GateW) == 201 * sizeof(float), "GateW layout is mirrored in Python");
  GateW g;
  load_gate_w(g, gate_w);
  gate_conv_fwd_kernel<<<gate_grid((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(out_main, out_aux, g, a1, stats, B, H, W);
  return check_launch("fusion_gate_conv_fwd");
}

extern "C" int eunet_fusion_gate_mid_fwd(const float* a1, const float* scale1, const float* shift1, const float* gate_w, float* a2,
                                         double* stats, long long M, void* stream) {
  EUNET_REQUIRE(M > 0 && a1 && scale1 && shift1 && gate_w && a2 && stats, "fusion_gate_mid_fwd: bad arguments");
  GateW g;
  load_gate_w(g, gate_w);
  gate_mid_fwd_kernel<<<gate_grid(M), 256, 0, (cudaStream_t)stream>>>(a1, scale1, shift1, g, a2, stats, M);
  return check_launch("fusion_gate_mid_fwd");
}

extern "C" int eunet_fusion_gate_apply_fwd(const float* out_main, const float* out_aux, const float* a2, const float* scale2,
                                           const float* shift2, const float* gate_w, void* fg16, int dtype, float* res4, int B, int H,
                                           int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && out_main && out_aux && a2 && scale2 && shift2 && gate_w && fg16 && res4,
                "fusion_gate_apply_fwd: bad arguments");
  GateW g;
  load_gate_w(g, gate_w);
  const long long HW = (long long)H * W;
  EUNET_DISPATCH_DTYPE(dtype, gate_apply_fwd_kernel<T><<<gate_grid(B * HW), 256, 0, (cudaStream_t)stream>>>(
                                  out_main, out_aux, a2, scale2, shift2, g, (T*)fg16, res4, B, HW));
  return check_launch("fusion_gate_apply_fwd");
}

extern "C" int eunet_fusion_gate_bwd(const float* out_main, const float* out_aux, const float* a1, const float* a2,
                                     const float* bn1 /* scale, shift, mean, invstd: 4 x [3] */,
                                     const float* bn2 /* scale, shift, mean, invstd: 4 x [6] */, const void* dfg16, int dtype,
                                     const float* dout4, const float* gate_w, float* dz1 /*[M][4] workspace*/,
                                     float* dz2 /*[M][8] workspace*/, double* acc /*[33 + 24 + 162], zeroed by the caller*/,
                                     float* dmain, float* daux, const float* gscale, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "fusion_gate_bwd: bad shape");
  EUNET_REQUIRE(out_main && out_aux && a1 && a2 && bn1 && bn2 && dfg16 && dout4 && gate_w && dz1 && dz2 && acc && dmain && daux,
                "fusion_gate_bwd: null operand");
  GateW g;
  load_gate_w(g, gate_w);
  cudaStream_t st = (cudaStream_t)stream;
  const long long HW = (long long)H * W, M = B * HW;
  const int grid = gate_grid(M);
  const float *sc1 = bn1, *sh1 = bn1 + 3, *mu1 = bn1 + 6, *is1 = bn1 + 9;
  const float *sc2 = bn2, *sh2 = bn2 + 6, *mu2 = bn2 + 12, *is2 = bn2 + 18;
  EUNET_DISPATCH_DTYPE(dtype, gate_bwd1_kernel<T><<<grid, 256, 0, st>>>(out_main, out_aux, a2, sc2, sh2, mu2, is2, (const T*)dfg16, dout4,
                                                                         g, dz2, acc, B, HW));
  if (check_launch("fusion_gate_bwd(1)")) return -2;
  gate_bwd2_kernel<<<grid, 256, 0, st>>>(a1, a2, dz2, sc1, sh1, mu1, is1, sc2, mu2, is2, acc, g, dz1, acc + 33, M);
  if (check_launch("fusion_gate_bwd(2)")) return -2;
  gate_bwd3_kernel<<<grid, 256, 0, st>>>(a1, dz1, sc1, mu1, is1, acc + 33, M);
  if (check_launch("fusion_gate_bwd(3)")) return -2;
  EUNET_DISPATCH_DTYPE(dtype, gate_bwd4_kernel<T><<<grid, 256, 0, st>>>(out_main, out_aux, a2, sc2, sh2, (const T*)dfg16, dout4, dz1, g,
                                                                         dmain, daux, acc + 57, gscale, B, H, W));
  return check_launch("fusion_gate_bwd(4)");
}

extern "C" int eunet_channel_scale(void* x, int ld, const float* scale_bc, int dtype, int B, long long HW, int C, void* stream) {
  EUNET_REQUIRE(B > 0 && HW > 0 && C > 0 && (C & 7) == 0 && (ld & 7) == 0 && ld >= C && x && scale_bc, "channel_scale: bad arguments");
  const long long M = (long long)B * HW;
  EUNET_DISPATCH_DTYPE(dtype, channel_scale_kernel<T><<<clamp_grid((M * (C / 8) + 255) / 256, 16), 256, 0, (cudaStream_t)stream>>>(
                                  (T*)x, ld, scale_bc, HW, M, C));
  return check_launch("channel_scale");
}
