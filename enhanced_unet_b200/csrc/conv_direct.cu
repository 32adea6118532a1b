// fp32 ("fp32 mode", 1e-4 tolerance) 3x3 convolution on CUDA cores: shared-memory tiled implicit GEMM,
// 64 pixels x 64 output channels per CTA, 4x4 register tile per thread, fp32 FMA.  Same interface and
// epilogue options as the tcgen05 bf16 path (conv_tc.cu); also the on-GPU cross-check for it.
#include "common.cuh"
#include "conv.cuh"

namespace eunet {

// y[p, co] = sum_{tap, ci} x[p + tap, ci] * w[co][tap][ci]
__global__ void __launch_bounds__(256)
conv3x3_fwd_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, float* __restrict__ y, int ldy, int B,
                       int H, int W, int Cin, int Cout, double* __restrict__ stats, const float* __restrict__ scale,
                       const float* __restrict__ shift, int relu) {
  __shared__ __align__(16) float Xs[16][64];
  __shared__ __align__(16) float Ws[16][64];
  __shared__ float red[2][16][64];
  const long long M = (long long)B * H * W;
  const long long p0 = (long long)blockIdx.x * 64;
  const int co0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  // loader role: pixel lp (0..63), channel quad lq (0..3)
  const int lp = threadIdx.x >> 2, lq = (threadIdx.x & 3) * 4;
  const long long pl = p0 + lp;
  int lb = 0, lyy = 0, lxx = 0;
  const bool lvalid = pl < M;
  if (lvalid) {
    lxx = (int)(pl % W);
    lyy = (int)((pl / W) % H);
    lb = (int)(pl / ((long long)W * H));
  }
  const int lco = co0 + lp;   // weight row loaded by this thread

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int tap = 0; tap < 9; ++tap) {
    const int yy = lyy + tap / 3 - 1, xx = lxx + tap % 3 - 1;
    const bool inb = lvalid && yy >= 0 && yy < H && xx >= 0 && xx < W;
    const float* xp = x + (((long long)lb * H + yy) * W + xx) * ldx + lq;
    const float* wp = w + ((long long)lco * 9 + tap) * Cin + lq;
    for (int c0 = 0; c0 < Cin; c0 += 16) {
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), wv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb) xv = *reinterpret_cast<const float4*>(xp + c0);
      if (lco < Cout) wv = *reinterpret_cast<const float4*>(wp + c0);
      __syncthreads();
      Xs[lq + 0][lp] = xv.x; Xs[lq + 1][lp] = xv.y; Xs[lq + 2][lp] = xv.z; Xs[lq + 3][lp] = xv.w;
      Ws[lq + 0][lp] = wv.x; Ws[lq + 1][lp] = wv.y; Ws[lq + 2][lp] = wv.z; Ws[lq + 3][lp] = wv.w;
      __syncthreads();
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[ci][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ws[ci][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }

  // batch statistics of the raw fp32 results (rows beyond M contribute zeros)
  if (stats != nullptr) {
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (p0 + ty * 4 + i < M)
#pragma unroll
        for (int j = 0; j < 4; ++j) { s[j] += acc[i][j]; q[j] += acc[i][j] * acc[i][j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) { red[0][ty][tx * 4 + j] = s[j]; red[1][ty][tx * 4 + j] = q[j]; }
    __syncthreads();
    if (threadIdx.x < 128) {
      const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += red[which][r][c];
      if (co0 + c < Cout) atomicAdd(&stats[which * Cout + co0 + c], (double)t);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long p = p0 + ty * 4 + i;
    if (p >= M) continue;
    const int co = co0 + tx * 4;
    if (co >= Cout) continue;   // Cout is a multiple of 4
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[i][j];
      if (scale != nullptr) v = fmaf(v, scale[co + j], shift[co + j]);
      if (relu) v = fmaxf(v, 0.f);
      o[j] = v;
    }
    *reinterpret_cast<float4*>(y + p * ldy + co) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// dw[co][tap][ci] += sum_p dy[p, co] * x[p + tap, ci];  grid = (pixel splits, 9 taps, co tiles * ci tiles)
__global__ void __launch_bounds__(256)
conv3x3_wgrad_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy, int lddy, float* __restrict__ dw,
                         int B, int H, int W, int Cin, int Cout, int ci_tiles, long long px_per_split) {
  __shared__ __align__(16) float Ds[16][64];
  __shared__ __align__(16) float Xs[16][64];
  const long long M = (long long)B * H * W;
  const int tap = blockIdx.y, dyy = tap / 3 - 1, dxx = tap % 3 - 1;
  const int co0 = (blockIdx.z / ci_tiles) * 64, ci0 = (blockIdx.z % ci_tiles) * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // tx -> 4 ci, ty -> 4 co
  const int lp = threadIdx.x >> 4, lq = (threadIdx.x & 15) * 4;   // loader: pixel 0..15, channel quad
  const long long pbeg = (long long)blockIdx.x * px_per_split;
  const long long pend = pbeg + px_per_split < M ? pbeg + px_per_split : M;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long pc = pbeg; pc < pend; pc += 16) {
    const long long p = pc + lp;
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f), xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < pend) {
      if (co0 + lq < Cout) dv = *reinterpret_cast<const float4*>(dy + p * lddy + co0 + lq);
      const int xx = (int)(p % W) + dxx, yy = (int)((p / W) % H) + dyy;
      if (ci0 + lq < Cin && xx >= 0 && xx < W && yy >= 0 && yy < H) {
        const long long b = p / ((long long)W * H);
        xv = *reinterpret_cast<const float4*>(x + ((b * H + yy) * W + xx) * ldx + ci0 + lq);
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Ds[lp][lq]) = dv;
    *reinterpret_cast<float4*>(&Xs[lp][lq]) = xv;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Ds[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + tx * 4 + j;
      if (ci < Cin) atomicAdd(&dw[((long long)co * 9 + tap) * Cin + ci], acc[i][j]);
    }
  }
}

int conv3x3_fwd_f32(const float* x, int ldx, const float* w, float* y, int ldy, int B, int H, int W, int Cin, int Cout,
                    double* stats, const float* scale, const float* shift, int relu, cudaStream_t st) {
  const long long M = (long long)B * H * W;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)((Cout + 63) / 64));
  conv3x3_fwd_f32_kernel<<<grid, 256, 0, st>>>(x, ldx, w, y, ldy, B, H, W, Cin, Cout, stats, scale, shift, relu);
  return check_launch("conv3x3_fwd_f32");
}

int conv3x3_wgrad_f32(const float* x, int ldx, const float* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                      cudaStream_t st) {
  const long long M = (long long)B * H * W;
  const int ci_tiles = (Cin + 63) / 64, co_tiles = (Cout + 63) / 64;
  const int tiles = 9 * ci_tiles * co_tiles;
  long long splits = ((long long)kNumSMs * 4 + tiles - 1) / tiles;
  const long long max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = ((M + splits - 1) / splits + 15) / 16 * 16;
  splits = (M + per - 1) / per;
  dim3 grid((unsigned)splits, 9, (unsigned)(ci_tiles * co_tiles));
  conv3x3_wgrad_f32_kernel<<<grid, 256, 0, st>>>(x, ldx, dy, lddy, dw, B, H, W, Cin, Cout, ci_tiles, per);
  return check_launch("conv3x3_wgrad_f32");
}

}  // namespace eunet
