// Bandwidth-bound kernels of the hot path: BatchNorm (finalize / apply+ReLU(+pool) / backward),
// 2x2 max-pool, bilinear x2 upsample (forward + adjoint), layout packing.  NHWC, 8 channels
// (128 bit for bf16) per thread access, fp32 math, fp64 cross-CTA accumulation.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

#define DISPATCH_DTYPE EUNET_DISPATCH_DTYPE

static void bn_ring_smem_attr(const void* kernel, int bytes) {
  if (bytes > 48 * 1024) (void)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// Persistent-style grid for a grid-stride REDUCTION kernel: exactly one resident wave (SMs x blocks that fit per SM), so
// no partially filled last wave (1184 blocks at 3 resident per SM = 2.67 waves) and 2.7x fewer end-of-block fp64 atomics.
// (The streaming apply kernels measured 2-3 % slower with it and keep 8 blocks per SM.)
template <typename K>
static int one_wave_grid(K kernel, int threads, int dyn_smem, long long max_blocks) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) {
    (void)cudaGetLastError();
    per_sm = 2;
  }
  long long g = (long long)kNumSMs * per_sm;
  if (g > max_blocks) g = max_blocks;
  return (int)(g < 1 ? 1 : g);
}

static inline int ew_grid(long long items, int threads = 256) {
  long long blocks = (items + threads - 1) / threads;
  long long cap = (long long)kNumSMs * 16;
  return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

// ------------------------------------------------------------------------------------------------
// BatchNorm finalize / eval fold
// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, long long count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, const float* __restrict__ conv_bias,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   long long* __restrict__ nbt, float momentum, float eps, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double n = (double)count;
  const double mean = stats[c] / n;
  double var = stats[C + c] / n - mean * mean;   // biased variance of the bias-free conv output
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma[c];
  const float sc = g * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;          // the conv bias cancels inside train-mode BN
  mean_out[c] = (float)mean;
  invstd_out[c] = invstd;
  if (running_mean != nullptr) {
    const double b = conv_bias ? (double)conv_bias[c] : 0.0;
    const double unbiased = count > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * (mean + b));
    running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
  }
}

__global__ void bn_fold_eval_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ conv_bias, const float* __restrict__ rm,
                                    const float* __restrict__ rv, float eps, float* __restrict__ scale,
                                    float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = gamma[c] / sqrtf(rv[c] + eps);
  const float b = conv_bias ? conv_bias[c] : 0.f;
  scale[c] = sc;
  shift[c] = beta[c] + (b - rm[c]) * sc;
}

// ------------------------------------------------------------------------------------------------
// BN apply + ReLU (+ fused 2x2 max-pool)
// ------------------------------------------------------------------------------------------------
template <typename T, typename TY>
__global__ void __launch_bounds__(256)
bn_apply_relu_kernel(const TY* __restrict__ y, int ldy, T* __restrict__ out, int ldo, long long M, int C,
                     const float* __restrict__ scale, const float* __restrict__ shift) {
  // block = 256 threads = G channel groups x R pixel lanes (G = C/8 divides 256); y streamed through a per-thread ring
  extern __shared__ __align__(16) char dyn_smem[];
  Stream8<TY, 6> sy(dyn_smem, 256);
  const int G = C >> 3, R = 256 / G;
  const int cg = threadIdx.x % G, r = threadIdx.x / G;
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8);
  const long long p0 = (long long)blockIdx.x * R + r, step = (long long)gridDim.x * R;
  const long long n = p0 < M ? (M - p0 + step - 1) / step : 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    if (i < n) sy.issue(i, y + (p0 + i * step) * ldy + cg * 8);
    cp_async_commit();
  }
  for (long long i = 0; i < n; ++i) {
    const long long j = i + 5;
    if (j < n) sy.issue((int)(j % 6), y + (p0 + j * step) * ldy + cg * 8);
    cp_async_commit();
    cp_async_wait<5>();
    F8 v = sy.get((int)(i % 6));
#pragma unroll
    for (int k = 0; k < 8; ++k) v.v[k] = fmaxf(fmaf(v.v[k], sc.v[k], sh.v[k]), 0.f);
    store8(out + (p0 + i * step) * ldo + cg * 8, v);
  }
  cp_async_wait<0>();
}

template <typename T, typename TY>
__global__ void bn_apply_relu_pool_kernel(const TY* __restrict__ y, int ldy, T* __restrict__ out, int ldo,
                                          T* __restrict__ pooled, int ldp, int B, int H, int W, int C,
                                          const float* __restrict__ scale, const float* __restrict__ shift) {
  const int G = C >> 3, Hp = H >> 1, Wp = W >> 1;
  const long long items = (long long)B * Hp * Wp * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long q = i / G;
    const int px = (int)(q % Wp); q /= Wp;
    const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8);
    F8 m;
#pragma unroll
    for (int k = 0; k < 8; ++k) m.v[k] = 0.f;   // post-ReLU values are >= 0
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long p = ((long long)b * H + (2 * py + dy)) * W + (2 * px + dx);
        F8 v = load8(y + p * ldy + cg * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          v.v[k] = fmaxf(fmaf(v.v[k], sc.v[k], sh.v[k]), 0.f);
          // the pooled value must equal the max of the STORED (rounded) activations
          m.v[k] = fmaxf(m.v[k], round_to<T>(v.v[k]));
        }
        store8(out + p * ldo + cg * 8, v);
      }
    const long long pp = ((long long)b * Hp + py) * Wp + px;
    store8(pooled + pp * ldp + cg * 8, m);
  }
}

// ------------------------------------------------------------------------------------------------
// max-pool 2x2
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool2_fwd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int B, int H, int W,
                                    int C) {
  const int G = C >> 3, Hp = H >> 1, Wp = W >> 1;
  const long long items = (long long)B * Hp * Wp * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long q = i / G;
    const int px = (int)(q % Wp); q /= Wp;
    const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    F8 m;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const long long p = ((long long)b * H + (2 * py + dy)) * W + (2 * px + dx);
        const F8 v = load8(x + p * ldx + cg * 8);
#pragma unroll
        for (int k = 0; k < 8; ++k) m.v[k] = (dy == 0 && dx == 0) ? v.v[k] : ((v.v[k] > m.v[k]) ? v.v[k] : m.v[k]);
      }
    store8(out + (((long long)b * Hp + py) * Wp + px) * ldo + cg * 8, m);
  }
}

template <typename T, bool ACC>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ dpool, int ldp, const T* __restrict__ x, int ldx, T* __restrict__ dx,
                                    int lddx, int B, int H, int W, int C) {
  const int G = C >> 3, Hp = H >> 1, Wp = W >> 1;
  const long long items = (long long)B * Hp * Wp * G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long q = i / G;
    const int px = (int)(q % Wp); q /= Wp;
    const int py = (int)(q % Hp);
    const int b = (int)(q / Hp);
    const F8 g = load8(dpool + (((long long)b * Hp + py) * Wp + px) * ldp + cg * 8);
    F8 v[4];
    long long pos[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      pos[w] = ((long long)b * H + (2 * py + (w >> 1))) * W + (2 * px + (w & 1));
      v[w] = load8(x + pos[w] * ldx + cg * 8);
    }
    int arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {   // first maximum in row-major window order (ATen: val > maxval)
      float m = v[0].v[k];
      int a = 0;
#pragma unroll
      for (int w = 1; w < 4; ++w)
        if (v[w].v[k] > m) { m = v[w].v[k]; a = w; }
      arg[k] = a;
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      F8 o;
      if (ACC) o = load8(dx + pos[w] * lddx + cg * 8);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float add = (arg[k] == w) ? g.v[k] : 0.f;
        o.v[k] = ACC ? o.v[k] + add : add;
      }
      store8(dx + pos[w] * lddx + cg * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// bilinear x2 upsample, align_corners=False.  Output rows 2k / 2k+1 of input row k:
//   out[2k]   = .25*in[k-1] + .75*in[k]   (k = 0: in[0]);   out[2k+1] = .75*in[k] + .25*in[k+1]  (k = H-1: in[H-1])
// One thread = 8 channels of one input row PAIR (forward) / one input row (adjoint) over a run of kUpSeg columns: the
// vertically interpolated columns (top / bottom output row) slide along the row in registers, so every step costs 2 loads
// + 4 stores (forward) or 8 loads + 1 store (adjoint) instead of 9 + 4 / 16 + 1.  Item order: channel group fastest, then
// the row, so the threads of a block share their input rows through L1.
// ------------------------------------------------------------------------------------------------
constexpr int kUpSeg = 8;

// AFF: the input is a RAW (pre-BatchNorm) conv output and relu(x * scale + shift) is applied on load - BN apply + ReLU +
// upsample in one pass for the blocks whose activation is consumed by the upsample only (enc4, dec4, dec3).
template <typename T, typename TI, bool AFF>
__global__ void __launch_bounds__(256)
upsample2_fwd_kernel(const TI* __restrict__ x, int ldx, T* __restrict__ out, int ldo, int B, int H, int W, int C,
                     const float* __restrict__ scale, const float* __restrict__ shift) {
  // One item = the PAIR of input rows (kk-1, kk), kk = 0..H, which alone determines the output rows 2kk-1 and 2kk
  //     out[2kk-1] = .75 in[kk-1] + .25 in[kk]   (kk = H: in[H-1]),      out[2kk] = .25 in[kk-1] + .75 in[kk]   (kk = 0: in[0])
  // - two row loads (and, AFF, two BatchNorm + ReLU evaluations) per column instead of the three of a mapping by input row.
  const int G = C >> 3, nseg = (W + kUpSeg - 1) / kUpSeg;
  const long long items = (long long)B * nseg * (H + 1) * G;
  const int Wo = 2 * W, Ho = 2 * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long q = i / G;
    const int kk = (int)(q % (H + 1)); q /= (H + 1);
    const int seg = (int)(q % nseg);
    const int b = (int)(q / nseg);
    const int j0 = seg * kUpSeg, j1 = min(W, j0 + kUpSeg);
    const bool has0 = kk > 0, has1 = kk < H;                 // output rows 2kk-1 / 2kk exist
    // the same products in the same order as the by-row formulas: lower row first
    const float wl0 = has1 ? 0.75f : 1.f, wu0 = has1 ? 0.25f : 0.f;      // row 2kk-1 (kk = H: the last input row itself)
    const float wl1 = has0 ? 0.25f : 0.f, wu1 = has0 ? 0.75f : 1.f;      // row 2kk   (kk = 0: the first input row itself)
    const TI* rl = x + (((long long)b * H + (has0 ? kk - 1 : 0)) * W) * ldx + cg * 8;
    const TI* ru = x + (((long long)b * H + (has1 ? kk : H - 1)) * W) * ldx + cg * 8;
    F8 sc, sh;
    if (AFF) { sc = load8(scale + cg * 8); sh = load8(shift + cg * 8); }
    F8 p0, p1, c0, c1, n0, n1;   // vertically interpolated columns j-1, j, j+1 for output rows 2kk-1 (0) and 2kk (1)
    auto vcol = [&](int j, F8& v0, F8& v1) {
      F8 a = load8(rl + (long long)j * ldx), c = load8(ru + (long long)j * ldx);
      if (AFF) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          a.v[e] = fmaxf(fmaf(a.v[e], sc.v[e], sh.v[e]), 0.f);
          c.v[e] = fmaxf(fmaf(c.v[e], sc.v[e], sh.v[e]), 0.f);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v0.v[e] = wl0 * a.v[e] + wu0 * c.v[e];
        v1.v[e] = wl1 * a.v[e] + wu1 * c.v[e];
      }
    };
    vcol(j0 > 0 ? j0 - 1 : 0, p0, p1);
    vcol(j0, c0, c1);
    // rows 2kk-1 and 2kk are adjacent; for kk = 0 the pointer of the (missing) upper row is never dereferenced
    T* o = out + (((long long)b * Ho + (2 * kk - 1)) * Wo + 2 * j0) * ldo + cg * 8;
#pragma unroll 2
    for (int j = j0; j < j1; ++j) {
      vcol(j < W - 1 ? j + 1 : W - 1, n0, n1);
      const float wl = j > 0 ? 0.25f : 0.f, wc0 = j > 0 ? 0.75f : 1.f;
      const float wr = j < W - 1 ? 0.25f : 0.f, wc1 = j < W - 1 ? 0.75f : 1.f;
      F8 o00, o01, o10, o11;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o00.v[e] = wl * p0.v[e] + wc0 * c0.v[e];
        o01.v[e] = wc1 * c0.v[e] + wr * n0.v[e];
        o10.v[e] = wl * p1.v[e] + wc0 * c1.v[e];
        o11.v[e] = wc1 * c1.v[e] + wr * n1.v[e];
      }
      if (has0) {
        store8(o, o00);
        store8(o + ldo, o01);
      }
      if (has1) {
        store8(o + (long long)Wo * ldo, o10);
        store8(o + (long long)Wo * ldo + ldo, o11);
      }
      o += 2 * (long long)ldo;
      p0 = c0; p1 = c1; c0 = n0; c1 = n1;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
upsample2_bwd_kernel(const T* __restrict__ dout, int ldo, T* __restrict__ dx, int ldx, int B, int H, int W, int C) {
  const int G = C >> 3, nseg = (W + kUpSeg - 1) / kUpSeg;
  const long long items = (long long)B * nseg * H * G;
  const int Wo = 2 * W, Ho = 2 * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int cg = (int)(i % G);
    long long q = i / G;
    const int k = (int)(q % H); q /= H;
    const int seg = (int)(q % nseg);
    const int b = (int)(q / nseg);
    const int j0 = seg * kUpSeg, j1 = min(W, j0 + kUpSeg);
    const float wy[4] = {k > 0 ? 0.25f : 0.f, k > 0 ? 0.75f : 1.f, k < H - 1 ? 0.75f : 1.f, k < H - 1 ? 0.25f : 0.f};
    const T* rows[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int oy = 2 * k - 1 + r;
      oy = oy < 0 ? 0 : (oy > Ho - 1 ? Ho - 1 : oy);     // clamped rows carry weight 0
      rows[r] = dout + (((long long)b * Ho + oy) * Wo) * ldo + cg * 8;
    }
    auto vcol = [&](int ox, F8& v) {
      ox = ox < 0 ? 0 : (ox > Wo - 1 ? Wo - 1 : ox);     // clamped columns carry weight 0
      const F8 g0 = load8(rows[0] + (long long)ox * ldo), g1 = load8(rows[1] + (long long)ox * ldo);
      const F8 g2 = load8(rows[2] + (long long)ox * ldo), g3 = load8(rows[3] + (long long)ox * ldo);
#pragma unroll
      for (int e = 0; e < 8; ++e) v.v[e] = wy[0] * g0.v[e] + wy[1] * g1.v[e] + wy[2] * g2.v[e] + wy[3] * g3.v[e];
    };
    F8 ca, cb, cc, cd;
    vcol(2 * j0 - 1, ca);
    vcol(2 * j0, cb);
    T* o = dx + (((long long)b * H + k) * W + j0) * ldx + cg * 8;
#pragma unroll 2
    for (int j = j0; j < j1; ++j) {
      vcol(2 * j + 1, cc);
      vcol(2 * j + 2, cd);
      const float w0 = j > 0 ? 0.25f : 0.f, w1 = j > 0 ? 0.75f : 1.f;
      const float w2 = j < W - 1 ? 0.75f : 1.f, w3 = j < W - 1 ? 0.25f : 0.f;
      F8 acc;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc.v[e] = w0 * ca.v[e] + w1 * cb.v[e] + w2 * cc.v[e] + w3 * cd.v[e];
      store8(o, acc);
      o += ldx;
      ca = cc; cb = cd;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm + ReLU backward
// ------------------------------------------------------------------------------------------------
// Block = 256 threads = G channel groups x R pixel lanes (G = C/8 divides 256).
constexpr int kBnStages = 4;   // 8-element vectors in flight per thread and stream

template <typename T, typename TY>
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const T* __restrict__ dact, int ldd, const TY* __restrict__ y, int ldy, long long M, int C,
                     const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                     const float* __restrict__ invstd, double* __restrict__ sums) {
  extern __shared__ __align__(16) char dyn_smem[];
  __shared__ float red[2][256 * 8];
  Stream8<T, kBnStages> sd(dyn_smem, 256);
  Stream8<TY, kBnStages> sy(dyn_smem + Stream8<T, kBnStages>::bytes(256), 256);
  const int G = C >> 3, R = 256 / G;
  const int cg = threadIdx.x % G, r = threadIdx.x / G;
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8), mu = load8(mean + cg * 8), is = load8(invstd + cg * 8);
  float sg[8], sgx[8], beta[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sg[e] = sgx[e] = 0.f;
    beta[e] = fmaf(sc.v[e], mu.v[e], sh.v[e]);
  }
  const long long p0 = (long long)blockIdx.x * R + r, step = (long long)gridDim.x * R;
  const long long n = p0 < M ? (M - p0 + step - 1) / step : 0;   // this thread's pixel count
#pragma unroll
  for (int i = 0; i < kBnStages - 1; ++i) {
    if (i < n) {
      const long long p = p0 + i * step;
      sd.issue(i, dact + p * ldd + cg * 8);
      sy.issue(i, y + p * ldy + cg * 8);
    }
    cp_async_commit();
  }
  for (long long i = 0; i < n; ++i) {
    const long long j = i + kBnStages - 1;
    if (j < n) {
      const long long p = p0 + j * step;
      sd.issue((int)(j % kBnStages), dact + p * ldd + cg * 8);
      sy.issue((int)(j % kBnStages), y + p * ldy + cg * 8);
    }
    cp_async_commit();
    cp_async_wait<kBnStages - 1>();
    const F8 d = sd.get((int)(i % kBnStages)), v = sy.get((int)(i % kBnStages));
    // pre = sc * (v - mean) + beta; masked sums with predicated accumulation (5 instructions per element), the
    // invstd factor of xhat is applied once at the end
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xc = v.v[e] - mu.v[e];
      if (fmaf(xc, sc.v[e], beta[e]) > 0.f) {
        sg[e] += d.v[e];
        sgx[e] = fmaf(d.v[e], xc, sgx[e]);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][r * C + cg * 8 + e] = sg[e];
    red[1][r * C + cg * 8 + e] = sgx[e] * is.v[e];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += 256) {
    const int which = t / C, c = t % C;
    float s = 0.f;
    for (int rr = 0; rr < R; ++rr) s += red[which][rr * C + c];
    atomicAdd(&sums[which * C + c], (double)s);
  }
}

template <typename T, typename TY>
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const T* __restrict__ dact, int ldd, const TY* __restrict__ y, int ldy, T* __restrict__ dy, int lddy,
                    long long M, int C, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd, const double* __restrict__ sums,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, const float* __restrict__ gscale) {
  // block = 256 threads = G channel groups x R pixel lanes (G = C/8 divides 256): every thread keeps ONE channel
  // group, so the per-channel constants are loaded once
  const int G = C >> 3, R = 256 / G;
  if (blockIdx.x == 0) {
    const double inv = (double)gscale_inv(gscale);     // fp16 mode: the gradients carry a power-of-two scale
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)(sums[c] * inv);
      if (dgamma) dgamma[c] = (float)(sums[C + c] * inv);
    }
  }
  const int cg = threadIdx.x % G, r = threadIdx.x / G;
  const double invM = 1.0 / (double)M;
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8), mu = load8(mean + cg * 8), is = load8(invstd + cg * 8);
  float k1[8], k2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    k1[e] = (float)(sums[cg * 8 + e] * invM);
    k2[e] = (float)(sums[C + cg * 8 + e] * invM);
  }
  extern __shared__ __align__(16) char dyn_smem[];
  Stream8<T, kBnStages> sd(dyn_smem, 256);
  Stream8<TY, kBnStages> sy(dyn_smem + Stream8<T, kBnStages>::bytes(256), 256);
  const long long p0 = (long long)blockIdx.x * R + r, step = (long long)gridDim.x * R;
  const long long n = p0 < M ? (M - p0 + step - 1) / step : 0;
#pragma unroll
  for (int i = 0; i < kBnStages - 1; ++i) {
    if (i < n) {
      const long long p = p0 + i * step;
      sd.issue(i, dact + p * ldd + cg * 8);
      sy.issue(i, y + p * ldy + cg * 8);
    }
    cp_async_commit();
  }
  for (long long i = 0; i < n; ++i) {
    const long long j = i + kBnStages - 1;
    if (j < n) {
      const long long p = p0 + j * step;
      sd.issue((int)(j % kBnStages), dact + p * ldd + cg * 8);
      sy.issue((int)(j % kBnStages), y + p * ldy + cg * 8);
    }
    cp_async_commit();
    cp_async_wait<kBnStages - 1>();
    const F8 d = sd.get((int)(i % kBnStages)), v = sy.get((int)(i % kBnStages));
    F8 o;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = fmaf(v.v[e], sc.v[e], sh.v[e]) > 0.f ? d.v[e] : 0.f;
      const float xhat = (v.v[e] - mu.v[e]) * is.v[e];
      o.v[e] = sc.v[e] * (g - k1[e] - xhat * k2[e]);
    }
    store8(dy + (p0 + i * step) * lddy + cg * 8, o);
  }
  cp_async_wait<0>();
}

// ---- BN + ReLU backward for an encoder output that ALSO feeds a 2x2 max-pool (enc1.4, enc2.4, enc3.4) ----
// The gradient of such an activation is  dskip (from the decoder's concat slice) + route(dpool)  where the pooled gradient
// goes to the first maximum of its window (ATen tie-break).  Instead of a max-pool backward pass that read-modify-writes
// the skip slice (6.5 B per element) followed by the two BN passes, both BN passes form that sum on the fly: one thread
// owns one 2x2 window x 8 channels, recomputes the four stored activations round_T(relu(y * scale + shift)) - exactly the
// values the forward pool compared - and routes dpool itself.
// One thread = one 2x2 window x eight channels.  The nine 16-byte vectors of a window stay PACKED in registers (36 for 16-bit
// storage) and are expanded a channel pair at a time: expanding all of them first costs 165 registers and one resident block
// per SM (measured 0.19-0.26 of the HBM roofline).
template <typename T> struct Pk8 { uint4 u; };
template <> struct Pk8<float> { float4 a, b; };
template <typename T> __device__ __forceinline__ Pk8<T> ldpk(const T* p) { return Pk8<T>{*reinterpret_cast<const uint4*>(p)}; }
template <> __device__ __forceinline__ Pk8<float> ldpk<float>(const float* p) {
  return Pk8<float>{*reinterpret_cast<const float4*>(p), *reinterpret_cast<const float4*>(p + 4)};
}
__device__ __forceinline__ uint32_t pk_word(const uint4& u, int i) { return i == 0 ? u.x : i == 1 ? u.y : i == 2 ? u.z : u.w; }
__device__ __forceinline__ float2 pk_pair(const Pk8<__half>& r, int kp) {
  const uint32_t w = pk_word(r.u, kp);
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
__device__ __forceinline__ float2 pk_pair(const Pk8<__nv_bfloat16>& r, int kp) {
  const uint32_t w = pk_word(r.u, kp);
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ float2 pk_pair(const Pk8<float>& r, int kp) {
  return kp == 0 ? make_float2(r.a.x, r.a.y) : kp == 1 ? make_float2(r.a.z, r.a.w) : kp == 2 ? make_float2(r.b.x, r.b.y) : make_float2(r.b.z, r.b.w);
}
__device__ __forceinline__ void pk_set(Pk8<__half>& r, int kp, float lo, float hi) {
  const uint32_t w = pack_f16x2(lo, hi);
  if (kp == 0) r.u.x = w; else if (kp == 1) r.u.y = w; else if (kp == 2) r.u.z = w; else r.u.w = w;
}
__device__ __forceinline__ void pk_set(Pk8<__nv_bfloat16>& r, int kp, float lo, float hi) {
  const uint32_t w = pack_bf16x2(lo, hi);
  if (kp == 0) r.u.x = w; else if (kp == 1) r.u.y = w; else if (kp == 2) r.u.z = w; else r.u.w = w;
}
__device__ __forceinline__ void pk_set(Pk8<float>& r, int kp, float lo, float hi) {
  if (kp == 0) { r.a.x = lo; r.a.y = hi; } else if (kp == 1) { r.a.z = lo; r.a.w = hi; }
  else if (kp == 2) { r.b.x = lo; r.b.y = hi; } else { r.b.z = lo; r.b.w = hi; }
}
template <typename T> __device__ __forceinline__ void stpk(T* p, const Pk8<T>& r) { *reinterpret_cast<uint4*>(p) = r.u; }
template <> __device__ __forceinline__ void stpk<float>(float* p, const Pk8<float>& r) {
  *reinterpret_cast<float4*>(p) = r.a;
  *reinterpret_cast<float4*>(p + 4) = r.b;
}

template <typename T, typename TY> struct PoolWindow {
  Pk8<T> dp, d[4];
  Pk8<TY> y[4];
  unsigned pos[4];          // pixel index of each window element (B*H*W < 2^32, checked on the host)
  __device__ __forceinline__ void load(const T* __restrict__ dskip, int ldd, const T* __restrict__ dpool, int ldp,
                                       const TY* __restrict__ yp, int ldy, int H, int W, unsigned item, int lgG) {
    const unsigned Hp = H >> 1, Wp = W >> 1;
    const unsigned c0 = (item & ((1u << lgG) - 1)) * 8;
    unsigned q = item >> lgG;                 // pooled pixel index (b, py, px)
    const unsigned px = q % Wp, r = q / Wp;   // r = b * Hp + py: the input row pair is rows 2r, 2r+1 of the [B*H] stack
    dp = ldpk(dpool + (size_t)q * ldp + c0);
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      pos[w] = (2 * r + (w >> 1)) * (unsigned)W + 2 * px + (w & 1);
      y[w] = ldpk(yp + (size_t)pos[w] * ldy + c0);
      d[w] = ldpk(dskip + (size_t)pos[w] * ldd + c0);
    }
    (void)Hp;
  }
  // Gradient entering BN for channels 2kp, 2kp+1 of the four window elements: ReLU mask of (skip gradient + routed pool gradient).
  // The pool gradient goes to the first maximum (row-major window order, ATen's tie-break) of the fp32 pre-activations: that
  // element is always one of the maxima of the stored (rounded) activations the forward pool compared, and where rounding
  // made several of those equal it is the one an fp32 forward would have picked.
  __device__ __forceinline__ void grads(int kp, float sc0, float sh0, float sc1, float sh1, float (&g)[4][2], float (&yy)[4][2]) const {
    const float2 dpv = pk_pair(dp, kp);
    float p[4][2];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float2 yv = pk_pair(y[w], kp);
      yy[w][0] = yv.x;
      yy[w][1] = yv.y;
      p[w][0] = fmaf(yv.x, sc0, sh0);
      p[w][1] = fmaf(yv.y, sc1, sh1);
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float m = fmaxf(fmaxf(p[0][e], p[1][e]), fmaxf(p[2][e], p[3][e]));
      const float dpe = e ? dpv.y : dpv.x;
      const bool e0 = p[0][e] == m, e1 = !e0 && p[1][e] == m, e2 = !e0 && !e1 && p[2][e] == m, e3 = !e0 && !e1 && !e2;
      const bool hit[4] = {e0, e1, e2, e3};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float2 dv = pk_pair(d[w], kp);
        g[w][e] = p[w][e] > 0.f ? (e ? dv.y : dv.x) + (hit[w] ? dpe : 0.f) : 0.f;
      }
    }
  }
};

__host__ __device__ inline int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}

template <typename T, typename TY>
__global__ void __launch_bounds__(256, 2)
bn_bwd_reduce_pool_kernel(const T* __restrict__ dskip, int ldd, const T* __restrict__ dpool, int ldp, const TY* __restrict__ y,
                          int ldy, int B, int H, int W, int C, int lgG, const float* __restrict__ scale,
                          const float* __restrict__ shift, const float* __restrict__ mean, const float* __restrict__ invstd,
                          double* __restrict__ sums) {
  __shared__ float red[2][256 * 8];
  const int G = 1 << lgG, R = 256 >> lgG;
  const int cg = threadIdx.x & (G - 1), r = threadIdx.x >> lgG;   // 256 % G == 0: a thread keeps its channel group over the grid stride
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8), mu = load8(mean + cg * 8);
  float sg[8], sgx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sg[e] = sgx[e] = 0.f;
  const unsigned items = (unsigned)B * (H >> 1) * (W >> 1) * G;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < items; i += gridDim.x * 256u) {
    PoolWindow<T, TY> win;
    win.load(dskip, ldd, dpool, ldp, y, ldy, H, W, i, lgG);
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      float g[4][2], yy[4][2];
      win.grads(kp, sc.v[2 * kp], sh.v[2 * kp], sc.v[2 * kp + 1], sh.v[2 * kp + 1], g, yy);
#pragma unroll
      for (int w = 0; w < 4; ++w)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sg[2 * kp + e] += g[w][e];
          sgx[2 * kp + e] = fmaf(g[w][e], yy[w][e] - mu.v[2 * kp + e], sgx[2 * kp + e]);
        }
    }
  }
  const F8 is = load8(invstd + cg * 8);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    red[0][r * C + cg * 8 + e] = sg[e];
    red[1][r * C + cg * 8 + e] = sgx[e] * is.v[e];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * C; t += 256) {
    const int which = t / C, c = t % C;
    float s = 0.f;
    for (int rr = 0; rr < R; ++rr) s += red[which][rr * C + c];
    atomicAdd(&sums[which * C + c], (double)s);
  }
}

template <typename T, typename TY>
__global__ void __launch_bounds__(256, 2)
bn_bwd_apply_pool_kernel(const T* __restrict__ dskip, int ldd, const T* __restrict__ dpool, int ldp, const TY* __restrict__ y,
                         int ldy, T* __restrict__ dy, int lddy, int B, int H, int W, int C, int lgG,
                         const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const double* __restrict__ sums, float* __restrict__ dgamma,
                         float* __restrict__ dbeta, const float* __restrict__ gscale) {
  if (blockIdx.x == 0) {
    const double inv = (double)gscale_inv(gscale);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = (float)(sums[c] * inv);
      if (dgamma) dgamma[c] = (float)(sums[C + c] * inv);
    }
  }
  const int cg = threadIdx.x & ((1 << lgG) - 1);
  const double invM = 1.0 / ((double)B * H * W);
  const F8 sc = load8(scale + cg * 8), sh = load8(shift + cg * 8);
  // dy = sc (g - k1 - (y - mean) invstd k2) = sc g + P + Q y
  float P[8], Q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float k1 = (float)(sums[cg * 8 + e] * invM), k2 = (float)(sums[C + cg * 8 + e] * invM);
    const float mu = mean[cg * 8 + e], is = invstd[cg * 8 + e];
    Q[e] = -sc.v[e] * is * k2;
    P[e] = -sc.v[e] * k1 - Q[e] * mu;
  }
  const unsigned items = (unsigned)B * (H >> 1) * (W >> 1) << lgG;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < items; i += gridDim.x * 256u) {
    PoolWindow<T, TY> win;
    win.load(dskip, ldd, dpool, ldp, y, ldy, H, W, i, lgG);
    Pk8<T> o[4];
#pragma unroll
    for (int kp = 0; kp < 4; ++kp) {
      float g[4][2], yy[4][2];
      win.grads(kp, sc.v[2 * kp], sh.v[2 * kp], sc.v[2 * kp + 1], sh.v[2 * kp + 1], g, yy);
#pragma unroll
      for (int w = 0; w < 4; ++w)
        pk_set(o[w], kp, fmaf(sc.v[2 * kp], g[w][0], fmaf(Q[2 * kp], yy[w][0], P[2 * kp])),
               fmaf(sc.v[2 * kp + 1], g[w][1], fmaf(Q[2 * kp + 1], yy[w][1], P[2 * kp + 1])));
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) stpk(dy + (size_t)win.pos[w] * lddy + cg * 8, o[w]);
  }
}

// ------------------------------------------------------------------------------------------------
// packing
// ------------------------------------------------------------------------------------------------
// x [B,C,H,W] fp32 -> NHWC, Cpad channels.  split (C == 3, Cpad >= 9): channels [0,3) = hi = T(x), [3,6) = lo = T(x - hi),
// [6,9) = hi again - with filters packed as {w_hi, w_hi, w_lo} the first convolution evaluates x*w to ~16 mantissa
// bits on the bf16 tensor cores using channels that would otherwise be zero padding.
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int C, int H, int W, int Cpad, int split) {
  const long long HW = (long long)H * W, items = (long long)B * HW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < items; p += (long long)gridDim.x * blockDim.x) {
    const long long hw = p % HW, b = p / HW;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
      F8 v;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int c = c0 + e;
        float val = 0.f;
        if (!split) {
          if (c < C) val = x[(b * C + c) * HW + hw];
        } else if (c < 9) {
          const float xv = x[(b * C + c % 3) * HW + hw];
          const float hi = round_to<T>(xv);
          val = (c >= 3 && c < 6) ? xv - hi : hi;
        }
        v.v[e] = val;
      }
      store8(out + p * Cpad + c0, v);
    }
  }
}

// packed element (r, tap, k) of one filter tensor w [Co][Ci][3][3]:
//   mode 0: out [CoPad][9][CiPad]            = w[r][k][tap]                       (forward operand)
//   mode 1: out [CiPad][9][CoPad]            = w[k][r][8 - tap]                   (dgrad operand: transposed + flipped)
//   mode 2: out [CoPad][9][CiPad], Ci == 3   = {w_hi, w_hi, w_lo} for the {x_hi, x_lo, x_hi} input split
template <typename T>
__device__ __forceinline__ T pack_weight_value(const float* __restrict__ w, int Co, int Ci, int r, int tap, int k, int mode) {
  float v = 0.f;
  if (mode == 0) {
    if (r < Co && k < Ci) v = w[((long long)r * Ci + k) * 9 + tap];
  } else if (mode == 1) {
    if (k < Co && r < Ci) v = w[((long long)k * Ci + r) * 9 + (8 - tap)];
  } else if (r < Co && k < 9) {
    const float wv = w[((long long)r * Ci + k % 3) * 9 + tap];
    const float hi = round_to<T>(wv);
    v = k < 6 ? hi : wv - hi;
  }
  return from_f32<T>(v);
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ out, int Co, int Ci, int CoPad, int CiPad,
                                   int transpose_flip) {
  // output rows R = transpose_flip ? CiPad : CoPad, inner K = transpose_flip ? CoPad : CiPad
  const int rows = transpose_flip == 1 ? CiPad : CoPad, inner = transpose_flip == 1 ? CoPad : CiPad;
  const long long items = (long long)rows * 9 * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % inner);
    const int tap = (int)((i / inner) % 9);
    const int r = (int)(i / ((long long)inner * 9));
    out[i] = pack_weight_value<T>(w, Co, Ci, r, tap, k, transpose_flip);
  }
}

// ONE launch packs every stale filter of the network (30 tensors per training step: forward + dgrad operand of 15
// convolutions).  The table travels by value as a kernel parameter; work unit = 1024 packed elements of one tensor.
constexpr int kPackMax = 32;
constexpr int kPackChunk = 1024;
struct PackTable {
  const float* w[kPackMax];
  void* out[kPackMax];
  int co[kPackMax], ci[kPackMax], copad[kPackMax], cipad[kPackMax], mode[kPackMax];
  int cstart[kPackMax + 1];
  int count;
};

template <typename T>
__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const __grid_constant__ PackTable t) {
  const int total = t.cstart[t.count];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int lo = 0, hi = t.count;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (t.cstart[mid] <= c) lo = mid; else hi = mid;
    }
    const int mode = t.mode[lo], Co = t.co[lo], Ci = t.ci[lo];
    const int rows = mode == 1 ? t.cipad[lo] : t.copad[lo], inner = mode == 1 ? t.copad[lo] : t.cipad[lo];
    const int items = rows * 9 * inner;
    const float* w = t.w[lo];
    T* out = reinterpret_cast<T*>(t.out[lo]);
#pragma unroll
    for (int u = 0; u < kPackChunk / 256; ++u) {
      const int i = (c - t.cstart[lo]) * kPackChunk + u * 256 + threadIdx.x;
      if (i < items) {
        const int k = i % inner, tap = (i / inner) % 9, r = i / (inner * 9);
        out[i] = pack_weight_value<T>(w, Co, Ci, r, tap, k, mode);
      }
    }
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co, int Ci, int CiPad, int hilo,
                                    const float* __restrict__ gscale) {
  const long long items = (long long)Co * Ci * 9;
  const float inv = gscale_inv(gscale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % Ci);
    const int co = (int)(i / (9LL * Ci));
    const float* row = dwp + ((long long)co * 9 + tap) * CiPad;
    dw[i] = (hilo ? row[ci] + row[3 + ci] : row[ci]) * inv;   // hi/lo input split: d/dw sums the x_hi and x_lo channels
  }
}

// the same unpacking for every filter gradient of a backward pass in ONE launch (15 launches of ~6 us of work each otherwise)
struct UnpackTable {
  const float* src[kPackMax];
  float* dst[kPackMax];
  int co[kPackMax], ci[kPackMax], cipad[kPackMax], hilo[kPackMax];
  int cstart[kPackMax + 1];
  int count;
};
__global__ void __launch_bounds__(256) unpack_wgrad_multi_kernel(const __grid_constant__ UnpackTable t, const float* __restrict__ gscale) {
  const float inv = gscale_inv(gscale);
  const int total = t.cstart[t.count];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    int lo = 0, hi = t.count;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (t.cstart[mid] <= c) lo = mid; else hi = mid;
    }
    const int Ci = t.ci[lo], CiPad = t.cipad[lo], hilo = t.hilo[lo];
    const int items = t.co[lo] * Ci * 9;
    const float* dwp = t.src[lo];
    float* dw = t.dst[lo];
    const int i0 = (c - t.cstart[lo]) * kPackChunk;
#pragma unroll 4
    for (int i = i0 + threadIdx.x; i < min(items, i0 + kPackChunk); i += 256) {
      const int tap = i % 9, ci = (i / 9) % Ci, co = i / (9 * Ci);
      const float* row = dwp + ((long long)co * 9 + tap) * CiPad;
      dw[i] = (hilo ? row[ci] + row[3 + ci] : row[ci]) * inv;
    }
  }
}

__global__ void cast_f64_f32_kernel(const double* __restrict__ s, float* __restrict__ d, long long n, const float* __restrict__ gscale) {
  const double inv = (double)gscale_inv(gscale);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = (float)(s[i] * inv);
}

// ---- fp16 mode: power-of-two gradient scale from the largest |dout| (all of backward is linear in dout) ----
__global__ void absmax_kernel(const float* __restrict__ g, long long n, unsigned int* __restrict__ bits) {
  float m = 0.f;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));   // fmaxf drops NaN operands
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(bits, __float_as_uint(m));   // non-negative floats order like their bits
}
__global__ void grad_scale_finalize_kernel(float* __restrict__ gscale, float target) {
  const float m = __uint_as_float(reinterpret_cast<const unsigned int*>(gscale)[2]);
  float S = 1.f;
  if (m > 0.f && m < 3.0e38f) {                      // zero / inf gradients: leave them unscaled
    int e = (int)floorf(log2f(target / m));
    e = e < -60 ? -60 : (e > 60 ? 60 : e);
    S = exp2f((float)e);
  }
  gscale[0] = S;
  gscale[1] = 1.f / S;
}

static int check_vec(const void* p, int ld, int C, const char* what) {
  EUNET_REQUIRE(C > 0 && (C & 7) == 0, "%s: channel count %d must be a multiple of 8", what, C);
  EUNET_REQUIRE((ld & 7) == 0 && ld >= C, "%s: ld %d must be a multiple of 8 and >= C=%d", what, ld, C);
  EUNET_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, "%s: pointer %p not 16-byte aligned", what, p);
  return 0;
}

}  // namespace eunet

namespace eunet {
int g_opt_bn_tma = 1;
int bn_apply_relu_tma(const void* y, int ldy, void* out, int ldo, long long M, int C, const float* scale, const float* shift,
                      bool out_f16, cudaStream_t st);
}  // namespace eunet

using namespace eunet;

extern "C" {

int eunet_bn_finalize(const double* stats, long long count, const float* gamma, const float* beta, const float* conv_bias,
                      float* running_mean, float* running_var, long long* nbt, float momentum, float eps, float* scale,
                      float* shift, float* mean, float* invstd, int C, void* stream) {
  EUNET_REQUIRE(C > 0 && count > 0, "bn_finalize: bad C=%d count=%lld", C, count);
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, conv_bias, running_mean,
                                                                         running_var, nbt, momentum, eps, scale, shift, mean,
                                                                         invstd, C);
  return check_launch("bn_finalize");
}

int eunet_bn_fold_eval(const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, int C, void* stream) {
  EUNET_REQUIRE(C > 0, "bn_fold_eval: bad C=%d", C);
  bn_fold_eval_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, conv_bias, running_mean, running_var,
                                                                          eps, scale, shift, C);
  return check_launch("bn_fold_eval");
}

int eunet_bn_apply_relu(const void* y, int ldy, void* out, int ldo, void* pooled, int ldp, int dtype, int B, int H, int W,
                        int C, const float* scale, const float* shift, void* stream) {
  if (check_vec(y, ldy, C, "bn_apply_relu(y)") || check_vec(out, ldo, C, "bn_apply_relu(out)")) return -1;
  const long long M = (long long)B * H * W;
  EUNET_REQUIRE(M > 0, "bn_apply_relu: empty tensor");
  cudaStream_t st = (cudaStream_t)stream;
  if (pooled) {
    if (check_vec(pooled, ldp, C, "bn_apply_relu(pooled)")) return -1;
    EUNET_REQUIRE((H & 1) == 0 && (W & 1) == 0, "bn_apply_relu: fused pool needs even H,W (got %dx%d)", H, W);
    DISPATCH_DTYPE(dtype, bn_apply_relu_pool_kernel<T, TY><<<ew_grid(M / 4 * (C / 8)), 256, 0, st>>>(
                              (const TY*)y, ldy, (T*)out, ldo, (T*)pooled, ldp, B, H, W, C, scale, shift));
  } else {
    if ((dtype == EUNET_BF16 || dtype == EUNET_F16) && g_opt_bn_tma) {
      const int rc = bn_apply_relu_tma(y, ldy, out, ldo, M, C, scale, shift, dtype == EUNET_F16, st);
      if (rc <= 0) return rc;      // launched or failed; 1 = shape not covered, ring kernel below
    }
    EUNET_REQUIRE(C <= 2048 && 256 % (C / 8) == 0, "bn_apply_relu: C/8=%d must divide 256", C / 8);
    DISPATCH_DTYPE(dtype, bn_ring_smem_attr((const void*)bn_apply_relu_kernel<T, TY>, Stream8<TY, 6>::bytes(256));
                   const int R = 256 / (C / 8);
                   const int grid = clamp_grid((M + R - 1) / R, 8);
                   bn_apply_relu_kernel<T, TY><<<grid, 256, Stream8<TY, 6>::bytes(256), st>>>((const TY*)y, ldy, (T*)out, ldo, M, C,
                                                                                         scale, shift));
  }
  return check_launch("bn_apply_relu");
}

int eunet_bn_bwd_reduce(const void* dact, int ldd, const void* y, int ldy, int dtype, long long M, int C, const float* scale,
                        const float* shift, const float* mean, const float* invstd, double* sums, void* stream) {
  if (check_vec(dact, ldd, C, "bn_bwd_reduce(dact)") || check_vec(y, ldy, C, "bn_bwd_reduce(y)")) return -1;
  EUNET_REQUIRE(C <= 2048 && 256 % (C / 8) == 0, "bn_bwd_reduce: C/8=%d must divide 256", C / 8);
  EUNET_REQUIRE(M > 0, "bn_bwd_reduce: empty tensor");
  const int R = 256 / (C / 8);
  DISPATCH_DTYPE(dtype, const int smem = (int)(8 * 256 * kBnStages * (sizeof(T) + sizeof(TY)));
                 bn_ring_smem_attr((const void*)bn_bwd_reduce_kernel<T, TY>, smem);
                 const int grid = one_wave_grid(bn_bwd_reduce_kernel<T, TY>, 256, smem, (M + R - 1) / R);
                 bn_bwd_reduce_kernel<T, TY><<<grid, 256, smem, (cudaStream_t)stream>>>((const T*)dact, ldd, (const TY*)y, ldy, M, C, scale,
                                                                                     shift, mean, invstd, sums));
  return check_launch("bn_bwd_reduce");
}

int eunet_bn_bwd_apply(const void* dact, int ldd, const void* y, int ldy, void* dy, int lddy, int dtype, long long M, int C,
                       const float* scale, const float* shift, const float* mean, const float* invstd, const double* sums,
                       float* dgamma, float* dbeta, const float* gscale, void* stream) {
  if (check_vec(dact, ldd, C, "bn_bwd_apply(dact)") || check_vec(y, ldy, C, "bn_bwd_apply(y)") ||
      check_vec(dy, lddy, C, "bn_bwd_apply(dy)"))
    return -1;
  EUNET_REQUIRE(M > 0, "bn_bwd_apply: empty tensor");
  EUNET_REQUIRE(C <= 2048 && 256 % (C / 8) == 0, "bn_bwd_apply: C/8=%d must divide 256", C / 8);
  DISPATCH_DTYPE(dtype, bn_ring_smem_attr((const void*)bn_bwd_apply_kernel<T, TY>, (int)(8 * 256 * kBnStages * (sizeof(T) + sizeof(TY))));
                 bn_bwd_apply_kernel<T, TY><<<clamp_grid((M + 256 / (C / 8) - 1) / (256 / (C / 8)), 8), 256, 8 * 256 * kBnStages * (sizeof(T) + sizeof(TY)), (cudaStream_t)stream>>>(
                            (const T*)dact, ldd, (const TY*)y, ldy, (T*)dy, lddy, M, C, scale, shift, mean, invstd, sums, dgamma,
                            dbeta, gscale));
  return check_launch("bn_bwd_apply");
}

int eunet_bn_bwd_reduce_pool(const void* dskip, int ldd, const void* dpool, int ldp, const void* y, int ldy, int dtype, int B, int H,
                             int W, int C, const float* scale, const float* shift, const float* mean, const float* invstd,
                             double* sums, void* stream) {
  if (check_vec(dskip, ldd, C, "bn_bwd_reduce_pool(dskip)") || check_vec(dpool, ldp, C, "bn_bwd_reduce_pool(dpool)") ||
      check_vec(y, ldy, C, "bn_bwd_reduce_pool(y)"))
    return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "bn_bwd_reduce_pool: H,W must be even (got %dx%d)", H, W);
  EUNET_REQUIRE(C <= 2048 && 256 % (C / 8) == 0, "bn_bwd_reduce_pool: C/8=%d must divide 256", C / 8);
  EUNET_REQUIRE((long long)B * H * W < (1ll << 31), "bn_bwd_reduce_pool: %lld pixels exceed the 32-bit index range", (long long)B * H * W);
  const long long items = (long long)B * (H / 2) * (W / 2) * (C / 8);
  const int lgG = ilog2_exact(C / 8);
  DISPATCH_DTYPE(dtype, const int grid = one_wave_grid(bn_bwd_reduce_pool_kernel<T, TY>, 256, 0, (items + 255) / 256);
                 bn_bwd_reduce_pool_kernel<T, TY><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)dskip, ldd, (const T*)dpool, ldp,
                                                                                        (const TY*)y, ldy, B, H, W, C, lgG, scale,
                                                                                        shift, mean, invstd, sums));
  return check_launch("bn_bwd_reduce_pool");
}

int eunet_bn_bwd_apply_pool(const void* dskip, int ldd, const void* dpool, int ldp, const void* y, int ldy, void* dy, int lddy,
                            int dtype, int B, int H, int W, int C, const float* scale, const float* shift, const float* mean,
                            const float* invstd, const double* sums, float* dgamma, float* dbeta, const float* gscale,
                            void* stream) {
  if (check_vec(dskip, ldd, C, "bn_bwd_apply_pool(dskip)") || check_vec(dpool, ldp, C, "bn_bwd_apply_pool(dpool)") ||
      check_vec(y, ldy, C, "bn_bwd_apply_pool(y)") || check_vec(dy, lddy, C, "bn_bwd_apply_pool(dy)"))
    return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "bn_bwd_apply_pool: H,W must be even (got %dx%d)", H, W);
  EUNET_REQUIRE(C <= 2048 && 256 % (C / 8) == 0, "bn_bwd_apply_pool: C/8=%d must divide 256", C / 8);
  EUNET_REQUIRE((long long)B * H * W < (1ll << 31), "bn_bwd_apply_pool: %lld pixels exceed the 32-bit index range", (long long)B * H * W);
  const long long items = (long long)B * (H / 2) * (W / 2) * (C / 8);
  const int lgG = ilog2_exact(C / 8);
  DISPATCH_DTYPE(dtype, bn_bwd_apply_pool_kernel<T, TY><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)dskip, ldd, (const T*)dpool, ldp, (const TY*)y, ldy, (T*)dy, lddy, B, H, W, C, lgG, scale, shift,
                            mean, invstd, sums, dgamma, dbeta, gscale));
  return check_launch("bn_bwd_apply_pool");
}

int eunet_maxpool2_fwd(const void* x, int ldx, void* out, int ldo, int dtype, int B, int H, int W, int C, void* stream) {
  if (check_vec(x, ldx, C, "maxpool2_fwd(x)") || check_vec(out, ldo, C, "maxpool2_fwd(out)")) return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "maxpool2_fwd: H,W must be even (got %dx%d)", H, W);
  const long long items = (long long)B * (H / 2) * (W / 2) * (C / 8);
  DISPATCH_DTYPE(dtype, maxpool2_fwd_kernel<T><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)out, ldo,
                                                                                                  B, H, W, C));
  return check_launch("maxpool2_fwd");
}

int eunet_maxpool2_bwd(const void* dpool, int ldp, const void* x, int ldx, void* dx, int lddx, int accumulate, int dtype,
                       int B, int H, int W, int C, void* stream) {
  if (check_vec(dpool, ldp, C, "maxpool2_bwd(dpool)") || check_vec(x, ldx, C, "maxpool2_bwd(x)") ||
      check_vec(dx, lddx, C, "maxpool2_bwd(dx)"))
    return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && (H & 1) == 0 && (W & 1) == 0, "maxpool2_bwd: H,W must be even (got %dx%d)", H, W);
  const long long items = (long long)B * (H / 2) * (W / 2) * (C / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (accumulate)
    DISPATCH_DTYPE(dtype, maxpool2_bwd_kernel<T, true><<<ew_grid(items), 256, 0, st>>>((const T*)dpool, ldp, (const T*)x, ldx,
                                                                                        (T*)dx, lddx, B, H, W, C));
  else
    DISPATCH_DTYPE(dtype, maxpool2_bwd_kernel<T, false><<<ew_grid(items), 256, 0, st>>>((const T*)dpool, ldp, (const T*)x, ldx,
                                                                                         (T*)dx, lddx, B, H, W, C));
  return check_launch("maxpool2_bwd");
}

int eunet_upsample2_fwd(const void* x, int ldx, void* out, int ldo, int dtype, int B, int H, int W, int C, void* stream) {
  if (check_vec(x, ldx, C, "upsample2_fwd(x)") || check_vec(out, ldo, C, "upsample2_fwd(out)")) return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "upsample2_fwd: empty tensor");
  const long long items = (long long)B * (H + 1) * ((W + kUpSeg - 1) / kUpSeg) * (C / 8);
  DISPATCH_DTYPE(dtype, upsample2_fwd_kernel<T, T, false><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>(
                            (const T*)x, ldx, (T*)out, ldo, B, H, W, C, nullptr, nullptr));
  return check_launch("upsample2_fwd");
}

int eunet_bn_apply_relu_upsample2(const void* y, int ldy, void* out, int ldo, int dtype, int B, int H, int W, int C,
                                  const float* scale, const float* shift, void* stream) {
  if (check_vec(y, ldy, C, "bn_apply_relu_upsample2(y)") || check_vec(out, ldo, C, "bn_apply_relu_upsample2(out)")) return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0 && scale && shift, "bn_apply_relu_upsample2: bad arguments");
  const long long items = (long long)B * (H + 1) * ((W + kUpSeg - 1) / kUpSeg) * (C / 8);
  DISPATCH_DTYPE(dtype, upsample2_fwd_kernel<T, TY, true><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>(
                            (const TY*)y, ldy, (T*)out, ldo, B, H, W, C, scale, shift));
  return check_launch("bn_apply_relu_upsample2");
}

int eunet_upsample2_bwd(const void* dout, int ldo, void* dx, int ldx, int dtype, int B, int H, int W, int C, void* stream) {
  if (check_vec(dout, ldo, C, "upsample2_bwd(dout)") || check_vec(dx, ldx, C, "upsample2_bwd(dx)")) return -1;
  EUNET_REQUIRE(B > 0 && H > 0 && W > 0, "upsample2_bwd: empty tensor");
  const long long items = (long long)B * H * ((W + kUpSeg - 1) / kUpSeg) * (C / 8);
  DISPATCH_DTYPE(dtype, upsample2_bwd_kernel<T><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>((const T*)dout, ldo, (T*)dx,
                                                                                                   ldx, B, H, W, C));
  return check_launch("upsample2_bwd");
}

int eunet_pack_input_nchw(const float* x, void* out, int dtype, int B, int C, int H, int W, int Cpad, int split_hilo,
                          void* stream) {
  EUNET_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && Cpad >= C && (Cpad & 7) == 0, "pack_input: bad shape");
  EUNET_REQUIRE(!split_hilo || (C == 3 && Cpad >= 9), "pack_input: the hi/lo split needs C == 3 and Cpad >= 9");
  DISPATCH_DTYPE(dtype, pack_input_kernel<T><<<ew_grid((long long)B * H * W), 256, 0, (cudaStream_t)stream>>>(x, (T*)out, B, C,
                                                                                                               H, W, Cpad, split_hilo));
  return check_launch("pack_input_nchw");
}

int eunet_pack_weight3x3(const float* w, void* out, int dtype, int Co, int Ci, int CoPad, int CiPad, int transpose_flip,
                         void* stream) {
  EUNET_REQUIRE(Co > 0 && Ci > 0 && CoPad >= Co && CiPad >= Ci, "pack_weight3x3: bad shape");
  EUNET_REQUIRE(transpose_flip >= 0 && transpose_flip <= 2 && (transpose_flip != 2 || (Ci == 3 && CiPad >= 9)),
                "pack_weight3x3: bad mode %d", transpose_flip);
  const long long items = (long long)CoPad * 9 * CiPad;
  DISPATCH_DTYPE(dtype, pack_weight_kernel<T><<<ew_grid(items), 256, 0, (cudaStream_t)stream>>>(w, (T*)out, Co, Ci, CoPad,
                                                                                                 CiPad, transpose_flip));
  return check_launch("pack_weight3x3");
}

int eunet_pack_weight3x3_multi(const void* const* w, void* const* out, const int* co, const int* ci, const int* copad,
                                const int* cipad, const int* mode, int count, int dtype, void* stream) {
  EUNET_REQUIRE(count > 0 && count <= kPackMax, "pack_weight3x3_multi: %d tensors (max %d)", count, kPackMax);
  PackTable t;
  t.count = count;
  t.cstart[0] = 0;
  for (int i = 0; i < count; ++i) {
    EUNET_REQUIRE(co[i] > 0 && ci[i] > 0 && copad[i] >= co[i] && cipad[i] >= ci[i] && mode[i] >= 0 && mode[i] <= 2 &&
                      (mode[i] != 2 || (ci[i] == 3 && cipad[i] >= 9)),
                  "pack_weight3x3_multi: bad entry %d", i);
    const long long items = (long long)copad[i] * 9 * cipad[i];
    EUNET_REQUIRE(items < (1LL << 30), "pack_weight3x3_multi: tensor %d too large", i);
    t.w[i] = (const float*)w[i]; t.out[i] = out[i];
    t.co[i] = co[i]; t.ci[i] = ci[i]; t.copad[i] = copad[i]; t.cipad[i] = cipad[i]; t.mode[i] = mode[i];
    t.cstart[i + 1] = t.cstart[i] + (int)((items + kPackChunk - 1) / kPackChunk);
  }
  const int grid = clamp_grid(t.cstart[count], 16);
  DISPATCH_DTYPE(dtype, pack_weight_multi_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(t));
  return check_launch("pack_weight3x3_multi");
}

int eunet_unpack_wgrad3x3(const float* dw_packed, float* dw, int Co, int Ci, int CiPad, int hilo, const float* gscale,
                          void* stream) {
  EUNET_REQUIRE(Co > 0 && Ci > 0 && CiPad >= Ci && (!hilo || (Ci == 3 && CiPad >= 6)), "unpack_wgrad3x3: bad shape");
  unpack_wgrad_kernel<<<ew_grid((long long)Co * Ci * 9), 256, 0, (cudaStream_t)stream>>>(dw_packed, dw, Co, Ci, CiPad, hilo, gscale);
  return check_launch("unpack_wgrad3x3");
}

int eunet_unpack_wgrad3x3_multi(const void* const* dw_packed, void* const* dw, const int* co, const int* ci, const int* cipad,
                                const int* hilo, int count, const float* gscale, void* stream) {
  EUNET_REQUIRE(count > 0 && count <= kPackMax, "unpack_wgrad3x3_multi: %d tensors (max %d)", count, kPackMax);
  UnpackTable t;
  t.count = count;
  t.cstart[0] = 0;
  for (int i = 0; i < count; ++i) {
    EUNET_REQUIRE(dw_packed[i] && dw[i] && co[i] > 0 && ci[i] > 0 && cipad[i] >= ci[i] && (!hilo[i] || (ci[i] == 3 && cipad[i] >= 6)),
                  "unpack_wgrad3x3_multi: bad entry %d", i);
    const long long items = (long long)co[i] * 9 * ci[i];
    EUNET_REQUIRE(items < (1LL << 30), "unpack_wgrad3x3_multi: tensor %d too large", i);
    t.src[i] = (const float*)dw_packed[i]; t.dst[i] = (float*)dw[i];
    t.co[i] = co[i]; t.ci[i] = ci[i]; t.cipad[i] = cipad[i]; t.hilo[i] = hilo[i];
    t.cstart[i + 1] = t.cstart[i] + (int)((items + kPackChunk - 1) / kPackChunk);
  }
  unpack_wgrad_multi_kernel<<<clamp_grid(t.cstart[count], 16), 256, 0, (cudaStream_t)stream>>>(t, gscale);
  return check_launch("unpack_wgrad3x3_multi");
}

int eunet_cast_f64_f32(const double* src, float* dst, long long n, const float* gscale, void* stream) {
  EUNET_REQUIRE(n > 0, "cast_f64_f32: n=%lld", n);
  cast_f64_f32_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n, gscale);
  return check_launch("cast_f64_f32");
}

int eunet_grad_scale(const float* g, long long n, float target_max, float* gscale, void* stream) {
  EUNET_REQUIRE(g && gscale && n > 0 && target_max > 0.f, "grad_scale: bad arguments");
  EUNET_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, "grad_scale: g must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(gscale + 2, 0, 4, st);
  EUNET_REQUIRE(e == cudaSuccess, "grad_scale: cudaMemsetAsync: %s", cudaGetErrorString(e));
  absmax_kernel<<<clamp_grid((n / 4 + 255) / 256, 8), 256, 0, st>>>(g, n, reinterpret_cast<unsigned int*>(gscale) + 2);
  grad_scale_finalize_kernel<<<1, 1, 0, st>>>(gscale, target_max);
  return check_launch("grad_scale");
}

}  // extern "C"
