// Fused training loss: logit resize (2x2 mean == bilinear 1/2, train_eval.py:306-310) + softmax +
// 2.5*focal (train_eval.py:37-60) + 2.5*dice (134-157) + 1.0*tversky (159-181), summed per sample and
// divided by the batch size (261-337).  One bandwidth pass forward (10 per-sample sums), a tiny
// finalize, one bandwidth pass backward.  fp32 per-pixel math, fp64 sums.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

// constants of Trainer.__init__ for model_name == 'enhanced_unet' (train_eval.py:74-87, 140, 164, 159)
__constant__ float kCEW[3] = {1.f, 20.f, 10.f};
__constant__ float kAlpha[3] = {1.f, 8.f, 5.f};
__constant__ double kDiceW[3] = {1.0, 15.0, 8.0};
__constant__ double kTvW[3] = {1.0, 12.0, 6.0};
static constexpr double kTvAlpha = 0.7;
static constexpr double kWFocal = 2.5, kWDice = 2.5, kWTv = 1.0;
static constexpr double kSmooth = 1e-6;

// mean-resized logits of target pixel (h,w) of sample b
template <int SCALE>
__device__ __forceinline__ void load_logits(const float* __restrict__ logits, int b, int h, int w, int H, int W, float z[3]) {
  if (SCALE == 2) {
    const long long plane = 4LL * H * W;
    const float* base = logits + (long long)b * 3 * plane + (long long)(2 * h) * (2 * W) + 2 * w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float2 r0 = *reinterpret_cast<const float2*>(base + c * plane);
      const float2 r1 = *reinterpret_cast<const float2*>(base + c * plane + 2 * W);
      z[c] = 0.25f * ((r0.x + r0.y) + (r1.x + r1.y));
    }
  } else {
    const long long plane = (long long)H * W;
    const float* base = logits + (long long)b * 3 * plane + (long long)h * W + w;
#pragma unroll
    for (int c = 0; c < 3; ++c) z[c] = base[c * plane];
  }
}

__device__ __forceinline__ void softmax3(const float z[3], float lp[3], float p[3]) {
  const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
  const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
  const float lse = logf(e0 + e1 + e2);
#pragma unroll
  for (int c = 0; c < 3; ++c) lp[c] = z[c] - m - lse;
  const float inv = 1.f / (e0 + e1 + e2);
  p[0] = e0 * inv; p[1] = e1 * inv; p[2] = e2 * inv;
}

__device__ __forceinline__ int clamp_target(long long t) { return t < 0 ? 0 : (t > 2 ? 2 : (int)t); }

// partial[b][0] = focal sum, [1..3] = I_c, [4..6] = Sp_c, [7..9] = St_c
template <int SCALE>
__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int H, int W,
                double* __restrict__ partial) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  float acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) acc[i] = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)HW; i += gridDim.x * blockDim.x) {   // HW < 2^31 (checked on the host)
    const int h = i / W, w = i % W;
    float z[3], lp[3], p[3];
    load_logits<SCALE>(logits, b, h, w, H, W, z);
    softmax3(z, lp, p);
    const long long traw = target[(long long)b * HW + i];
    const int t = clamp_target(traw);
    const float ce = -lp[t] * kCEW[t];
    const float pt = expf(-ce);
    const float u = 1.f - pt;
    const float u2 = u * u;
    acc[0] += kAlpha[t] * (u2 * u2 * u) * ce;
    // a label outside {0,1,2} (e.g. the ignore value 255) makes the reference's F.cross_entropy raise; a kernel cannot
    // raise, so it poisons the loss with NaN instead of silently training on a clamped label
    if (traw < 0 || traw > 2) acc[0] = __int_as_float(0x7fc00000);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      acc[1 + c] += (t == c) ? p[c] : 0.f;
      acc[4 + c] += p[c];
      acc[7 + c] += (t == c) ? 1.f : 0.f;
    }
  }
  __shared__ float red[8][10];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double s = 0.0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) s += (double)red[wv][threadIdx.x];
    atomicAdd(&partial[b * 10 + threadIdx.x], s);
  }
}

// coef[b][0..2] = dL/dI_c, [3..5] = dL/dSp_c, [6] = focal scale (all already divided by the batch size)
__global__ void loss_finalize_kernel(const double* __restrict__ partial, int B, long long HW, float* __restrict__ loss,
                                     float* __restrict__ per_sample, double* __restrict__ coef) {
  __shared__ double tot[1024];
  double mine = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const double* s = partial + b * 10;
    const double focal = s[0] / (double)HW;
    double dice_l = 0.0, tv_l = 0.0;
    for (int c = 0; c < 3; ++c) {
      const double I = s[1 + c], Sp = s[4 + c], St = s[7 + c];
      const double U = Sp + St + kSmooth;
      const double dice = (2.0 * I + kSmooth) / U;
      dice_l += kDiceW[c] * (1.0 - dice) / 3.0;
      const double fp = Sp - I, fn = St - I;
      const double D = I + kTvAlpha * fp + (1.0 - kTvAlpha) * fn + kSmooth;
      const double tv = (I + kSmooth) / D;
      tv_l += kTvW[c] * (1.0 - tv) / 3.0;
      // gradients wrt the sums (dD/dI = 1 - alpha - (1 - alpha), dD/dSp = alpha)
      const double dDdI = 1.0 - kTvAlpha - (1.0 - kTvAlpha);
      const double dLdI = kWDice * (-(kDiceW[c] / 3.0) * 2.0 / U) +
                          kWTv * (-(kTvW[c] / 3.0) * (1.0 / D - (I + kSmooth) * dDdI / (D * D)));
      const double dLdSp = kWDice * ((kDiceW[c] / 3.0) * (2.0 * I + kSmooth) / (U * U)) +
                           kWTv * ((kTvW[c] / 3.0) * (I + kSmooth) * kTvAlpha / (D * D));
      coef[b * 8 + c] = dLdI / (double)B;
      coef[b * 8 + 3 + c] = dLdSp / (double)B;
    }
    coef[b * 8 + 6] = kWFocal / ((double)HW * (double)B);
    coef[b * 8 + 7] = 0.0;
    const double L = kWFocal * focal + kWDice * dice_l + kWTv * tv_l;
    if (per_sample) per_sample[b] = (float)L;
    mine += L;
  }
  tot[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)blockDim.x; ++i) s += tot[i];
    *loss = (float)(s / (double)B);
  }
}

template <int SCALE>
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ logits, const long long* __restrict__ target, int H, int W,
                const double* __restrict__ coef, const float* __restrict__ grad_out, float* __restrict__ dlogits) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  const float go = grad_out ? *grad_out : 1.f;
  float cI[3], cSp[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    cI[c] = (float)coef[b * 8 + c];
    cSp[c] = (float)coef[b * 8 + 3 + c];
  }
  const float cF = (float)coef[b * 8 + 6];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int)HW; i += gridDim.x * blockDim.x) {   // HW < 2^31 (checked on the host)
    const int h = i / W, w = i % W;
    float z[3], lp[3], p[3];
    load_logits<SCALE>(logits, b, h, w, H, W, z);
    softmax3(z, lp, p);
    const int t = clamp_target(target[(long long)b * HW + i]);
    const float wt = kCEW[t];
    const float ce = -lp[t] * wt;
    const float pt = expf(-ce);
    const float u = 1.f - pt;
    const float u2 = u * u, u4 = u2 * u2;
    // d focal_pixel / d ce = alpha * ((1-pt)^5 + 5 ce (1-pt)^4 pt);  d ce / d z_k = w_t (p_k - [k == t])
    const float dfdce = cF * kAlpha[t] * (u4 * u + 5.f * ce * u4 * pt) * wt;
    float A[3], dot = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      A[c] = cSp[c] + ((t == c) ? cI[c] : 0.f);
      dot += A[c] * p[c];
    }
    float g[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = go * (dfdce * (p[k] - ((t == k) ? 1.f : 0.f)) + p[k] * (A[k] - dot));
    if (SCALE == 2) {
      const long long plane = 4LL * H * W;
      float* base = dlogits + (long long)b * 3 * plane + (long long)(2 * h) * (2 * W) + 2 * w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float q = 0.25f * g[c];
        *reinterpret_cast<float2*>(base + c * plane) = make_float2(q, q);
        *reinterpret_cast<float2*>(base + c * plane + 2 * W) = make_float2(q, q);
      }
    } else {
      float* base = dlogits + (long long)b * 3 * HW + i;
#pragma unroll
      for (int c = 0; c < 3; ++c) base[c * HW] = g[c];
    }
  }
}

}  // namespace eunet

using namespace eunet;

extern "C" int eunet_loss_fwd(const float* logits, const long long* target, int B, int H, int W, int logits_scale,
                              double* partial, float* loss, float* per_sample, double* coef, void* stream) {
  EUNET_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1LL << 30), "loss_fwd: bad shape B=%d H=%d W=%d", B, H, W);
  EUNET_REQUIRE(logits_scale == 1 || logits_scale == 2, "loss_fwd: logits_scale must be 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(partial, 0, sizeof(double) * 10 * B, st);
  EUNET_REQUIRE(e == cudaSuccess, "loss_fwd: memset: %s", cudaGetErrorString(e));
  const long long HW = (long long)H * W;
  long long bx_want = (HW + 255) / 256, bx_cap = ((long long)kNumSMs * 8 + B - 1) / B;
  const int bx = (int)(bx_want < bx_cap ? bx_want : bx_cap);
  dim3 grid(bx, B);
  if (logits_scale == 2) loss_fwd_kernel<2><<<grid, 256, 0, st>>>(logits, target, H, W, partial);
  else loss_fwd_kernel<1><<<grid, 256, 0, st>>>(logits, target, H, W, partial);
  if (check_launch("loss_fwd")) return -2;
  loss_finalize_kernel<<<1, B < 1024 ? ((B + 31) / 32) * 32 : 1024, 0, st>>>(partial, B, HW, loss, per_sample, coef);
  return check_launch("loss_finalize");
}

extern "C" int eunet_loss_bwd(const float* logits, const long long* target, int B, int H, int W, int logits_scale,
                              const double* coef, const float* grad_out, float* dlogits, void* stream) {
  EUNET_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1LL << 30), "loss_bwd: bad shape B=%d H=%d W=%d", B, H, W);
  EUNET_REQUIRE(logits_scale == 1 || logits_scale == 2, "loss_bwd: logits_scale must be 1 or 2");
  const long long HW = (long long)H * W;
  long long bx_want = (HW + 255) / 256, bx_cap = ((long long)kNumSMs * 8 + B - 1) / B;
  const int bx = (int)(bx_want < bx_cap ? bx_want : bx_cap);
  dim3 grid(bx, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (logits_scale == 2) loss_bwd_kernel<2><<<grid, 256, 0, st>>>(logits, target, H, W, coef, grad_out, dlogits);
  else loss_bwd_kernel<1><<<grid, 256, 0, st>>>(logits, target, H, W, coef, grad_out, dlogits);
  return check_launch("loss_bwd");
}
