// Inference post-processing between logits and metrics ("next" row 1 of SURVEY.md §8f):
//   * logits -> probabilities: bilinear 2x-down resize (== 2x2 mean) + softmax over 3 classes
//     (Evaluator._run_model_single, train_eval.py:411-412);
//   * probabilities -> semantic mask: argmax + the order-dependent rule cascade and the two global
//     pixel-ratio filters of Evaluator._convert_probs_to_mask (train_eval.py:455-568).
// Per-pixel float32 arithmetic in the same order as the reference's tensor ops; the ratio filters need
// the per-image live/dead pixel counts first, hence two bandwidth passes with an integer reduction between.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

template <int SCALE>
__global__ void __launch_bounds__(256)
softmax_probs_kernel(const float* __restrict__ logits, float* __restrict__ probs, int H, int W) {
  const int b = blockIdx.y;
  const long long HW = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    float z[3];
    if (SCALE == 2) {
      const int h = (int)(i / W), w = (int)(i % W);
      const long long plane = 4 * HW;
      const float* base = logits + (long long)b * 3 * plane + (long long)(2 * h) * (2 * W) + 2 * w;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float2 r0 = *reinterpret_cast<const float2*>(base + c * plane);
        const float2 r1 = *reinterpret_cast<const float2*>(base + c * plane + 2 * W);
        z[c] = 0.25f * ((r0.x + r0.y) + (r1.x + r1.y));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) z[c] = logits[((long long)b * 3 + c) * HW + i];
    }
    const float m = fmaxf(z[0], fmaxf(z[1], z[2]));
    const float e0 = expf(z[0] - m), e1 = expf(z[1] - m), e2 = expf(z[2] - m);
    const float s = e0 + e1 + e2;
    float* o = probs + (long long)b * 3 * HW + i;
    o[0] = e0 / s;
    o[HW] = e1 / s;
    o[2 * HW] = e2 / s;
  }
}

// the per-pixel cascade, train_eval.py:470-524
__device__ __forceinline__ int cascade(float bg, float live, float dead) {
  int pred = 0;   // torch.argmax: first maximum
  float mx = bg;
  if (live > mx) { mx = live; pred = 1; }
  if (dead > mx) { mx = dead; pred = 2; }
  if (pred == 1 && ((live < 0.42f) || (live <= bg * 1.15f))) pred = 0;
  if (pred == 2 && ((dead < 0.5f) || (dead <= bg * 1.3f) || (bg > 0.3f) || (live > dead * 0.9f))) pred = 0;
  const bool hi_live = (pred == 0) && (live > 0.42f) && (live > bg * 1.15f) && (live > dead * 1.05f);
  if (hi_live) pred = 1;
  const bool hi_dead = (pred == 0) && (dead > 0.5f) && (dead > bg * 1.3f) && (dead > live * 1.1f) && (bg < 0.3f) && !hi_live;
  if (hi_dead) pred = 2;
  if (pred == 1 && (dead > live * 1.15f) && (dead > 0.45f)) pred = 2;
  if (pred == 2 && (live > dead * 1.15f) && (live > 0.42f)) pred = 1;
  if (mx < 0.3f) pred = 0;
  return pred;
}

__global__ void __launch_bounds__(256)
mask_pass1_kernel(const float* __restrict__ probs, unsigned char* __restrict__ mask, int* __restrict__ counts, long long HW) {
  const int b = blockIdx.y;
  const float* p = probs + (long long)b * 3 * HW;
  int nl = 0, nd = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const int c = cascade(p[i], p[HW + i], p[2 * HW + i]);
    mask[(long long)b * HW + i] = (unsigned char)c;
    nl += (c == 1);
    nd += (c == 2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nl += __shfl_xor_sync(0xffffffffu, nl, o);
    nd += __shfl_xor_sync(0xffffffffu, nd, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (nl) atomicAdd(&counts[2 * b], nl);
    if (nd) atomicAdd(&counts[2 * b + 1], nd);
  }
}

// the global ratio filters, train_eval.py:526-563 (ratios are float64 in the reference: python ints / ints)
__global__ void __launch_bounds__(256)
mask_pass2_kernel(const float* __restrict__ probs, unsigned char* __restrict__ mask, const int* __restrict__ counts,
                  long long HW) {
  const int b = blockIdx.y;
  const double live_ratio = (double)counts[2 * b] / (double)HW, dead_ratio = (double)counts[2 * b + 1] / (double)HW;
  const bool f_live = live_ratio > 0.5, f_dead = dead_ratio > 0.15;
  if (!f_live && !f_dead) return;
  const int tier = dead_ratio > 0.4 ? 2 : (dead_ratio > 0.25 ? 1 : 0);
  const float* p = probs + (long long)b * 3 * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    int c = mask[(long long)b * HW + i];
    if (c == 0) continue;
    const float bg = p[i], live = p[HW + i], dead = p[2 * HW + i];
    if (c == 1 && f_live) {
      const bool keep = (live > 0.5f) && (live > bg * 1.3f) && (bg < 0.3f);
      if (!keep) c = 0;
    } else if (c == 2 && f_dead) {
      bool keep;
      if (tier == 2) keep = (dead > 0.65f) && (dead > bg * 1.6f) && (bg < 0.2f) && (live < dead * 0.7f);
      else if (tier == 1) keep = (dead > 0.6f) && (dead > bg * 1.5f) && (bg < 0.25f) && (live < dead * 0.8f);
      else keep = (dead > 0.55f) && (dead > bg * 1.4f) && (bg < 0.25f);
      if (!keep) c = 0;
    }
    mask[(long long)b * HW + i] = (unsigned char)c;
  }
}

static inline dim3 img_grid(long long HW, int B) {
  long long want = (HW + 255) / 256, cap = ((long long)kNumSMs * 8 + B - 1) / B;
  return dim3((unsigned)(want < cap ? want : cap), (unsigned)B);
}

}  // namespace eunet

using namespace eunet;

namespace eunet {
// F.interpolate(mode='bilinear', align_corners=False, antialias=False) on planar fp32 [planes][Hin][Win] -> [planes][Hout][Wout]
// (the multi-scale views of Evaluator._run_tta_inference, train_eval.py:441-451).  ATen's source-index rule:
// src = ratio * (dst + 0.5) - 0.5, clamped at 0; i0 = (int)src, i1 = i0 + (i0 < in - 1), lambda = src - i0.
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float* __restrict__ src, float* __restrict__ dst, int Hin, int Win, int Hout, int Wout, float rh,
                       float rw) {
  const long long plane = blockIdx.y;
  const float* s = src + plane * (long long)Hin * Win;
  float* d = dst + plane * (long long)Hout * Wout;
  const long long n = (long long)Hout * Wout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int oy = (int)(i / Wout), ox = (int)(i % Wout);
    float sy = rh * (oy + 0.5f) - 0.5f, sx = rw * (ox + 0.5f) - 0.5f;
    sy = sy < 0.f ? 0.f : sy;
    sx = sx < 0.f ? 0.f : sx;
    int y0 = (int)sy, x0 = (int)sx;
    y0 = y0 < Hin - 1 ? y0 : Hin - 1;
    x0 = x0 < Win - 1 ? x0 : Win - 1;
    const int y1 = y0 + (y0 < Hin - 1 ? 1 : 0), x1 = x0 + (x0 < Win - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v00 = __ldg(s + (long long)y0 * Win + x0), v01 = __ldg(s + (long long)y0 * Win + x1);
    const float v10 = __ldg(s + (long long)y1 * Win + x0), v11 = __ldg(s + (long long)y1 * Win + x1);
    d[i] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
  }
}

// F.pad(x, (0, w_pad, 0, h_pad), mode='reflect') on planar fp32 [planes][H][W] -> [planes][Hp][Wp] (train_eval.py:249-253,
// 400-406: pad to multiples of 32 at the bottom / right).  Reflection without repeating the edge: index Hp > i >= H reads
// 2 (H - 1) - i.
__global__ void __launch_bounds__(256)
reflect_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int Hp, int Wp) {
  const long long plane = blockIdx.y;
  const float* s = src + plane * (long long)H * W;
  float* d = dst + plane * (long long)Hp * Wp;
  const long long n = (long long)Hp * Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / Wp), x = (int)(i % Wp);
    const int sy = y < H ? y : 2 * (H - 1) - y, sx = x < W ? x : 2 * (W - 1) - x;
    d[i] = __ldg(s + (long long)sy * W + sx);
  }
}

// Mean of the five TTA views of Evaluator._run_tta_inference (train_eval.py:419-453) in ONE pass: the horizontally /
// vertically flipped views are un-flipped by index (no flipped copies), summed in the reference's order and divided by 5.
__global__ void __launch_bounds__(256)
tta_combine_kernel(const float* __restrict__ p0, const float* __restrict__ ph, const float* __restrict__ pv,
                   const float* __restrict__ pa, const float* __restrict__ pb, float* __restrict__ out, int H, int W) {
  const long long off = (long long)blockIdx.y * H * W;
  const long long n = (long long)H * W;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i % W);
    float v = __ldg(p0 + off + i);
    v += __ldg(ph + off + (long long)y * W + (W - 1 - x));
    v += __ldg(pv + off + (long long)(H - 1 - y) * W + x);
    v += __ldg(pa + off + i);
    v += __ldg(pb + off + i);
    out[off + i] = v / 5.0f;
  }
}
}  // namespace eunet

extern "C" int eunet_reflect_pad(const float* src, float* dst, int planes, int H, int W, int Hp, int Wp, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0 && Hp >= H && Wp >= W, "reflect_pad: bad shape");
  EUNET_REQUIRE(Hp - H < H && Wp - W < W, "reflect_pad: padding (%d, %d) must be smaller than the image (%d, %d)", Hp - H, Wp - W, H, W);
  reflect_pad_kernel<<<img_grid((long long)Hp * Wp, planes), 256, 0, (cudaStream_t)stream>>>(src, dst, H, W, Hp, Wp);
  return check_launch("reflect_pad");
}

extern "C" int eunet_tta_combine(const float* p_base, const float* p_hflip, const float* p_vflip, const float* p_s075,
                                 const float* p_s125, float* out, int planes, int H, int W, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(planes > 0 && planes <= 65535 && H > 0 && W > 0, "tta_combine: bad shape");
  EUNET_REQUIRE(p_base && p_hflip && p_vflip && p_s075 && p_s125 && out, "tta_combine: null operand");
  tta_combine_kernel<<<img_grid((long long)H * W, planes), 256, 0, (cudaStream_t)stream>>>(p_base, p_hflip, p_vflip, p_s075, p_s125, out, H, W);
  return check_launch("tta_combine");
}

extern "C" int eunet_resize_bilinear(const float* src, float* dst, int planes, int Hin, int Win, int Hout, int Wout, float ratio_h,
                                     float ratio_w, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(planes > 0 && planes <= 65535 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "resize_bilinear: bad shape");
  EUNET_REQUIRE(ratio_h > 0.f && ratio_w > 0.f, "resize_bilinear: ratios must be positive");
  const dim3 grid = img_grid((long long)Hout * Wout, planes);
  resize_bilinear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, Hin, Win, Hout, Wout, ratio_h, ratio_w);
  return check_launch("resize_bilinear");
}

extern "C" int eunet_softmax_probs(const float* logits, float* probs, int B, int H, int W, int logits_scale, void* stream) {
  EUNET_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "softmax_probs: bad shape");
  EUNET_REQUIRE(logits_scale == 1 || logits_scale == 2, "softmax_probs: logits_scale must be 1 or 2");
  const dim3 grid = img_grid((long long)H * W, B);
  if (logits_scale == 2) softmax_probs_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, probs, H, W);
  else softmax_probs_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(logits, probs, H, W);
  return check_launch("softmax_probs");
}

extern "C" int eunet_probs_to_mask(const float* probs, unsigned char* mask, int* counts, int B, int H, int W, void* stream) {
  EUNET_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0, "probs_to_mask: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(int) * 2 * B, st);
  EUNET_REQUIRE(e == cudaSuccess, "probs_to_mask: memset: %s", cudaGetErrorString(e));
  const long long HW = (long long)H * W;
  EUNET_REQUIRE(HW < (1LL << 31), "probs_to_mask: image too large");
  const dim3 grid = img_grid(HW, B);
  mask_pass1_kernel<<<grid, 256, 0, st>>>(probs, mask, counts, HW);
  if (check_launch("probs_to_mask(pass1)")) return -2;
  mask_pass2_kernel<<<grid, 256, 0, st>>>(probs, mask, counts, HW);
  return check_launch("probs_to_mask(pass2)");
}
