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

}  // namespace eunet

using namespace eunet;

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
