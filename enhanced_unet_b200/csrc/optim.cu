// Optimiser step over flat fp32 buffers (train_eval.py:120 AdamW(lr, weight_decay=1e-4, betas=(.9,.999)),
// 341 clip_grad_norm_(max_norm=1.0), 343 optimizer.step()): one reduction + one fused elementwise pass.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  float s = 0.f;
  const long long n4 = ((reinterpret_cast<uintptr_t>(g) & 15) == 0) ? n / 4 : 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nthr) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + tid; i < n; i += nthr) s += g[i] * g[i];
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += (double)red[i];
    atomicAdd(out, t);
  }
}

// torch.optim.AdamW semantics (decoupled weight decay, bias-corrected moments), with the global-norm
// clip coefficient min(1, max_norm / (||g|| + 1e-6)) of torch.nn.utils.clip_grad_norm_ applied on the fly.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             long long n, const double* __restrict__ gradsq, float max_norm, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float bc1, float bc2_sqrt, float grad_scale) {
  float clip = 1.f;
  if (gradsq != nullptr && max_norm > 0.f) {
    const float norm = (float)sqrt(*gradsq) * fabsf(grad_scale);
    clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * weight_decay);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

}  // namespace eunet

using namespace eunet;

extern "C" int eunet_sumsq(const float* g, long long n, double* out, void* stream) {
  EUNET_REQUIRE(n > 0, "sumsq: n=%lld", n);
  sumsq_kernel<<<clamp_grid((n / 4 + 255) / 256, 4), 256, 0, (cudaStream_t)stream>>>(g, n, out);
  return check_launch("sumsq");
}

extern "C" int eunet_adamw_step(float* p, const float* g, float* m, float* v, long long n, const double* gradsq, float max_norm,
                                float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                void* stream) {
  EUNET_REQUIRE(n > 0 && step >= 1, "adamw_step: n=%lld step=%d", n, step);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<clamp_grid((n + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, gradsq, max_norm, lr, beta1, beta2,
                                                                                 eps, weight_decay, bc1, bc2_sqrt, grad_scale);
  return check_launch("adamw_step");
}
