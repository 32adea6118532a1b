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

// ---- multi-tensor variants: ONE launch covers up to kMaxTensors parameter tensors.  The pointer table travels by
// value as a kernel parameter (no H2D copy); work unit = a 1024-element chunk of one tensor. ----
constexpr int kMaxTensors = 64;
constexpr int kChunk = 1024;
struct TensorTable {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  long long n[kMaxTensors];          // elements per tensor
  int cstart[kMaxTensors + 1];       // prefix sums of chunk counts
  int count;
};

__device__ __forceinline__ int find_tensor(const TensorTable& t, int chunk) {
  int lo = 0, hi = t.count;          // cstart[lo] <= chunk < cstart[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t.cstart[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) sumsq_multi_kernel(const __grid_constant__ TensorTable t, double* __restrict__ out) {
  float s = 0.f;
  const int total = t.cstart[t.count];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    const int k = find_tensor(t, c);
    const long long e0 = (long long)(c - t.cstart[k]) * kChunk + threadIdx.x * 4;
    const float* g = t.g[k];
    const long long n = t.n[k];
    if (e0 + 3 < n && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(g + e0));
      s += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
    } else {
      for (long long i = e0; i < n && i < e0 + 4; ++i) s += g[i] * g[i];
    }
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += (double)red[i];
    atomicAdd(out, tot);
  }
}

__device__ __forceinline__ void adamw_elem(float& p, float g, float& m, float& v, float gs, float lr, float beta1, float beta2,
                                           float eps, float wd, float bc1, float bc2_sqrt) {
  const float gi = g * gs;
  float pi = p * (1.f - lr * wd);
  m = beta1 * m + (1.f - beta1) * gi;
  v = beta2 * v + (1.f - beta2) * gi * gi;
  pi -= (lr / bc1) * (m / (sqrtf(v) / bc2_sqrt + eps));
  p = pi;
}

// CUDA-graph friendly step bookkeeping: the step counter and the learning rate live in device memory, so a captured
// training step advances on every replay.  hyper[0] = lr (written by the host), state[0] = step (incremented here);
// out = {lr, 1 - beta1^step, sqrt(1 - beta2^step), step}.
__global__ void adamw_prepare_kernel(int* __restrict__ step, const float* __restrict__ lr, float beta1, float beta2,
                                     float* __restrict__ out) {
  const int s = *step + 1;
  *step = s;
  out[0] = *lr;
  out[1] = (float)(1.0 - pow((double)beta1, (double)s));
  out[2] = (float)sqrt(1.0 - pow((double)beta2, (double)s));
  out[3] = (float)s;
}

__global__ void __launch_bounds__(256)
adamw_multi_kernel(const __grid_constant__ TensorTable t, const double* __restrict__ gradsq, float max_norm, float lr, float beta1,
                   float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale, const float* __restrict__ hyper) {
  if (hyper != nullptr) {      // device-resident {lr, bc1, bc2_sqrt} (adamw_prepare_kernel)
    lr = hyper[0]; bc1 = hyper[1]; bc2_sqrt = hyper[2];
  }
  float clip = 1.f;
  if (gradsq != nullptr && max_norm > 0.f) {
    const float norm = (float)sqrt(*gradsq) * fabsf(grad_scale);
    clip = fminf(1.f, max_norm / (norm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  const int total = t.cstart[t.count];
  for (int c = blockIdx.x; c < total; c += gridDim.x) {
    const int k = find_tensor(t, c);
    const long long e0 = (long long)(c - t.cstart[k]) * kChunk + threadIdx.x * 4;
    const long long n = t.n[k];
    float* p = t.p[k];
    const float* g = t.g[k];
    float* m = t.m[k];
    float* v = t.v[k];
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                      reinterpret_cast<uintptr_t>(v)) & 15) == 0;
    if (e0 + 3 < n && al) {
      float4 pp = *reinterpret_cast<float4*>(p + e0), mm = *reinterpret_cast<float4*>(m + e0), vv = *reinterpret_cast<float4*>(v + e0);
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g + e0));
      adamw_elem(pp.x, gg.x, mm.x, vv.x, gs, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
      adamw_elem(pp.y, gg.y, mm.y, vv.y, gs, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
      adamw_elem(pp.z, gg.z, mm.z, vv.z, gs, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
      adamw_elem(pp.w, gg.w, mm.w, vv.w, gs, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
      *reinterpret_cast<float4*>(p + e0) = pp;
      *reinterpret_cast<float4*>(m + e0) = mm;
      *reinterpret_cast<float4*>(v + e0) = vv;
    } else {
      for (long long i = e0; i < n && i < e0 + 4; ++i) adamw_elem(p[i], g[i], m[i], v[i], gs, lr, beta1, beta2, eps, wd, bc1, bc2_sqrt);
    }
  }
}

static int fill_table(TensorTable& t, void* const* p, const void* const* g, void* const* m, void* const* v, const long long* n,
                      int first, int count) {
  t.count = count;
  t.cstart[0] = 0;
  for (int i = 0; i < count; ++i) {
    t.p[i] = p ? (float*)p[first + i] : nullptr;
    t.g[i] = (const float*)g[first + i];
    t.m[i] = m ? (float*)m[first + i] : nullptr;
    t.v[i] = v ? (float*)v[first + i] : nullptr;
    t.n[i] = n[first + i];
    EUNET_REQUIRE(n[first + i] > 0 && n[first + i] < (1LL << 40), "multi-tensor: bad element count");
    const long long chunks = (n[first + i] + kChunk - 1) / kChunk;
    EUNET_REQUIRE(t.cstart[i] + chunks < 0x7fffffffLL, "multi-tensor: too many chunks");
    t.cstart[i + 1] = t.cstart[i] + (int)chunks;
  }
  return 0;
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
  // bias corrections in double, as torch.optim.AdamW forms them in Python floats (1 - 0.999f^step carries ~6e-5 in fp32)
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adamw_kernel<<<clamp_grid((n + 255) / 256, 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, gradsq, max_norm, lr, beta1, beta2,
                                                                                 eps, weight_decay, bc1, bc2_sqrt, grad_scale);
  return check_launch("adamw_step");
}

/* multi-tensor forms: host arrays of `count` device pointers / sizes */
extern "C" int eunet_sumsq_multi(const void* const* g, const long long* n, int count, double* out, void* stream) {
  EUNET_REQUIRE(count > 0, "sumsq_multi: count=%d", count);
  for (int first = 0; first < count; first += kMaxTensors) {
    TensorTable t;
    const int c = count - first < kMaxTensors ? count - first : kMaxTensors;
    if (fill_table(t, nullptr, g, nullptr, nullptr, n, first, c)) return -1;
    sumsq_multi_kernel<<<clamp_grid(t.cstart[c], 4), 256, 0, (cudaStream_t)stream>>>(t, out);
    if (check_launch("sumsq_multi")) return -2;
  }
  return 0;
}

extern "C" int eunet_adamw_prepare(int* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_out, void* stream) {
  EUNET_REQUIRE(step_dev && lr_dev && hyper_out, "adamw_prepare: null operand");
  adamw_prepare_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, lr_dev, beta1, beta2, hyper_out);
  return check_launch("adamw_prepare");
}

extern "C" int eunet_adamw_multi(void* const* p, const void* const* g, void* const* m, void* const* v, const long long* n, int count,
                                 const double* gradsq, float max_norm, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int step, float grad_scale, const float* hyper_dev, void* stream) {
  EUNET_REQUIRE(count > 0 && (step >= 1 || hyper_dev != nullptr), "adamw_multi: count=%d step=%d", count, step);
  // bias corrections in double, as torch.optim.AdamW forms them in Python floats (1 - 0.999f^step carries ~6e-5 in fp32)
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  for (int first = 0; first < count; first += kMaxTensors) {
    TensorTable t;
    const int c = count - first < kMaxTensors ? count - first : kMaxTensors;
    if (fill_table(t, p, g, m, v, n, first, c)) return -1;
    adamw_multi_kernel<<<clamp_grid(t.cstart[c], 8), 256, 0, (cudaStream_t)stream>>>(t, gradsq, max_norm, lr, beta1, beta2, eps,
                                                                                    weight_decay, bc1, bc2_sqrt, grad_scale, hyper_dev);
    if (check_launch("adamw_multi")) return -2;
  }
  return 0;
}
