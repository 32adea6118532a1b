// Shared device/host helpers for the Enhanced-UNet B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace eunet {

// ---- error reporting (C-ABI: 0 = ok, negative = error; message via eunet_last_error()) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);   // cudaPeekAtLastError based; sets error text

#define EUNET_REQUIRE(cond, ...)                      \
  do {                                                \
    if (!(cond)) {                                    \
      ::eunet::set_error(__VA_ARGS__);                \
      return -1;                                      \
    }                                                 \
  } while (0)

constexpr int kNumSMs = 148;

enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2 };

// dtype code -> (T = activation / gradient storage type, TY = storage type of RAW pre-BatchNorm conv outputs).
//   EUNET_BF16: bf16 tensors, fp16 raw;  EUNET_F16: fp16 everywhere (gradients carry the power-of-two scale of
//   eunet_grad_scale);  EUNET_F32: fp32 everywhere.
#define EUNET_DISPATCH_DTYPE(dtype, ...)                    \
  do {                                                      \
    if ((dtype) == EUNET_BF16) {                            \
      using T = __nv_bfloat16;                              \
      using TY = __half;                                    \
      __VA_ARGS__;                                          \
    } else if ((dtype) == EUNET_F16) {                      \
      using T = __half;                                     \
      using TY = __half;                                    \
      __VA_ARGS__;                                          \
    } else if ((dtype) == EUNET_F32) {                      \
      using T = float;                                      \
      using TY = float;                                     \
      __VA_ARGS__;                                          \
    } else {                                                \
      set_error("unknown dtype %d", (int)(dtype));          \
      return -1;                                            \
    }                                                       \
  } while (0)

// ---- element access: 8 consecutive channels as fp32, from fp32 or bf16 storage ----
struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 load8(const float* p) {
  F8 r;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
  F8 r;
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void store8(float* p, const F8& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(r.v[4], r.v[5], r.v[6], r.v[7]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
  uint4 u;
  u.x = pack_bf16x2(r.v[0], r.v[1]);
  u.y = pack_bf16x2(r.v[2], r.v[3]);
  u.z = pack_bf16x2(r.v[4], r.v[5]);
  u.w = pack_bf16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// fp16 storage is used for the RAW (pre-BatchNorm) convolution outputs in bf16 mode: 11 mantissa bits keep the
// rounding error of values whose mean is large against their spread 8x below bf16 ("raw fp16").
__device__ __forceinline__ F8 load8(const __half* p) {
  F8 r;
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  lo = fminf(fmaxf(lo, -65504.f), 65504.f);   // saturate instead of overflowing to inf
  hi = fminf(fmaxf(hi, -65504.f), 65504.f);
  __half2 t = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store8(__half* p, const F8& r) {
  uint4 u;
  u.x = pack_f16x2(r.v[0], r.v[1]);
  u.y = pack_f16x2(r.v[2], r.v[3]);
  u.z = pack_f16x2(r.v[4], r.v[5]);
  u.w = pack_f16x2(r.v[6], r.v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// two 16-bit elements in one 32-bit word, F16 ? fp16 : bf16 (the TMA-staged kernels work on raw 32-bit words)
template <bool F16> __device__ __forceinline__ void unpack16x2(uint32_t w, float& lo, float& hi) {
  if (F16) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w));
    lo = f.x; hi = f.y;
  } else {
    lo = __uint_as_float(w << 16); hi = __uint_as_float(w & 0xffff0000u);
  }
}
template <bool F16> __device__ __forceinline__ uint32_t pack16x2(float lo, float hi) {
  return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi);
}
// the power-of-two gradient scale of fp16 mode: gscale = {S, 1/S} on the device, or NULL (= 1)
__device__ __forceinline__ float gscale_fwd(const float* gscale) { return gscale ? __ldg(gscale) : 1.f; }
__device__ __forceinline__ float gscale_inv(const float* gscale) { return gscale ? __ldg(gscale + 1) : 1.f; }

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__half x) { return __half2float(x); }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(fminf(fmaxf(x, -65504.f), 65504.f)); }

// value after a round trip through the storage type (so that statistics / masks computed from
// fp32 accumulators agree with what later kernels read back)
template <typename T> __device__ __forceinline__ float round_to(float x) { return to_f32(from_f32<T>(x)); }

// ---- per-thread cp.async rings: keep several 8-element vectors per thread in flight without holding registers.
// Each thread only ever reads the slots it filled itself, so cp.async.wait_group is the only synchronisation.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int STAGES>
struct Stream8 {
  char* base;
  int stride;
  // smem region of STAGES * nthreads * 8 * sizeof(T) bytes
  __device__ __forceinline__ Stream8(char* smem, int nthreads)
      : base(smem + threadIdx.x * 8 * (int)sizeof(T)), stride(nthreads * 8 * (int)sizeof(T)) {}
  __device__ __forceinline__ void issue(int stage, const T* src) {
    cp_async16(base + stage * stride, src);
    if (sizeof(T) == 4) cp_async16(base + stage * stride + 16, reinterpret_cast<const char*>(src) + 16);
  }
  __device__ __forceinline__ F8 get(int stage) const { return load8(reinterpret_cast<const T*>(base + stage * stride)); }
  static constexpr int bytes(int nthreads) { return STAGES * nthreads * 8 * (int)sizeof(T); }
};

// per-thread ring of float4 values (e.g. a 3-channel fp32 pixel padded to 4)
template <int STAGES>
struct Stream4f {
  char* base;
  int stride;
  __device__ __forceinline__ Stream4f(char* smem, int nthreads) : base(smem + threadIdx.x * 16), stride(nthreads * 16) {}
  __device__ __forceinline__ void issue(int stage, const float* src) { cp_async16(base + stage * stride, src); }
  __device__ __forceinline__ float4 get(int stage) const { return *reinterpret_cast<const float4*>(base + stage * stride); }
  static constexpr int bytes(int nthreads) { return STAGES * nthreads * 16; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int clamp_grid(long long want, int per_sm = 8) {
  long long cap = (long long)kNumSMs * per_sm;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

}  // namespace eunet
