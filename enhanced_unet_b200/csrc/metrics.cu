// Integer confusion-count reduction behind metrics.py:12-58 and visualization.py:294-311 / 1484-1492.
// Per image a 4x4 int64 matrix CM[g][p] (class 3 = "other value").  Bit-exact: integer adds only.
//
// Bandwidth kernel: 128-bit loads of both masks; per element the (g,p) pair selects one of 16
// nibble counters packed in two 32-bit registers (a shift by >= 32 yields 0 in PTX shl), flushed
// into 32-bit per-thread counters every 15 elements, so the inner loop costs ~7 integer ops/pixel.
#include "common.cuh"
#include "../../include/eunet.h"

namespace eunet {

__device__ __forceinline__ uint32_t shl_clamp(uint32_t v, uint32_t s) {
  uint32_t r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));   // s >= 32 -> 0 (PTX clamps the shift amount)
  return r;
}

struct NibbleAcc {
  uint32_t lo = 0, hi = 0;   // 16 x 4-bit counters
  uint32_t cnt[16];
  int pending = 0;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt[k] = 0;
  }
  __device__ __forceinline__ void flush() {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      cnt[k] += (lo >> (4 * k)) & 15u;
      cnt[8 + k] += (hi >> (4 * k)) & 15u;
    }
    lo = hi = 0;
    pending = 0;
  }
  // g, p already clamped to [0,3]
  __device__ __forceinline__ void add(uint32_t g, uint32_t p) {
    if (pending == 15) flush();   // a nibble holds at most 15
    ++pending;
    const uint32_t s = (g * 4u + p) * 4u;
    lo += shl_clamp(1u, s);
    hi += shl_clamp(1u, s - 32u);
  }
};

template <typename E> struct VecOf;
template <> struct VecOf<long long> { static constexpr int N = 2; };
template <> struct VecOf<int> { static constexpr int N = 4; };
template <> struct VecOf<unsigned char> { static constexpr int N = 16; };

template <typename E>
__device__ __forceinline__ uint32_t clamp_class(E v) {
  // values outside {0,1,2} (negative, 255, ...) -> 3
  const unsigned long long u = (unsigned long long)(long long)v;
  return u > 2ull ? 3u : (uint32_t)u;
}
template <>
__device__ __forceinline__ uint32_t clamp_class<unsigned char>(unsigned char v) {
  return v > 2 ? 3u : (uint32_t)v;
}

template <typename E>
__global__ void __launch_bounds__(256)
confusion_kernel(const E* __restrict__ pred, const E* __restrict__ gt, long long px_per_image,
                 unsigned long long* __restrict__ counts) {
  constexpr int N = VecOf<E>::N;
  const long long img = blockIdx.y;
  const E* p = pred + img * px_per_image;
  const E* g = gt + img * px_per_image;
  NibbleAcc acc;
  acc.init();
  // vector path only when both image bases are 16-byte aligned
  const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
  const long long nvec = aligned ? px_per_image / N : 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthr = (long long)gridDim.x * blockDim.x;
  for (long long v = tid; v < nvec; v += nthr) {
    const uint4 pv = __ldg(reinterpret_cast<const uint4*>(p) + v);
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(g) + v);
    alignas(16) E pe[N];
    alignas(16) E ge[N];
    *reinterpret_cast<uint4*>(pe) = pv;
    *reinterpret_cast<uint4*>(ge) = gv;
#pragma unroll
    for (int i = 0; i < N; ++i) acc.add(clamp_class<E>(ge[i]), clamp_class<E>(pe[i]));
  }
  acc.flush();
  for (long long i = nvec * N + tid; i < px_per_image; i += nthr) {   // scalar tail (and unaligned images)
    acc.add(clamp_class<E>(g[i]), clamp_class<E>(p[i]));
  }
  acc.flush();

  // block reduction: warp shuffle (64-bit) then shared atomics, one global atomic per counter per block
  __shared__ unsigned long long sm[16];
  if (threadIdx.x < 16) sm[threadIdx.x] = 0ull;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    unsigned long long v = acc.cnt[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sm[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 16 && sm[threadIdx.x]) atomicAdd(&counts[img * 16 + threadIdx.x], sm[threadIdx.x]);
}

}  // namespace eunet

extern "C" int eunet_confusion4x4(const void* pred, const void* gt, int elem_bytes, long long n_images,
                                  long long px_per_image, long long* counts, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(n_images >= 0 && px_per_image >= 0, "confusion4x4: negative sizes");
  EUNET_REQUIRE(elem_bytes == 1 || elem_bytes == 4 || elem_bytes == 8, "confusion4x4: elem_bytes %d not in {1,4,8}", elem_bytes);
  if (n_images == 0) return 0;
  EUNET_REQUIRE(n_images <= 65535, "confusion4x4: at most 65535 images per call (got %lld)", n_images);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(long long) * 16 * n_images, st);
  EUNET_REQUIRE(e == cudaSuccess, "confusion4x4: memset: %s", cudaGetErrorString(e));
  if (px_per_image == 0) return 0;
  const int vec = 16 / elem_bytes;
  long long want = (px_per_image / vec + 255) / 256;           // one vector per thread
  want = (want + 3) / 4;                                        // ~4 vectors per thread
  long long per_image_cap = ((long long)kNumSMs * 8 + n_images - 1) / n_images;
  if (per_image_cap < 1) per_image_cap = 1;
  int bx = (int)(want < 1 ? 1 : (want < per_image_cap ? want : per_image_cap));
  dim3 grid(bx, (unsigned)n_images);
  auto* c = reinterpret_cast<unsigned long long*>(counts);
  if (elem_bytes == 8)
    confusion_kernel<long long><<<grid, 256, 0, st>>>((const long long*)pred, (const long long*)gt, px_per_image, c);
  else if (elem_bytes == 4)
    confusion_kernel<int><<<grid, 256, 0, st>>>((const int*)pred, (const int*)gt, px_per_image, c);
  else
    confusion_kernel<unsigned char><<<grid, 256, 0, st>>>((const unsigned char*)pred, (const unsigned char*)gt, px_per_image, c);
  return check_launch("confusion4x4");
}
