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

// ------------------------------------------------------------------------------------------------------------
// Instance matching support (reference metrics.py:61-194, calculate_instance_metrics): the reference evaluates
// calculate_iou(pred_mask, gt_mask) - two full-image logical passes - for every (prediction, ground truth) pair.
// Here every mask is packed ONCE into a bit plane and all |P| x |G| intersection counts come from AND + popcount
// over the bit planes ("binary GEMM"); unions follow from the areas (U = A_p + A_g - I).  Integer work: bit-exact.
// ------------------------------------------------------------------------------------------------------------
namespace eunet {

// masks [n][hw] uint8 (non-zero = member) -> bits [n][words] (bit i of word w = pixel 32 w + i), area [n]
__global__ void __launch_bounds__(256) pack_mask_bits_kernel(const unsigned char* __restrict__ masks, long long hw, long long words,
                                                             unsigned int* __restrict__ bits, unsigned long long* __restrict__ area) {
  const long long img = blockIdx.y;
  const unsigned char* m = masks + img * hw;
  unsigned long long cnt = 0;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long w = warp0; w < words; w += nwarps) {
    const long long px = w * 32 + lane;
    const bool on = px < hw && m[px] != 0;
    const unsigned int word = __ballot_sync(0xffffffffu, on);
    if (lane == 0) {
      bits[img * words + w] = word;
      cnt += __popc(word);
    }
  }
  if (lane == 0 && cnt) atomicAdd(&area[img], cnt);
}

// inter[p][g] += popc(a[p][w] & b[g][w]) over this block's word range.  Block = 8 x 8 pairs x 4 word lanes.
constexpr int kPairChunk = 512;   // words per block iteration (2 KB per mask row)
__global__ void __launch_bounds__(256) pair_intersections_kernel(const unsigned int* __restrict__ a, int na,
                                                                 const unsigned int* __restrict__ b, int nb, long long words,
                                                                 unsigned long long* __restrict__ inter) {
  __shared__ unsigned int sa[8][kPairChunk], sb[8][kPairChunk];
  const int pa0 = blockIdx.x * 8, pb0 = blockIdx.y * 8;
  const int pair = threadIdx.x >> 2, wl = threadIdx.x & 3;
  const int ia = pair >> 3, ib = pair & 7;
  unsigned long long cnt = 0;
  for (long long w0 = (long long)blockIdx.z * kPairChunk; w0 < words; w0 += (long long)gridDim.z * kPairChunk) {
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * kPairChunk; i += 256) {
      const int r = i / kPairChunk, c = i % kPairChunk;
      const long long w = w0 + c;
      sa[r][c] = (pa0 + r < na && w < words) ? a[(long long)(pa0 + r) * words + w] : 0u;
      sb[r][c] = (pb0 + r < nb && w < words) ? b[(long long)(pb0 + r) * words + w] : 0u;
    }
    __syncthreads();
    unsigned int c32 = 0;
#pragma unroll 8
    for (int c = wl; c < kPairChunk; c += 4) c32 += __popc(sa[ia][c] & sb[ib][c]);
    cnt += c32;
  }
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 1);
  cnt += __shfl_xor_sync(0xffffffffu, cnt, 2);
  if (wl == 0 && cnt && pa0 + ia < na && pb0 + ib < nb) atomicAdd(&inter[(long long)(pa0 + ia) * nb + pb0 + ib], cnt);
}

}  // namespace eunet

extern "C" int eunet_pack_mask_bits(const unsigned char* masks, int n, long long hw, unsigned int* bits, long long* area,
                                    void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(n >= 0 && hw >= 0, "pack_mask_bits: negative sizes");
  if (n == 0) return 0;
  EUNET_REQUIRE(n <= 65535, "pack_mask_bits: at most 65535 masks per call (got %d)", n);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(area, 0, sizeof(long long) * n, st);
  EUNET_REQUIRE(e == cudaSuccess, "pack_mask_bits: memset: %s", cudaGetErrorString(e));
  const long long words = (hw + 31) / 32;
  if (words == 0) return 0;
  long long bx = (words + 7) / 8;                      // 8 warps per block, one word per warp and iteration
  const long long cap = ((long long)kNumSMs * 8 + n - 1) / n;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n);
  pack_mask_bits_kernel<<<grid, 256, 0, st>>>(masks, hw, words, bits, reinterpret_cast<unsigned long long*>(area));
  return check_launch("pack_mask_bits");
}

extern "C" int eunet_pair_intersections(const unsigned int* a_bits, int na, const unsigned int* b_bits, int nb, long long words,
                                        long long* inter, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(na >= 0 && nb >= 0 && words >= 0, "pair_intersections: negative sizes");
  if (na == 0 || nb == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(inter, 0, sizeof(long long) * (size_t)na * nb, st);
  EUNET_REQUIRE(e == cudaSuccess, "pair_intersections: memset: %s", cudaGetErrorString(e));
  if (words == 0) return 0;
  const int gx = (na + 7) / 8, gy = (nb + 7) / 8;
  EUNET_REQUIRE(gy <= 65535, "pair_intersections: too many masks (%d)", nb);
  long long gz = (words + kPairChunk - 1) / kPairChunk;
  const long long cap = ((long long)kNumSMs * 4 + (long long)gx * gy - 1) / ((long long)gx * gy);
  if (gz > cap) gz = cap;
  if (gz < 1) gz = 1;
  if (gz > 65535) gz = 65535;
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  pair_intersections_kernel<<<grid, 256, 0, st>>>(a_bits, na, b_bits, nb, words, reinterpret_cast<unsigned long long*>(inter));
  return check_launch("pair_intersections");
}
