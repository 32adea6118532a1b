// v2 forward / dgrad 3x3 convolution: persistent, halo-tile implicit GEMM on tcgen05.
//
// One work item = a block of (16*MT) x 8 output pixels of one image x BN output channels.  For every 64- (or 16-)
// channel chunk ONE 4-D TMA box {KC, 10, 16*MT+2, 1} brings the block's halo tile into shared memory; all nine
// filter taps and all MT 128-row accumulators read it through shifted UMMA descriptors:
//     GEMM row r = (ty, tx) of tile mt, tap (dy, dx)  ->  halo row (16*mt + ty + dy) * 10 + (tx + dx)
// i.e. descriptor start = ((16*mt + dy) * 10 + dx) rows, stride between 8-row groups (one output row of 8 pixels) =
// 10 rows.  (The absolute-address swizzle makes any row shift legal - pinned by tests/test_gpu_probe.py.)
// Activation bytes through L2 drop from 9x (one box per tap, v1) to 1.4x (halo overhead); filter tiles are either
// streamed once per (tap, chunk) for all MT tiles (ring) or kept resident in shared memory for the whole kernel.
// TMEM holds NACC = 2 accumulator sets when they fit (512 columns), so the epilogue of item i overlaps the MMAs of
// item i+1.  BN batch statistics are accumulated in registers across items and flushed once per CTA.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue.
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct ConvHaloParams {
  void* y;           // bf16 or fp16 elements
  int ldy;
  int B, H, W, Cin, Cout;
  int blocks_x, blocks_y, items;   // pixel blocks per image row / column, total pixel blocks
  double* stats;
  const float* scale;
  const float* shift;
  int relu;
  int out_f16;
  uint32_t fmt16;     // tensor-core operand format of x and w: 1 = bf16, 0 = fp16
  int a_ahead;        // option a_ahead (default 1): halo tiles are requested as soon as their slot is free (0: after the previous chunk's last tap)
  float* amax;        // fp16 outputs: atomicMax of |y| over the stored values when it exceeds the fp16 range (may be NULL)
  // EPI = 1 (fused 2Hx2W tail, reference models.py:310-313 + 337): out = d14 + b3 + W3 . relu(acc * scale + shift)
  const float* w3;    // [3][64]
  const float* b3;    // [3]
  const float* d14;   // fp32 [pixels][4]: the residual d1
  float* out;         // fp32 NCHW [B,3,H,W]
};

__device__ __forceinline__ float warp_transpose_sum32h(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = upper ? v[i + s] : v[i];
      const float send = upper ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int KC, int BN, int MT, bool WRES, bool TST = false>
struct HaloCfg {
  static constexpr int ROWB = KC * 2;                                   // bytes per pixel row of a tile
  static constexpr int HALO_ROWS = (16 * MT + 2) * 10;
  static constexpr int A_BYTES = HALO_ROWS * ROWB;                      // exact TMA transaction size
  static constexpr int A_SLOT = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_BYTES = BN * ROWB;
  static constexpr int NACC = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_NEED = NACC * MT * BN;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static constexpr int ASTAGES = (KC == 16 || BN == 128 || BN == 192) ? 3 : 2;      // BN = 128: a third 44 KB halo slot (384 -> 128 at H/2: +12 %)
  static constexpr int BSTAGES = WRES ? 0 : (BN == 192 ? 6 : BN >= 128 ? 4 : 6);   // BN = 192: 6 x 24 KB filter taps in flight
  static constexpr int CH = BN >= 32 ? 32 : 16;
  static constexpr int NCHUNK = BN / CH;
  // TST: the epilogue stages an item's output tile in shared memory (128-byte swizzled rows = one pixel's BN = 64
  // channels) and ONE thread writes it with a TMA tensor store.  Per-thread 16-byte global stores touch 32 different
  // 128-byte lines per warp instruction and made the store-bound layers (3->64 at 2Hx2W) L1/LSU-bound.
  static constexpr int STG_BYTES = TST ? MT * 128 * BN * 2 : 0;
  static_assert(!TST || (BN == 64 && WRES), "TMA-store epilogue: BN = 64 (one 128-byte row per pixel), resident filters");
  static int smem_bytes(int cchunks) {
    return 1024 + ASTAGES * A_SLOT + (WRES ? 9 * cchunks * B_BYTES : BSTAGES * B_BYTES) + 2 * STG_BYTES;
  }
};

template <int KC, int BN, int MT, bool WRES, bool TST, int EPI>
__global__ void __launch_bounds__(320, 1)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmY, const ConvHaloParams p) {
  using C = HaloCfg<KC, BN, MT, WRES, TST>;
  constexpr uint32_t LAYOUT = KC == 64 ? tc::kSwizzle128 : tc::kSwizzle32;
  constexpr int AS = C::ASTAGES, BS = WRES ? 1 : C::BSTAGES, NACC = C::NACC, CH = C::CH;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[AS], a_empty[AS], b_full[BS], b_empty[BS], acc_full[2], acc_empty[2], w_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_scale[BN], s_shift[BN];   // folded-BN affine of this CTA's channel tile (eval epilogue)
  // EPI = 1: enhance.3 weights and the exchange buffer through which the two warps of a TMEM lane quarter (32 channels
  // each) combine their partial class sums: [item parity][quarter][tile][lane][3]
  __shared__ __align__(16) float s_w3[EPI ? 3 * BN : 4];
  __shared__ float s_xch[EPI ? 2 * 4 * MT * 32 * 3 : 1];

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = sbase, b_base = sbase + AS * C::A_SLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.scale != nullptr)
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      s_scale[i] = p.scale[blockIdx.x * BN + i];
      s_shift[i] = p.shift[blockIdx.x * BN + i];
    }
  if (EPI == 1)
    for (int i = threadIdx.x; i < 3 * BN; i += blockDim.x) s_w3[i] = p.w3[i];
  // blockIdx.x = N tile (fastest in launch order: the CTAs that read the same halo tiles run together and share them
  // through L2), blockIdx.y = persistent CTA index over the pixel blocks
  const int n0 = blockIdx.x * BN;
  const int cchunks = p.Cin / KC;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < AS; ++s) { tc::mbar_init(tc::smem_u32(&a_full[s]), 1); tc::mbar_init(tc::smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < BS; ++s) { tc::mbar_init(tc::smem_u32(&b_full[s]), 1); tc::mbar_init(tc::smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tc::smem_u32(&acc_full[s]), 1); tc::mbar_init(tc::smem_u32(&acc_empty[s]), (C::NCHUNK >= 2) ? 8 : 4); }
    tc::mbar_init(tc::smem_u32(&w_full), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW);
    if (TST) tc::tma_prefetch_desc(&tmY);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), C::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      if (WRES) {
        const uint32_t wb = tc::smem_u32(&w_full);
        tc::mbar_expect_tx(wb, 9u * cchunks * C::B_BYTES);
        for (int cc = 0; cc < cchunks; ++cc)
          for (int tap = 0; tap < 9; ++tap)
            tc::tma_load_2d(b_base + (cc * 9 + tap) * C::B_BYTES, &tmW, wb, tap * p.Cin + cc * KC, n0);
      }
      // The halo tiles run AHEAD of the filter taps: this one thread issues both streams, and a tile issued only after the last
      // tap of the previous chunk (which itself waits for a ring slot) would be requested just BSTAGES taps before the MMA
      // warp needs it - less than its load time for the 44 / 85 KB tiles (ncu, 192 -> 64 at H: tensor pipe 46 % busy against the
      // 67 % the shared-memory port allows).  So before every tap the next tile is issued as soon as its slot is free
      // (non-blocking probe), i.e. up to a whole chunk early.
      uint32_t a_it = 0, b_it = 0, a_need = 0;
      int a_item = blockIdx.y, a_cc = 0;
      auto issue_a = [&](bool block) {
        if (a_item >= p.items) return;
        const uint32_t s = a_it % AS;
        const uint32_t eb = tc::smem_u32(&a_empty[s]), par = ((a_it / AS) & 1u) ^ 1u;
        if (block) tc::mbar_wait(eb, par);
        else if (!tc::mbar_test(eb, par)) return;
        const int bx = a_item % p.blocks_x, by = (a_item / p.blocks_x) % p.blocks_y, b = a_item / (p.blocks_x * p.blocks_y);
        const uint32_t fb = tc::smem_u32(&a_full[s]);
        tc::mbar_expect_tx(fb, C::A_BYTES);
        tc::tma_load_4d(a_base + s * C::A_SLOT, &tmX, fb, a_cc * KC, bx * 8 - 1, by * (16 * MT) - 1, b);
        ++a_it;
        if (++a_cc == cchunks) { a_cc = 0; a_item += gridDim.y; }
      };
      for (int item = blockIdx.y; item < p.items; item += gridDim.y) {
        for (int cc = 0; cc < cchunks; ++cc) {
          while (a_it <= a_need) issue_a(true);       // this chunk's tile is in flight before its filter taps
          ++a_need;
          if (!WRES) {
            for (int tap = 0; tap < 9; ++tap) {
              if (p.a_ahead && a_it < a_need + AS - 1) issue_a(false);
              const uint32_t sb = b_it % BS;
              tc::mbar_wait(tc::smem_u32(&b_empty[sb]), ((b_it / BS) & 1u) ^ 1u);
              const uint32_t bb = tc::smem_u32(&b_full[sb]);
              tc::mbar_expect_tx(bb, C::B_BYTES);
              tc::tma_load_2d(b_base + sb * C::B_BYTES, &tmW, bb, tap * p.Cin + cc * KC, n0);
              ++b_it;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(128, BN, 0, 0, p.fmt16);
      if (WRES) {
        tc::mbar_wait(tc::smem_u32(&w_full), 0);
        tc::tc_fence_after();
      }
      uint32_t a_it = 0, b_it = 0, li = 0;
      for (int item = blockIdx.y; item < p.items; item += gridDim.y, ++li) {
        const uint32_t ab = li % NACC;
        tc::mbar_wait(tc::smem_u32(&acc_empty[ab]), ((li / NACC) & 1u) ^ 1u);
        tc::tc_fence_after();
        for (int cc = 0; cc < cchunks; ++cc) {
          const uint32_t s = a_it % AS;
          tc::mbar_wait(tc::smem_u32(&a_full[s]), (a_it / AS) & 1u);
          tc::tc_fence_after();
          const uint32_t a_addr = a_base + s * C::A_SLOT;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t b_addr, sb = 0;
            if (WRES) {
              b_addr = b_base + (cc * 9 + tap) * C::B_BYTES;
            } else {
              sb = b_it % BS;
              tc::mbar_wait(tc::smem_u32(&b_full[sb]), (b_it / BS) & 1u);
              tc::tc_fence_after();
              b_addr = b_base + sb * C::B_BYTES;
            }
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_tap = a_addr + (uint32_t)(((16 * mt + dy) * 10 + dx) * C::ROWB);
#pragma unroll
              for (int j = 0; j < KC / 16; ++j) {
                const uint64_t adesc = tc::make_smem_desc(a_tap + j * 32, 16, 10 * C::ROWB, LAYOUT);
                const uint64_t bdesc = tc::make_smem_desc(b_addr + j * 32, 16, 8 * C::ROWB, LAYOUT);
                tc::umma_bf16(tmem_base + (ab * MT + mt) * BN, adesc, bdesc, idesc, (cc | tap | j) != 0 ? 1u : 0u);
              }
            }
            if (!WRES) {
              tc::umma_commit(tc::smem_u32(&b_empty[sb]));
              ++b_it;
            }
          }
          tc::umma_commit(tc::smem_u32(&a_empty[s]));
          ++a_it;
        }
        tc::umma_commit(tc::smem_u32(&acc_full[ab]));
      }
    }
  } else {
    // ===================== epilogue: 8 warps =====================
    // warp w owns TMEM lanes [32 (w%4), +32); the two warps of a lane quarter split the column chunks (even / odd).
    // BN statistics: per-thread fp32 partial sums (packed f32x2 adds / fmas) over its rows, reduced over the warp's 32
    // rows by ONE 31-shuffle transpose-reduce per flush into fp64 per-lane column totals, which go out with fp64
    // atomics when the CTA is done.  With one column chunk per warp (BN <= 64) the partials stay in registers across
    // kFlush items (<= 32 addends per fp32 partial); otherwise they are flushed per chunk and item.
    // With two epilogue warps per scheduler the epilogue is issue-latency bound, so instruction count is what matters.
    const int e = warp - 2, q = warp & 3, half = e >> 2;
    // TMA-store staging (2 buffers, 1024-byte aligned: the filters / B ring in front are multiples of 1 KB)
    const uint32_t stg_base = b_base + (WRES ? 9u * cchunks * C::B_BYTES : (uint32_t)(C::BSTAGES * C::B_BYTES));
    constexpr int NCW = (C::NCHUNK >= 2) ? C::NCHUNK / 2 : 1;
    constexpr int kFlush = 8;
    const bool active = (C::NCHUNK >= 2) || half == 0;
    if (active) {
      const int r = q * 32 + lane;                 // GEMM row inside a 128-row tile: (ty, tx) = (r / 8, r % 8)
      const int ty = r >> 3, tx = r & 7;
      const bool want_stats = p.stats != nullptr, affine = EPI == 0 && p.scale != nullptr, do_store = p.y != nullptr;
      const bool want_amax = p.amax != nullptr && p.out_f16 != 0;
      // every 32-channel chunk of every pixel row starts on a 32-byte boundary (channel-slice views of a wider buffer qualify
      // when their offset and leading dimension are multiples of 16 elements)
      const bool wide = !TST && (reinterpret_cast<uintptr_t>(p.y) & 31) == 0 && (p.ldy & 15) == 0;
      float amax = 0.f;
      double tot1[NCW], tot2[NCW];
#pragma unroll
      for (int cw = 0; cw < NCW; ++cw) tot1[cw] = tot2[cw] = 0.0;
      uint64_t run1[16], run2[16];             // f32x2 pairs: columns (2i, 2i+1) of the current chunk
#pragma unroll
      for (int i = 0; i < 16; ++i) run1[i] = run2[i] = 0ull;
      auto flush = [&](int cw) {
        float a[32], c[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          tc::unpack_f32x2(run1[i], a[2 * i], a[2 * i + 1]);
          tc::unpack_f32x2(run2[i], c[2 * i], c[2 * i + 1]);
          run1[i] = run2[i] = 0ull;
        }
        tot1[cw] += (double)warp_transpose_sum32h(a, lane);
        tot2[cw] += (double)warp_transpose_sum32h(c, lane);
      };
      uint32_t li = 0;
      int pending = 0;
      for (int item = blockIdx.y; item < p.items; item += gridDim.y, ++li) {
        const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
        const uint32_t ab = li % NACC;
        tc::mbar_wait(tc::smem_u32(&acc_full[ab]), (li / NACC) & 1u);
        tc::tc_fence_after();
        const int gx = bx * 8 + tx;
        float part[EPI ? MT : 1][3];
        // EPI = 1: the residual d1 of this thread's pixels is requested NOW and consumed after the accumulator chunks have been
        // reduced (a load issued where it is used put ~1 us of HBM latency on every item of the epilogue chain)
        float4 dres[EPI ? MT : 1];
        if (EPI == 1 && half == 0) {
          const long long HW = (long long)p.H * p.W;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const int gy = by * (16 * MT) + 16 * mt + ty;
            dres[EPI ? mt : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx < p.W && gy < p.H)
              dres[EPI ? mt : 0] = __ldg(reinterpret_cast<const float4*>(p.d14) + (long long)b * HW + (long long)gy * p.W + gx);
          }
        }
        // The T = NCW x MT accumulator chunks of an item are drained through TWO register buffers: the TMEM load of chunk t+1 is
        // in flight while chunk t is converted and stored (one buffer: every chunk paid the load latency in full, and the
        // K = 576 layers - 4608 MMA clocks per item - were bound by this loop: ncu tc pipe 43-65 % busy).
        auto issue = [&](int cw, int mt, uint32_t (&raw)[32]) {
          const int c0 = ((C::NCHUNK >= 2) ? 2 * cw + half : 0) * CH;
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ab * MT + mt) * BN + c0);
          if (CH == 32) tc::tmem_ld32(taddr, raw);
          else {
            tc::tmem_ld16(taddr, raw);
#pragma unroll
            for (int i = 16; i < 32; ++i) raw[i] = 0u;
          }
        };
        auto process = [&](int cw, int mt, uint32_t (&raw)[32]) {
          const int c0 = ((C::NCHUNK >= 2) ? 2 * cw + half : 0) * CH;
          const int gy = by * (16 * MT) + 16 * mt + ty;
          const bool valid = gx < p.W && gy < p.H;
          if (want_stats && valid) {
#pragma unroll
            for (int i = 0; i < CH / 2; ++i) {
              const uint64_t v = tc::pack_f32x2(raw[2 * i], raw[2 * i + 1]);
              run1[i] = tc::add_f32x2(run1[i], v);
              run2[i] = tc::fma_f32x2(v, v, run2[i]);
            }
          }
          if (EPI == 1) {
            // this warp's 32 channels of relu(bn(acc)) . W3 for the pixel of this thread
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, u0 = 0.f, u1 = 0.f, u2 = 0.f;     // two chains per class (ILP)
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              const float4 sc = *reinterpret_cast<const float4*>(&s_scale[c0 + i]), sh = *reinterpret_cast<const float4*>(&s_shift[c0 + i]);
              const float4 w0 = *reinterpret_cast<const float4*>(&s_w3[c0 + i]), w1 = *reinterpret_cast<const float4*>(&s_w3[BN + c0 + i]);
              const float4 w2 = *reinterpret_cast<const float4*>(&s_w3[2 * BN + c0 + i]);
              const float a0 = fmaxf(fmaf(__uint_as_float(raw[i]), sc.x, sh.x), 0.f), a1 = fmaxf(fmaf(__uint_as_float(raw[i + 1]), sc.y, sh.y), 0.f);
              const float a2 = fmaxf(fmaf(__uint_as_float(raw[i + 2]), sc.z, sh.z), 0.f), a3 = fmaxf(fmaf(__uint_as_float(raw[i + 3]), sc.w, sh.w), 0.f);
              t0 = fmaf(a2, w0.z, fmaf(a0, w0.x, t0)); u0 = fmaf(a3, w0.w, fmaf(a1, w0.y, u0));
              t1 = fmaf(a2, w1.z, fmaf(a0, w1.x, t1)); u1 = fmaf(a3, w1.w, fmaf(a1, w1.y, u1));
              t2 = fmaf(a2, w2.z, fmaf(a0, w2.x, t2)); u2 = fmaf(a3, w2.w, fmaf(a1, w2.y, u2));
            }
            t0 += u0; t1 += u1; t2 += u2;
            part[EPI ? mt : 0][0] = t0; part[EPI ? mt : 0][1] = t1; part[EPI ? mt : 0][2] = t2;
          }
          if (do_store && (TST || valid)) {
            uint32_t dst_s = 0;
            uint16_t* dst_g = nullptr;
            if (TST) {   // every row is staged (rows outside the image are clipped by the tensor store)
              dst_s = stg_base + (li & 1u) * C::STG_BYTES + (uint32_t)((mt * 128 + r) * 128);
            } else {
              const long long pix = ((long long)b * p.H + gy) * p.W + gx;
              dst_g = reinterpret_cast<uint16_t*>(p.y) + pix * p.ldy + n0 + c0;
            }
            uint4 held = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int g8 = 0; g8 < CH / 8; ++g8) {
              float o[8];
#pragma unroll
              for (int ee = 0; ee < 8; ++ee) o[ee] = __uint_as_float(raw[g8 * 8 + ee]);
              if (affine) {
#pragma unroll
                for (int h4 = 0; h4 < 2; ++h4) {
                  const float4 sc = *reinterpret_cast<const float4*>(&s_scale[c0 + g8 * 8 + h4 * 4]);
                  const float4 sh = *reinterpret_cast<const float4*>(&s_shift[c0 + g8 * 8 + h4 * 4]);
                  o[h4 * 4 + 0] = fmaf(o[h4 * 4 + 0], sc.x, sh.x); o[h4 * 4 + 1] = fmaf(o[h4 * 4 + 1], sc.y, sh.y);
                  o[h4 * 4 + 2] = fmaf(o[h4 * 4 + 2], sc.z, sh.z); o[h4 * 4 + 3] = fmaf(o[h4 * 4 + 3], sc.w, sh.w);
                }
              }
              if (p.relu) {
#pragma unroll
                for (int ee = 0; ee < 8; ++ee) o[ee] = fmaxf(o[ee], 0.f);
              }
              if (want_amax) {
#pragma unroll
                for (int ee = 0; ee < 8; ++ee) amax = fmaxf(amax, fabsf(o[ee]));
              }
              uint4 u;
              if (p.out_f16) {
                u.x = tc::cvt_f16x2_sat(o[0], o[1]); u.y = tc::cvt_f16x2_sat(o[2], o[3]);
                u.z = tc::cvt_f16x2_sat(o[4], o[5]); u.w = tc::cvt_f16x2_sat(o[6], o[7]);
              } else {
                u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]); u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
              }
              if (TST) {
                const uint32_t chunk = (uint32_t)(c0 / 8 + g8);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst_s + ((chunk ^ (uint32_t)(r & 7)) << 4)),
                             "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                             : "memory");
              } else if (wide && CH == 32) {
                // 32-byte stores (STG.256): a lane writes whole 32-byte sectors, half as many store instructions
                if (g8 & 1) {
                  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst_g + (g8 - 1) * 8), "r"(held.x),
                               "r"(held.y), "r"(held.z), "r"(held.w), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w)
                               : "memory");
                } else {
                  held = u;
                }
              } else {
                *reinterpret_cast<uint4*>(dst_g + g8 * 8) = u;
              }
            }
          }
        };
        constexpr int T = NCW * MT;
        uint32_t rawA[32], rawB[32];
        if constexpr (EPI == 1 && !TST && (MT % 2 == 0) && NCW == 1) {
          // inference tail (nothing stored per channel): TWO tiles per pass share the per-channel constants (5 shared-memory
          // loads per 4 channels for both) and give the scheduler two independent dot-product chains per class - with two
          // epilogue warps per scheduler this loop is latency-bound, not issue-bound
          const int c0 = half * CH;
#pragma unroll 1
          for (int mt = 0; mt < MT; mt += 2) {
            issue(0, mt, rawA);
            issue(0, mt + 1, rawB);
            tc::tmem_ld_wait_dep(rawA);
            tc::tmem_ld_wait_dep(rawB);
            float t[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}}, u[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
            for (int i = 0; i < CH; i += 4) {
              const float4 sc = *reinterpret_cast<const float4*>(&s_scale[c0 + i]), sh = *reinterpret_cast<const float4*>(&s_shift[c0 + i]);
              const float4 w0 = *reinterpret_cast<const float4*>(&s_w3[c0 + i]), w1 = *reinterpret_cast<const float4*>(&s_w3[BN + c0 + i]);
              const float4 w2 = *reinterpret_cast<const float4*>(&s_w3[2 * BN + c0 + i]);
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const uint32_t* raw = h2 ? rawB : rawA;
                const float a0 = fmaxf(fmaf(__uint_as_float(raw[i]), sc.x, sh.x), 0.f), a1 = fmaxf(fmaf(__uint_as_float(raw[i + 1]), sc.y, sh.y), 0.f);
                const float a2 = fmaxf(fmaf(__uint_as_float(raw[i + 2]), sc.z, sh.z), 0.f), a3 = fmaxf(fmaf(__uint_as_float(raw[i + 3]), sc.w, sh.w), 0.f);
                t[h2][0] = fmaf(a2, w0.z, fmaf(a0, w0.x, t[h2][0])); u[h2][0] = fmaf(a3, w0.w, fmaf(a1, w0.y, u[h2][0]));
                t[h2][1] = fmaf(a2, w1.z, fmaf(a0, w1.x, t[h2][1])); u[h2][1] = fmaf(a3, w1.w, fmaf(a1, w1.y, u[h2][1]));
                t[h2][2] = fmaf(a2, w2.z, fmaf(a0, w2.x, t[h2][2])); u[h2][2] = fmaf(a3, w2.w, fmaf(a1, w2.y, u[h2][2]));
              }
            }
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
              for (int k3 = 0; k3 < 3; ++k3) part[EPI ? mt + h2 : 0][k3] = t[h2][k3] + u[h2][k3];
          }
        } else {
          issue(0, 0, rawA);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const int cw = t / MT, mt = t % MT;
            if (t & 1) {
              tc::tmem_ld_wait_dep(rawB);
              if (t + 1 < T) issue((t + 1) / MT, (t + 1) % MT, rawA);
              process(cw, mt, rawB);
            } else {
              tc::tmem_ld_wait_dep(rawA);
              if (t + 1 < T) issue((t + 1) / MT, (t + 1) % MT, rawB);
              process(cw, mt, rawA);
            }
            if (NCW > 1 && want_stats && mt == MT - 1) flush(cw);
          }
        }
        if (NCW == 1 && want_stats && ++pending == kFlush) {
          flush(0);
          pending = 0;
        }
        // this accumulator set may be overwritten once every participating epilogue warp has drained it
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[ab]));
        if (EPI == 1) {
          // combine the two 32-channel halves of every pixel: the upper-half warp hands its partial sums over through
          // shared memory (double-buffered by item parity: ONE 64-thread barrier per quarter and item), the lower-half
          // warp adds the residual d1 and the bias and writes the three logit planes
          float* xq = s_xch + (((li & 1u) * 4 + q) * MT) * 96;
          if (half == 1) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              xq[mt * 96 + lane] = part[EPI ? mt : 0][0];
              xq[mt * 96 + 32 + lane] = part[EPI ? mt : 0][1];
              xq[mt * 96 + 64 + lane] = part[EPI ? mt : 0][2];
            }
          }
          tc::named_bar_sync(2 + q, 64);
          if (half == 0) {
            const float bb0 = __ldg(p.b3), bb1 = __ldg(p.b3 + 1), bb2 = __ldg(p.b3 + 2);
            const long long HW = (long long)p.H * p.W;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const int gy = by * (16 * MT) + 16 * mt + ty;
              if (gx < p.W && gy < p.H) {
                const long long hw = (long long)gy * p.W + gx;
                const float4 d = dres[EPI ? mt : 0];
                float* o = p.out + (long long)b * 3 * HW + hw;
                o[0] = d.x + bb0 + (part[EPI ? mt : 0][0] + xq[mt * 96 + lane]);
                o[HW] = d.y + bb1 + (part[EPI ? mt : 0][1] + xq[mt * 96 + 32 + lane]);
                o[2 * HW] = d.z + bb2 + (part[EPI ? mt : 0][2] + xq[mt * 96 + 64 + lane]);
              }
            }
          }
        }
        if (TST && do_store) {
          // the store of item li-1 has finished reading its buffer before anyone passes this barrier, so item li+1 may
          // overwrite that buffer without a second barrier; the store of item li is issued right after it
          tc::fence_proxy_async_smem();
          if (e == 0 && lane == 0) tc::tma_store_wait_read<0>();
          tc::named_bar_sync(1, 256);
          if (e == 0 && lane == 0) {
            tc::tma_store_4d(&tmY, stg_base + (li & 1u) * C::STG_BYTES, n0, bx * 8, by * (16 * MT), b);
            tc::tma_store_commit();
          }
        }
      }
      if (TST && do_store && e == 0 && lane == 0) tc::tma_store_wait<0>();
      // fp16 saturation is loud, not silent: the largest stored magnitude goes out only when it exceeded the fp16 range
      if (want_amax && amax > 65504.f) atomicMax(reinterpret_cast<unsigned int*>(p.amax), __float_as_uint(amax));
      if (want_stats) {
        if (NCW == 1 && pending > 0) flush(0);
#pragma unroll
        for (int cw = 0; cw < NCW; ++cw) {
          const int c0 = ((C::NCHUNK >= 2) ? 2 * cw + half : 0) * CH;
          if (lane < CH) {
            atomicAdd(p.stats + n0 + c0 + lane, tot1[cw]);
            atomicAdd(p.stats + p.Cout + n0 + c0 + lane, tot2[cw]);
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int KC, int BN, int MT, bool WRES, bool TST = false, int EPI = 0>
static int launch_halo(const void* x, int ldx, const void* w, ConvHaloParams p, cudaStream_t st) {
  using C = HaloCfg<KC, BN, MT, WRES, TST>;
  const int cchunks = p.Cin / KC;
  const int smem = C::smem_bytes(cchunks);
  if (smem > 227 * 1024) return 1;   // does not fit: caller falls back to the per-tap kernel
  p.blocks_x = (p.W + 7) / 8;
  p.blocks_y = (p.H + 16 * MT - 1) / (16 * MT);
  const long long items = (long long)p.blocks_x * p.blocks_y * p.B;
  if (items > 0x7fffffffLL) return 1;
  p.items = (int)items;
  CUtensorMap tmX, tmW, tmY;
  if (TST && p.y != nullptr) {
    uint64_t dims[4] = {(uint64_t)p.Cout, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {(uint64_t)p.ldy * 2, (uint64_t)p.ldy * 2 * p.W, (uint64_t)p.ldy * 2 * p.W * p.H};
    uint32_t box[4] = {(uint32_t)BN, 8u, (uint32_t)(16 * MT), 1u};
    if (tc::encode_tensor_map_bf16(&tmY, p.y, 4, dims, str, box, 128)) return -1;   // 2-byte elements (bf16 or fp16 bits)
  } else {
    memset(&tmY, 0, sizeof(tmY));
  }
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * p.W, (uint64_t)ldx * 2 * p.W * p.H};
    uint32_t box[4] = {(uint32_t)KC, 10u, (uint32_t)(16 * MT + 2), 1u};
    if (tc::encode_tensor_map_bf16(&tmX, x, 4, dims, str, box, KC == 64 ? 128 : 32)) return -1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * p.Cin, (uint64_t)p.Cout}, str[1] = {(uint64_t)9 * p.Cin * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (tc::encode_tensor_map_bf16(&tmW, w, 2, dims, str, box, KC == 64 ? 128 : 32)) return -1;
  }
  auto kern = conv3x3_halo_kernel<KC, BN, MT, WRES, TST, EPI>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_halo: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e));
  }
  const int ntiles = p.Cout / BN;
  int per = kNumSMs / ntiles;
  if (per < 1) per = 1;
  if (per > p.items) per = p.items;
  dim3 grid((unsigned)ntiles, (unsigned)per);
  kern<<<grid, 320, smem, st>>>(tmX, tmW, tmY, p);
  return check_launch("conv3x3_halo");
}

// returns 0 = launched, 1 = shape not covered (caller uses the per-tap kernel), < 0 = error
int conv3x3_fwd_halo_bf16(const void* x, int ldx, const void* w, void* y, int ldy, int B, int H, int W, int Cin, int Cout,
                          double* stats, const float* scale, const float* shift, int relu, int out_raw, int f16, float* amax,
                          cudaStream_t st) {
  ConvHaloParams p;
  p.y = y; p.ldy = ldy; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.blocks_x = p.blocks_y = p.items = 0;
  p.stats = stats; p.scale = scale; p.shift = shift; p.relu = relu; p.out_f16 = (out_raw || f16) ? 1 : 0;
  p.fmt16 = f16 ? 0u : 1u; p.amax = amax; p.a_ahead = g_opt_a_ahead;
  p.w3 = p.b3 = p.d14 = nullptr; p.out = nullptr;
  if (H < 8 || W < 8) return 1;
  if (y == nullptr && !(Cin == 16 && Cout == 64)) return 1;   // statistics-only pass: only the 16 -> 64 kernel skips stores           // tiny images: the batch-folding per-tap kernel wastes less
  if (Cin % 64 == 0) {
    if (Cout % 256 == 0) {
      // BN = 256: one 128-row tile per item leaves room for two accumulator sets (2 x 256 columns), so the statistics
      // epilogue overlaps the next item's MMAs; without statistics (dgrad) the per-tap kernel is already at ~0.85 of peak
      return launch_halo<64, 256, 1, false>(x, ldx, w, p, st);
    }
    // N = 192 (option bn192, off by default; dgrad of dec2.0: 64 -> 192 channels, = 2 also dgrad of dec3.0, 384 = 2 x 192):
    // a 128 x 64 x 16 MMA reads 6 KB of operands in 32 tensor clocks - bound by the 128 B/clk shared-memory port at 2/3 of
    // peak - while 128 x 192 x 16 reads 10 KB in 96.  Measured SLOWER all the same (1.37 vs 1.16 ms): with one 128-pixel
    // tile per item the 48 KB of output go out as per-thread 16-byte stores (32 lines per warp instruction), where the three
    // N = 64 CTAs use the TMA-store epilogue; a staged epilogue for N = 192 does not fit beside the filter ring.
    if (g_opt_bn192 >= 1 && Cout % 192 == 0 && (Cout % 128 != 0 || g_opt_bn192 >= 2)) return launch_halo<64, 192, 1, false>(x, ldx, w, p, st);
    if (Cout % 128 == 0) return launch_halo<64, 128, 2, false>(x, ldx, w, p, st);
    if (Cout % 64 == 0) {
      if (Cin == 64) return g_opt_tma_store ? launch_halo<64, 64, 2, true, true>(x, ldx, w, p, st)
                                            : launch_halo<64, 64, 2, true>(x, ldx, w, p, st);
      return launch_halo<64, 64, 4, false>(x, ldx, w, p, st);
    }
    if (Cout == 16 && Cin == 64) return launch_halo<64, 16, 4, true>(x, ldx, w, p, st);
    return 1;
  }
  if (Cin == 16 && Cout % 64 == 0 && Cout % 128 != 0)
    return g_opt_tma_store ? launch_halo<16, 64, 4, true, true>(x, ldx, w, p, st) : launch_halo<16, 64, 4, true>(x, ldx, w, p, st);
  return 1;
}

// Fused forward of the 2Hx2W tail (see ConvHaloParams): 3x3 convolution of the padded d1 (16 -> 64), folded / batch
// BatchNorm affine + ReLU, the 1x1 enhance.3 projection, residual and bias, straight from the TMEM accumulators.
// mid != nullptr additionally stores the RAW fp16 convolution output (the training backward reads it).
int conv3x3_tail_fwd_bf16(const void* x16, const void* w, void* mid, const float* scale, const float* shift, const float* w3,
                          const float* b3, const float* d14, float* out, int B, int H, int W, int f16, cudaStream_t st) {
  ConvHaloParams p;
  p.y = mid; p.ldy = 64; p.B = B; p.H = H; p.W = W; p.Cin = 16; p.Cout = 64;
  p.blocks_x = p.blocks_y = p.items = 0;
  p.stats = nullptr; p.scale = scale; p.shift = shift; p.relu = 0; p.out_f16 = 1;
  p.fmt16 = f16 ? 0u : 1u; p.amax = nullptr; p.a_ahead = g_opt_a_ahead;
  p.w3 = w3; p.b3 = b3; p.d14 = d14; p.out = out;
  if (H < 8 || W < 8) return 1;
  if (mid != nullptr) return launch_halo<16, 64, 4, true, true, 1>(x16, 16, w, p, st);
  return launch_halo<16, 64, 4, true, false, 1>(x16, 16, w, p, st);
}

}  // namespace eunet

extern "C" int eunet_conv3x3_tail_fwd(const void* d1p16, const void* w_packed, void* mid_raw, const float* scale, const float* shift,
                                      const float* w3, const float* b3, const float* d14, float* out, int dtype, int B, int H2,
                                      int W2, void* stream) {
  using namespace eunet;
  EUNET_REQUIRE(B > 0 && H2 >= 8 && W2 >= 8, "conv3x3_tail_fwd: needs B > 0 and a >= 8x8 output grid (got %d, %dx%d)", B, H2, W2);
  EUNET_REQUIRE(dtype == EUNET_BF16 || dtype == EUNET_F16, "conv3x3_tail_fwd: tensor-core path only (dtype %d)", dtype);
  EUNET_REQUIRE(d1p16 && w_packed && scale && shift && w3 && b3 && d14 && out, "conv3x3_tail_fwd: null operand");
  EUNET_REQUIRE((reinterpret_cast<uintptr_t>(d14) & 15) == 0 && (reinterpret_cast<uintptr_t>(mid_raw) & 15) == 0,
                "conv3x3_tail_fwd: d14 / mid must be 16-byte aligned");
  const int rc = conv3x3_tail_fwd_bf16(d1p16, w_packed, mid_raw, scale, shift, w3, b3, d14, out, B, H2, W2, dtype == EUNET_F16,
                                       (cudaStream_t)stream);
  EUNET_REQUIRE(rc <= 0, "conv3x3_tail_fwd: shape not covered");
  return rc;
}
