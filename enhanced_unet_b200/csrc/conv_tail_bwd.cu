// Fused backward of the 2Hx2W enhance head (reference models.py:309-311 under loss.backward(), train_eval.py:338):
//   dmid = BatchNorm+ReLU backward of enhance.1/2 applied to W3^T dout      (was tail_bwd_dmid: 2.15 GB written)
//   dW0  = wgrad of enhance.0   (dmid x d1 patches)                          (was conv3x3_wgrad16_halo: 2.15 GB read)
//   dd1  = dgrad of enhance.0   (3 real channels)                            (was conv3x3_dgrad_few:    2.15 GB read)
// in ONE persistent kernel: the 64-channel gradient never leaves the SM.  Per 16x8-pixel tile
//   1. TMA lands the halo tiles of the raw conv output `mid` (fp16, 18x10 pixels x 64 ch, 128B swizzle), of the packed
//      logit gradient dout4 (fp32 x 4 per pixel) and of the padded d1 (16 ch, 32B swizzle);
//   2. eight TRANSFORM warps turn `mid` IN PLACE into dmid (bf16): per element
//          dmid = A + Bc v + [sc v + sh > 0] sum_k g_k (sc w3_k)      (Bc = -sc k2 invstd, A = -sc k1 - Bc mean)
//      with zeros for halo pixels outside the image (the transposed convolution's boundary condition);
//   3. the MMA warp issues, on that one staged tile,
//        U^T[(i,tap), p] = Wt . dmid^T      M = 64 (27 real), N = 184 halo rows, K = 64      (dgrad, see conv_dgrad_few.cu)
//        dW0[co, (tap, ci)] += dmid^T . X   M = 64, N = 3 x 48, K = 128 interior pixels      (wgrad, see conv_wgrad_halo.cu)
//      the same shared-memory bytes serve as a K-major operand (rows = N) for the first and as an MN-major operand
//      (rows = K, interior rows addressed with a 10-row group stride) for the second;
//   4. four warps move U^T from TMEM to shared memory and then sum the 9 taps per pixel (col2im) -> dd1 fp32x4;
//      the wgrad accumulator stays in TMEM for the CTA's whole lifetime and is added to the packed gradient at the end.
// Warp roles (576 threads): 0 TMA producer, 1 MMA issuer + TMEM owner, 4/5/8/9 drain + gather, the other twelve transform
// (the kernel is bound by the transform stage: ~7 instructions per element on 180 x 64 elements per 128 pixels).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct TailBwdParams {
  float* dx4;         // [pixels][4]
  float* dw;          // packed [64][9][16]
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* w3;    // [3][64]
  const double* acc;  // [0,64) sum g', [64,128) sum g' xhat (tail_bwd_reduce)
  int B, H, W;
  int blocks_x, blocks_y, items;
  uint32_t fmt16;     // operand format (d1p16, flipped filters, the dmid tile formed on chip): 1 = bf16, 0 = fp16
  int dbg;            // timing experiments only (option "tail_dbg"): 1 skip transform, 2 skip wgrad MMAs, 4 skip U^T MMAs, 8 skip drain,
                      // 16 fp32 transform arithmetic in fp16 mode
};
int g_opt_tail_dbg = 0;

constexpr int kTbStages = 5;
constexpr int kTbMid = 184 * 128;            // 23552: mid / dmid halo tile slot (180 rows used)
constexpr int kTbDout = 3072;                // 180 x 16 B
constexpr int kTbX = 6144;                   // 180 x 32 B
constexpr int kTbStage = kTbMid + kTbDout + kTbX;   // 32768
constexpr int kTbSPitch = 188;
constexpr int kTbSBytes = 27 * kTbSPitch * 4;
constexpr int kTbUCols = 184;                // U^T accumulator: N = 184 halo rows; TWO of them at columns [0, 184) and [184, 368)
constexpr int kTbWgCol = 368;                // first TMEM column of the wgrad accumulator (144 columns: 368 + 144 = 512 exactly)

// HMATH (fp16 operands): the transform runs in packed half2 arithmetic, see the transform branch
template <bool HMATH>
__global__ void __launch_bounds__(576, 1)
tail_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmMid, const __grid_constant__ CUtensorMap tmDout,
                      const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const TailBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t tma_full[kTbStages], xf_full[kTbStages], st_empty[kTbStages], acc_full[2], acc_empty[2], w_full, fin_bar;
  __shared__ uint32_t tmem_base_s;

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = sbase, a_base = sbase + 8192, s_base = a_base + kTbStages * kTbStage;
  uint8_t* const gen0 = smem_raw + (sbase - tc::smem_u32(smem_raw));     // generic pointer to sbase
  float* const s_gen = reinterpret_cast<float*>(gen0 + (s_base - sbase));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long M = (long long)p.B * p.H * p.W;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kTbStages; ++s) {
      tc::mbar_init(tc::smem_u32(&tma_full[s]), 1);
      tc::mbar_init(tc::smem_u32(&xf_full[s]), 12);
      tc::mbar_init(tc::smem_u32(&st_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(tc::smem_u32(&acc_full[s]), 1);
      tc::mbar_init(tc::smem_u32(&acc_empty[s]), 4);
    }
    tc::mbar_init(tc::smem_u32(&w_full), 1);
    tc::mbar_init(tc::smem_u32(&fin_bar), 1);
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmMid);
    tc::tma_prefetch_desc(&tmDout);
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tc::tmem_alloc(tc::smem_u32(&tmem_base_s), 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (tc::elect_one()) {
      const uint32_t wb = tc::smem_u32(&w_full);
      tc::mbar_expect_tx(wb, 8192);
      tc::tma_load_2d(w_base, &tmW, wb, 0, 0);
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
        const uint32_t s = it % kTbStages;
        tc::mbar_wait(tc::smem_u32(&st_empty[s]), ((it / kTbStages) & 1u) ^ 1u);
        const uint32_t fb = tc::smem_u32(&tma_full[s]);
        tc::mbar_expect_tx(fb, 180 * 128 + 180 * 16 + 180 * 32);
        const uint32_t base = a_base + s * kTbStage;
        tc::tma_load_4d(base, &tmMid, fb, 0, bx * 8 - 1, by * 16 - 1, b);
        tc::tma_load_4d(base + kTbMid, &tmDout, fb, 0, bx * 8 - 1, by * 16 - 1, b);
        tc::tma_load_4d(base + kTbMid + kTbDout, &tmX, fb, 0, bx * 8 - 1, by * 16 - 1, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (tc::elect_one()) {
      const uint32_t idesc_u = tc::make_idesc_16(64, 184, 0, 0, p.fmt16);     // both K-major
      const uint32_t idesc_w = tc::make_idesc_16(64, 48, 1, 1, p.fmt16);      // both MN-major
      tc::mbar_wait(tc::smem_u32(&w_full), 0);
      tc::tc_fence_after();
      uint32_t it = 0;
      for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
        const uint32_t s = it % kTbStages;
        tc::mbar_wait(tc::smem_u32(&xf_full[s]), (it / kTbStages) & 1u);
        tc::tc_fence_after();
        const uint32_t t_addr = a_base + s * kTbStage, x_addr = t_addr + kTbMid + kTbDout;
        // the 24 wgrad MMAs accumulate into their own TMEM columns for the CTA's whole lifetime; U^T alternates between two
        // accumulators so that the drain warps copy tile t out while tile t+1 is being multiplied
        if (!(p.dbg & 2))
#pragma unroll
        for (int j = 0; j < 8; ++j) {      // K = 16 pixels = image rows 2j, 2j+1 of the tile = halo rows 2j+1, 2j+2 (columns 1..8)
          const uint64_t adesc = tc::make_smem_desc(t_addr + (uint32_t)(((2 * j + 1) * 10 + 1) * 128), 0, 10 * 128, tc::kSwizzle128);
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint64_t bdesc = tc::make_smem_desc(x_addr + (uint32_t)((dy * 10 + j * 20) * 32), 32, 10 * 32, tc::kSwizzle32);
            tc::umma_bf16(tmem_base + kTbWgCol + dy * 48, adesc, bdesc, idesc_w, (it | j) != 0 ? 1u : 0u);
          }
        }
        const uint32_t ab = it & 1u;
        tc::mbar_wait(tc::smem_u32(&acc_empty[ab]), ((it >> 1) & 1u) ^ 1u);
        tc::tc_fence_after();
        if (!(p.dbg & 4))
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t adesc = tc::make_smem_desc(w_base + j * 32, 16, 1024, tc::kSwizzle128);
          const uint64_t bdesc = tc::make_smem_desc(t_addr + j * 32, 16, 1024, tc::kSwizzle128);
          tc::umma_bf16(tmem_base + ab * kTbUCols, adesc, bdesc, idesc_u, j != 0 ? 1u : 0u);
        }
        tc::umma_commit(tc::smem_u32(&acc_full[ab]));
        tc::umma_commit(tc::smem_u32(&st_empty[s]));
      }
      tc::umma_commit(tc::smem_u32(&fin_bar));
    }
  } else if (warp == 4 || warp == 5 || warp == 8 || warp == 9) {
    // ===================== drain + gather (see conv_dgrad_few.cu) =====================
    // M = 64 accumulator layout (cta_group::1): row 16 q + l lives in TMEM lane 32 q + l, l < 16; rows 0..26 are real, so
    // the two warps on lane quarter 0 and the two on quarter 1 drain U^T (splitting the 192 columns), then the same 128
    // threads gather one interior pixel each.  The kernel is bound by the transform stage, so this group runs its two
    // phases back to back behind ONE 128-thread named barrier per tile (U^T double-buffered in shared memory).
    const int q = warp & 3, e = (warp >> 2) - 1;           // warps 4, 5 -> e = 0; warps 8, 9 -> e = 1 (column half)
    const int row = q * 16 + lane;
    const int tid = (e * 2 + q) * 32 + lane;               // 0..127: pixel (ty, tx) = (tid / 8, tid % 8)
    const int ty = tid >> 3, tx = tid & 7;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
      float* S = s_gen + (it & 1u) * (kTbSBytes / 4);
      const uint32_t ab = it & 1u;
      tc::mbar_wait(tc::smem_u32(&acc_full[ab]), (it >> 1) & 1u);
      tc::tc_fence_after();
      if (p.dbg & 8) {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[ab]));
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < 3; ++c) {
        const int col0 = e * 96 + c * 32;
        uint32_t raw[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + ab * kTbUCols + (uint32_t)col0, raw);
        tc::tmem_ld_wait();
        if (lane < 16 && row < 27) {
          float4* dst = reinterpret_cast<float4*>(S + row * kTbSPitch + col0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (col0 + 4 * i < kTbSPitch)
              dst[i] = make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                                   __uint_as_float(raw[4 * i + 3]));
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&acc_empty[ab]));    // this accumulator is drained: tile t+2 may use it
      // everyone's part of U^T is in S[it & 1]; the other buffer was gathered by all four warps before they got here
      tc::named_bar_sync(1, 128);
      const int gx = bx * 8 + tx, gy = by * 16 + ty;
      float o[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) a += S[(i * 9 + t) * kTbSPitch + (ty + t / 3) * 10 + tx + t % 3];
        o[i] = a;
      }
      if (gx < p.W && gy < p.H)
        reinterpret_cast<float4*>(p.dx4)[((long long)b * p.H + gy) * p.W + gx] = make_float4(o[0], o[1], o[2], 0.f);
    }
  } else {
    // ===================== transform: mid (fp16) -> dmid (bf16), in place =====================
    // twelve warps: 2, 3, 6, 7, 10..17
    const int xw = warp < 4 ? warp - 2 : (warp < 8 ? warp - 4 : warp - 6);
    const int xt = xw * 32 + lane;             // 0..383
    const int j = xt & 7, r0 = xt >> 3;        // 16-byte chunk (channels 8j..8j+7) and first halo row (0..47) of this thread
    // dmid = A + Bc v + [sc v + sh > 0] sum_k g_k (sc w3_k) with A = -sc (k1 - k2 invstd mean), Bc = -sc k2 invstd:
    // 7 constants per channel (56 registers), 7 instructions per element
    // HMATH: the same expression on channel PAIRS in half2 arithmetic (v is fp16 already, the result is stored as fp16):
    //     dmid2 = fma(m2, S2, fma(Bc2, v2, A2)),  S2 = g0 ws0 + g1 ws1 + g2 ws2 (3 HFMA2),  m2 = [(v2 ^ sgn) > thr] in {0, 1}
    // = 3.5 instructions per element.  The ReLU mask stays EXACT: sc v + sh > 0 <=> sigma v > sigma tau, tau = -sh / sc,
    // sigma = sign(sc), and for an fp16 v that is v' > rd(sigma tau) with rd = the largest fp16 not above (a compare, no
    // arithmetic); only the addends are rounded to fp16 (2^-11 relative, like the stored result itself).
    float sc[HMATH ? 1 : 8], sh[HMATH ? 1 : 8], A[HMATH ? 1 : 8], Bc[HMATH ? 1 : 8], ws[3][HMATH ? 1 : 8];
    __half2 hA[HMATH ? 4 : 1], hBc[HMATH ? 4 : 1], hws[3][HMATH ? 4 : 1], hthr[HMATH ? 4 : 1];
    uint32_t hsgn[HMATH ? 4 : 1];
    if constexpr (HMATH) {
      const F8 a = load8(p.scale + j * 8), b = load8(p.shift + j * 8), m = load8(p.mean + j * 8), is = load8(p.invstd + j * 8);
      const F8 w0 = load8(p.w3 + j * 8), w1 = load8(p.w3 + 64 + j * 8), w2 = load8(p.w3 + 128 + j * 8);
      float fA[8], fBc[8], thr[8];
      uint32_t sg[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float k1 = (float)(p.acc[j * 8 + e] / (double)M), k2 = (float)(p.acc[64 + j * 8 + e] / (double)M);
        fBc[e] = -a.v[e] * k2 * is.v[e];
        fA[e] = -a.v[e] * k1 - fBc[e] * m.v[e];
        const float s_ = a.v[e], h_ = b.v[e];
        sg[e] = s_ < 0.f ? 0x8000u : 0u;
        // sigma tau rounded DOWN twice (division, conversion): the largest fp16 not above it; sc = 0: constant mask
        thr[e] = s_ == 0.f ? (h_ > 0.f ? -INFINITY : INFINITY) : __fdiv_rd(-h_, fabsf(s_));
      }
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        hA[e2] = __floats2half2_rn(fA[2 * e2], fA[2 * e2 + 1]);
        hBc[e2] = __floats2half2_rn(fBc[2 * e2], fBc[2 * e2 + 1]);
        hws[0][e2] = __floats2half2_rn(w0.v[2 * e2] * a.v[2 * e2], w0.v[2 * e2 + 1] * a.v[2 * e2 + 1]);
        hws[1][e2] = __floats2half2_rn(w1.v[2 * e2] * a.v[2 * e2], w1.v[2 * e2 + 1] * a.v[2 * e2 + 1]);
        hws[2][e2] = __floats2half2_rn(w2.v[2 * e2] * a.v[2 * e2], w2.v[2 * e2 + 1] * a.v[2 * e2 + 1]);
        hthr[e2] = __halves2half2(__float2half_rd(thr[2 * e2]), __float2half_rd(thr[2 * e2 + 1]));
        hsgn[e2] = sg[2 * e2] | (sg[2 * e2 + 1] << 16);
      }
    } else {
      const F8 a = load8(p.scale + j * 8), b = load8(p.shift + j * 8), m = load8(p.mean + j * 8), is = load8(p.invstd + j * 8);
      const F8 w0 = load8(p.w3 + j * 8), w1 = load8(p.w3 + 64 + j * 8), w2 = load8(p.w3 + 128 + j * 8);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float k1 = (float)(p.acc[j * 8 + e] / (double)M), k2 = (float)(p.acc[64 + j * 8 + e] / (double)M);
        sc[e] = a.v[e];
        sh[e] = b.v[e];
        Bc[e] = -a.v[e] * k2 * is.v[e];
        A[e] = -a.v[e] * k1 - Bc[e] * m.v[e];
        ws[0][e] = w0.v[e] * a.v[e];
        ws[1][e] = w1.v[e] * a.v[e];
        ws[2][e] = w2.v[e] * a.v[e];
      }
    }
    // this thread's four halo rows (r0 + 48 k; the last one exists only for r0 < 36) and their shared-memory offsets are
    // fixed for the whole kernel.  The loop body is BRANCH-FREE: rows outside the image were zero-filled by TMA and are
    // zeroed again by a select, so that the compiler can issue the loads of all rows up front (twelve warps per SM have
    // to hide the shared-memory latency themselves: with a bounds-check branch per row this stage ran at 2400 clk per tile
    // against ~900 clk of issue slots and bounded the kernel)
    int ry[4], rx[4];
    uint32_t off[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = r0 + 48 * k;
      ry[k] = r / 10;
      rx[k] = r - ry[k] * 10;
      off[k] = (uint32_t)(r * 128) + ((uint32_t)(j ^ (r & 7)) << 4);
    }
    const bool last_ok = r0 + 144 < 180;
    auto row = [&](uint32_t t_addr, uint32_t d_addr, int k, int gy0, int gx0) {
      const int r = r0 + 48 * k;
      const int gy = gy0 + ry[k], gx = gx0 + rx[k];
      const bool inb = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
      uint32_t h0, h1, h2, h3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h0), "=r"(h1), "=r"(h2), "=r"(h3) : "r"(t_addr + off[k]));
      float g0, g1, g2, gpad;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g0), "=f"(g1), "=f"(g2), "=f"(gpad) : "r"(d_addr + (uint32_t)(r * 16)));
      const uint32_t hw[4] = {h0, h1, h2, h3};
      if constexpr (HMATH) {
        const __half2 q0 = __float2half2_rn(g0), q1 = __float2half2_rn(g1), q2 = __float2half2_rn(g2);
        const __half2 one = __float2half2_rn(1.f);
        uint32_t u[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
          const __half2 v = *reinterpret_cast<const __half2*>(&hw[e2]);
          const uint32_t vx = hw[e2] ^ hsgn[e2];
          const __half2 msk = __hgt2(*reinterpret_cast<const __half2*>(&vx), hthr[e2]);      // 1.0 / 0.0 per channel
          const __half2 S = __hfma2(q2, hws[2][e2], __hfma2(q1, hws[1][e2], __hmul2(q0, hws[0][e2])));
          const __half2 o2 = __hfma2(msk, S, __hfma2(hBc[e2], v, hA[e2]));
          u[e2] = inb ? *reinterpret_cast<const uint32_t*>(&o2) : 0u;      // the transposed convolution's boundary condition
        }
        (void)one; (void)gpad;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_addr + off[k]), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]) : "memory");
        return;
      }
      float o[8];
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e2]));
        const float v[2] = {f.x, f.y};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = 2 * e2 + h;
          float rr = fmaf(Bc[e], v[h], A[e]);
          if (fmaf(v[h], sc[e], sh[e]) > 0.f) rr = fmaf(g2, ws[2][e], fmaf(g1, ws[1][e], fmaf(g0, ws[0][e], rr)));
          o[e] = rr;
        }
      }
      uint32_t u0, u1, u2, u3;
      if (p.fmt16) { u0 = pack_bf16x2(o[0], o[1]); u1 = pack_bf16x2(o[2], o[3]); u2 = pack_bf16x2(o[4], o[5]); u3 = pack_bf16x2(o[6], o[7]); }
      else { u0 = tc::cvt_f16x2_sat(o[0], o[1]); u1 = tc::cvt_f16x2_sat(o[2], o[3]); u2 = tc::cvt_f16x2_sat(o[4], o[5]); u3 = tc::cvt_f16x2_sat(o[6], o[7]); }
      if (!inb) u0 = u1 = u2 = u3 = 0u;          // the transposed convolution's boundary condition
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_addr + off[k]), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
    };
    uint32_t it = 0;
    for (int item = blockIdx.x; item < p.items; item += gridDim.x, ++it) {
      const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y;
      const uint32_t s = it % kTbStages;
      const uint32_t t_addr = a_base + s * kTbStage, d_addr = t_addr + kTbMid;
      tc::mbar_wait(tc::smem_u32(&tma_full[s]), (it / kTbStages) & 1u);
      if (!(p.dbg & 1)) {
        const int gy0 = by * 16 - 1, gx0 = bx * 8 - 1;
        row(t_addr, d_addr, 0, gy0, gx0);
        row(t_addr, d_addr, 1, gy0, gx0);
        row(t_addr, d_addr, 2, gy0, gx0);
        if (last_ok) row(t_addr, d_addr, 3, gy0, gx0);
      }
      tc::fence_proxy_async_smem();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(tc::smem_u32(&xf_full[s]));
    }
  }
  if (warp >= 2 && warp <= 5) {
    // the wgrad accumulator (M = 64: row 16 q + l in lane 32 q + l): warps 2..5 sit on the four TMEM lane quarters
    const int q = warp & 3, co = q * 16 + lane;
    tc::mbar_wait(tc::smem_u32(&fin_bar), 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      uint32_t raw[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(kTbWgCol + tap * 16), raw);
      tc::tmem_ld_wait();
      if (lane < 16) {
        float* dst = p.dw + ((long long)co * 9 + tap) * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) atomicAdd(dst + i, __uint_as_float(raw[i]));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace eunet

using namespace eunet;

extern "C" int eunet_tail_bwd_fused(const float* dout4, const void* mid_raw, const void* d1p16, const void* w_packed_flip,
                                    const float* scale, const float* shift, const float* mean, const float* invstd,
                                    const float* w3, const double* acc, float* dx4, float* dw_packed, int dtype, int B, int H2,
                                    int W2, void* stream) {
  EUNET_REQUIRE(B > 0 && H2 >= 8 && W2 >= 8, "tail_bwd_fused: needs B > 0 and a >= 8x8 grid (got %d, %dx%d)", B, H2, W2);
  EUNET_REQUIRE(dout4 && mid_raw && d1p16 && w_packed_flip && scale && shift && mean && invstd && w3 && acc && dx4 && dw_packed,
                "tail_bwd_fused: null operand");
  EUNET_REQUIRE(dtype == EUNET_BF16 || dtype == EUNET_F16, "tail_bwd_fused: tensor-core path only (dtype %d)", dtype);
  TailBwdParams p;
  p.fmt16 = dtype == EUNET_F16 ? 0u : 1u;
  p.dbg = g_opt_tail_dbg;
  p.dx4 = dx4; p.dw = dw_packed; p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd; p.w3 = w3; p.acc = acc;
  p.B = B; p.H = H2; p.W = W2;
  p.blocks_x = (W2 + 7) / 8;
  p.blocks_y = (H2 + 15) / 16;
  const long long items = (long long)p.blocks_x * p.blocks_y * B;
  EUNET_REQUIRE(items <= 0x7fffffffLL, "tail_bwd_fused: too many tiles");
  p.items = (int)items;
  CUtensorMap tmMid, tmDout, tmX, tmW;
  {
    uint64_t dims[4] = {64ull, (uint64_t)W2, (uint64_t)H2, (uint64_t)B};
    uint64_t str[3] = {128ull, 128ull * W2, 128ull * W2 * H2};
    uint32_t box[4] = {64u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmMid, mid_raw, 4, dims, str, box, 128)) return -1;     // fp16 bits, 2-byte elements
  }
  {
    uint64_t dims[4] = {8ull, (uint64_t)W2, (uint64_t)H2, (uint64_t)B};                    // fp32 x 4 per pixel = 8 two-byte units
    uint64_t str[3] = {16ull, 16ull * W2, 16ull * W2 * H2};
    uint32_t box[4] = {8u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmDout, dout4, 4, dims, str, box, 0)) return -1;
  }
  {
    uint64_t dims[4] = {16ull, (uint64_t)W2, (uint64_t)H2, (uint64_t)B};
    uint64_t str[3] = {32ull, 32ull * W2, 32ull * W2 * H2};
    uint32_t box[4] = {16u, 10u, 18u, 1u};
    if (tc::encode_tensor_map_bf16(&tmX, d1p16, 4, dims, str, box, 32)) return -1;
  }
  {
    uint64_t dims[2] = {64ull, 16ull * 9}, str[1] = {128ull};
    uint32_t box[2] = {64u, 64u};
    if (tc::encode_tensor_map_bf16(&tmW, w_packed_flip, 2, dims, str, box, 128)) return -1;
  }
  constexpr int SMEM = 1024 + 8192 + kTbStages * kTbStage + 2 * kTbSBytes;
  const bool hmath = p.fmt16 == 0u && !(g_opt_tail_dbg & 16);     // option tail_dbg bit 16: fp32 transform in fp16 mode (A/B)
  auto kern = hmath ? tail_bwd_fused_kernel<true> : tail_bwd_fused_kernel<false>;
  {   // set on every launch: the attribute is per device, and a process may drive more than one GPU
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "tail_bwd_fused: cudaFuncSetAttribute(%d): %s", SMEM, cudaGetErrorString(e));
  }
  const int grid = p.items < kNumSMs ? p.items : kNumSMs;
  kern<<<grid, 576, SMEM, (cudaStream_t)stream>>>(tmMid, tmDout, tmX, tmW, p);
  return check_launch("tail_bwd_fused");
}
