// CTA-pair (cta_group::2) variant of the halo-tile 3x3 convolution (conv_halo.cu) for the N <= 128 layers.
//
// conv_halo.cu's M = 128 MMAs read 4 KB of A plus 32 N bytes of B from shared memory every N/2 clocks: 128 B/clk at N = 128 and
// 192 B/clk at N = 64 against a 128 B/clk port - those layers (enc1.3, enc2.*, dec3.*, dec2.*, 40 % of the conv FLOPs) sit at
// 39-68 % tensor-pipe utilisation by construction.  Here two CTAs of a cluster (the two SMs of a TPC) execute ONE M = 256 MMA:
// each SM stages its own halo tile (its 128 rows of A, i.e. its own pixel block) and only HALF of the filter tile, so the read
// demand per SM drops to 4 KB + 16 N bytes per N/2 clocks (96 B/clk at N = 128, 160 B/clk at N = 64).
//
// STATUS (measured on B200, profiles/conv_halo2_r2_ncu_summary.txt): numerically correct (the whole conv test matrix passes
// through it), shared-memory traffic per SM drops as predicted (l1tex throughput 61 % -> 25 %), but the tensor pipe is
// active only 35 % of the time against 65 % for the single-CTA kernel (128 -> 128 channels at 256^2: 436 us vs 265 us): the
// pair's lock-stepped pipeline (every "full" barrier collects transactions from both SMs, every accumulator hand-over needs
// both epilogues) stalls the single issuing thread.  It is therefore OFF by default (option "cta_pair") and kept as the
// starting point for the next round: deeper B staging per pair, and issuing from both CTAs' schedulers.
//
// Same work decomposition as conv_halo.cu (item = (16 MT) x 8 pixels x BN channels, one 4-D TMA halo box per 64-channel chunk,
// nine taps by descriptor shift, two TMEM accumulator sets); a PAIR takes items (2 i, 2 i + 1).  Pipeline roles per CTA:
// warp 0 TMA producer (own halo tile + own half of every filter tile; all transaction bytes are reported to the LEADER's
// "full" barriers), warp 1 of the leader = the only MMA issuer (tcgen05.mma.cta_group::2; tcgen05.commit multicasts the
// "empty" / "accumulator full" arrivals to both CTAs), warps 2..9 epilogue on the CTA's own TMEM half (arriving on the
// leader's "accumulator empty" barrier, remotely from the peer).
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "../../include/eunet.h"

namespace eunet {

struct ConvHalo2Params {
  void* y;
  int ldy;
  int B, H, W, Cin, Cout;
  int blocks_x, blocks_y, items;
  double* stats;
  const float* scale;
  const float* shift;
  int relu;
  int out_f16;
  uint32_t fmt16;
  float* amax;
};

__device__ __forceinline__ float warp_transpose_sum32p(float (&v)[32], int lane) {
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

template <int BN, int MT>
struct Halo2Cfg {
  static constexpr int KC = 64, ROWB = 128;
  static constexpr int HALO_ROWS = (16 * MT + 2) * 10;
  static constexpr int A_BYTES = HALO_ROWS * ROWB;
  static constexpr int A_SLOT = (A_BYTES + 1023) / 1024 * 1024;
  static constexpr int BH = BN / 2;                       // filter rows staged by each CTA
  static constexpr int B_BYTES = BH * ROWB;
  static constexpr int NACC = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = NACC * MT * BN <= 128 ? 128 : NACC * MT * BN <= 256 ? 256 : 512;
  static constexpr int ASTAGES = 2;
  static constexpr int BSTAGES = BN == 128 ? 16 : 12;      // deep filter ring: every refill is a cross-SM round trip
  static constexpr int NCHUNK = BN / 32;
  static constexpr int SMEM = 1024 + ASTAGES * A_SLOT + BSTAGES * B_BYTES;
  static_assert(BN == 64 || BN == 128, "CTA-pair kernel: N tiles of 64 or 128");
};

template <int BN, int MT>
__global__ void __launch_bounds__(320, 1)
conv3x3_halo2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvHalo2Params p) {
  using C = Halo2Cfg<BN, MT>;
  constexpr int AS = C::ASTAGES, BS = C::BSTAGES, NACC = C::NACC;
  constexpr int NCW = C::NCHUNK / 2;                        // column chunks per epilogue warp (two warps per lane quarter)
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[AS], a_empty[AS], b_full[BS], b_empty[BS], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_scale[BN], s_shift[BN];

  const uint32_t sbase = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = sbase, b_base = sbase + AS * C::A_SLOT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();              // 0 = leader
  // grid = (CTAs of the persistent pairs, N tiles): the pair must be adjacent in x (kernels that use 2-CTA tcgen05 are
  // rejected at launch - "cluster misconfiguration" - unless the cluster's x extent is even)
  const int n0 = blockIdx.y * BN;
  const int cchunks = p.Cin / 64;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int npair_items = (p.items + 1) >> 1;               // item pairs (the last one may hold a single valid item)

  if (p.scale != nullptr)
    for (int i = threadIdx.x; i < BN; i += blockDim.x) {
      s_scale[i] = p.scale[n0 + i];
      s_shift[i] = p.shift[n0 + i];
    }
  if (warp == 0 && lane == 0) {
    // "full" barriers live in the leader and collect one expect_tx arrival from each CTA's producer
    for (int s = 0; s < AS; ++s) { tc::mbar_init(tc::smem_u32(&a_full[s]), 2); tc::mbar_init(tc::smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < BS; ++s) { tc::mbar_init(tc::smem_u32(&b_full[s]), 2); tc::mbar_init(tc::smem_u32(&b_empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(tc::smem_u32(&acc_full[s]), 1); tc::mbar_init(tc::smem_u32(&acc_empty[s]), 16); }
    tc::mbar_fence_init();
    tc::tma_prefetch_desc(&tmX);
    tc::tma_prefetch_desc(&tmW);
  }
  if (warp == 1) tc::tmem_alloc_2sm(tc::smem_u32(&tmem_base_s), C::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();                                   // both CTAs' barriers are initialised before any remote arrival
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (tc::elect_one()) {
      uint32_t a_it = 0, b_it = 0;
      for (int pi = pair; pi < npair_items; pi += npairs) {
        int item = 2 * pi + (int)rank;
        if (item >= p.items) item = p.items - 1;            // odd tail: the peer re-reads the last item (its output is dropped)
        const int bx = item % p.blocks_x, by = (item / p.blocks_x) % p.blocks_y, b = item / (p.blocks_x * p.blocks_y);
        const int x0 = bx * 8, y0 = by * (16 * MT);
        for (int cc = 0; cc < cchunks; ++cc) {
          const uint32_t s = a_it % AS;
          tc::mbar_wait(tc::smem_u32(&a_empty[s]), ((a_it / AS) & 1u) ^ 1u);
          const uint32_t fb = tc::mapa_cluster(tc::smem_u32(&a_full[s]), 0);
          tc::mbar_expect_tx_cluster(fb, C::A_BYTES);
          tc::tma_load_4d_2sm(a_base + s * C::A_SLOT, &tmX, fb, cc * 64, x0 - 1, y0 - 1, b);
          ++a_it;
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t sb = b_it % BS;
            tc::mbar_wait(tc::smem_u32(&b_empty[sb]), ((b_it / BS) & 1u) ^ 1u);
            const uint32_t bb = tc::mapa_cluster(tc::smem_u32(&b_full[sb]), 0);
            tc::mbar_expect_tx_cluster(bb, C::B_BYTES);
            tc::tma_load_2d_2sm(b_base + sb * C::B_BYTES, &tmW, bb, tap * p.Cin + cc * 64, n0 + (int)rank * C::BH);
            ++b_it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (rank == 0 && tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_16(256, BN, 0, 0, p.fmt16);
      uint32_t a_it = 0, b_it = 0, li = 0;
      for (int pi = pair; pi < npair_items; pi += npairs, ++li) {
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
            const uint32_t sb = b_it % BS;
            tc::mbar_wait(tc::smem_u32(&b_full[sb]), (b_it / BS) & 1u);
            tc::tc_fence_after();
            const uint32_t b_addr = b_base + sb * C::B_BYTES;
            const int dy = tap / 3, dx = tap - dy * 3;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint32_t a_tap = a_addr + (uint32_t)(((16 * mt + dy) * 10 + dx) * 128);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint64_t adesc = tc::make_smem_desc(a_tap + j * 32, 16, 10 * 128, tc::kSwizzle128);
                const uint64_t bdesc = tc::make_smem_desc(b_addr + j * 32, 16, 8 * 128, tc::kSwizzle128);
                tc::umma_2sm(tmem_base + (ab * MT + mt) * BN, adesc, bdesc, idesc, (cc | tap | j) != 0 ? 1u : 0u);
              }
            }
            tc::umma_commit_2sm(tc::smem_u32(&b_empty[sb]), 3);
            ++b_it;
          }
          tc::umma_commit_2sm(tc::smem_u32(&a_empty[s]), 3);
          ++a_it;
        }
        tc::umma_commit_2sm(tc::smem_u32(&acc_full[ab]), 3);
      }
    }
  } else {
    // ===================== epilogue: 8 warps per CTA on the CTA's own 128 accumulator rows =====================
    const int e = warp - 2, q = warp & 3, half = e >> 2;
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    const bool want_stats = p.stats != nullptr, affine = p.scale != nullptr;
    const bool want_amax = p.amax != nullptr && p.out_f16 != 0;
    float amax = 0.f;
    double tot1[NCW], tot2[NCW];
#pragma unroll
    for (int cw = 0; cw < NCW; ++cw) tot1[cw] = tot2[cw] = 0.0;
    uint64_t run1[16], run2[16];
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
      tot1[cw] += (double)warp_transpose_sum32p(a, lane);
      tot2[cw] += (double)warp_transpose_sum32p(c, lane);
    };
    const uint32_t acc_empty_leader0 = tc::mapa_cluster(tc::smem_u32(&acc_empty[0]), 0);
    const uint32_t acc_empty_leader1 = tc::mapa_cluster(tc::smem_u32(&acc_empty[1]), 0);
    uint32_t li = 0;
    for (int pi = pair; pi < npair_items; pi += npairs, ++li) {
      const int item = 2 * pi + (int)rank;
      const bool item_ok = item < p.items;
      const int ic = item_ok ? item : p.items - 1;
      const int bx = ic % p.blocks_x, by = (ic / p.blocks_x) % p.blocks_y, b = ic / (p.blocks_x * p.blocks_y);
      const uint32_t ab = li % NACC;
      tc::mbar_wait(tc::smem_u32(&acc_full[ab]), (li / NACC) & 1u);
      tc::tc_fence_after();
      const int gx = bx * 8 + tx;
#pragma unroll
      for (int cw = 0; cw < NCW; ++cw) {
        const int c0 = (2 * cw + half) * 32;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          const int gy = by * (16 * MT) + 16 * mt + ty;
          const bool valid = item_ok && gx < p.W && gy < p.H;
          uint32_t raw[32];
          tc::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ab * MT + mt) * BN + c0), raw);
          tc::tmem_ld_wait();
          if (want_stats && valid) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint64_t v = tc::pack_f32x2(raw[2 * i], raw[2 * i + 1]);
              run1[i] = tc::add_f32x2(run1[i], v);
              run2[i] = tc::fma_f32x2(v, v, run2[i]);
            }
          }
          if (p.y != nullptr && valid) {
            const long long pix = ((long long)b * p.H + gy) * p.W + gx;
            uint16_t* dst_g = reinterpret_cast<uint16_t*>(p.y) + pix * p.ldy + n0 + c0;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
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
              *reinterpret_cast<uint4*>(dst_g + g8 * 8) = u;
            }
          }
        }
        if (want_stats) flush(cw);
      }
      // this accumulator set (in BOTH CTAs) may be overwritten once all 16 epilogue warps of the pair have drained it
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(ab == 0 ? acc_empty_leader0 : acc_empty_leader1);
    }
    if (want_amax && amax > 65504.f) atomicMax(reinterpret_cast<unsigned int*>(p.amax), __float_as_uint(amax));
    if (want_stats) {
#pragma unroll
      for (int cw = 0; cw < NCW; ++cw) {
        const int c0 = (2 * cw + half) * 32;
        atomicAdd(p.stats + n0 + c0 + lane, tot1[cw]);
        atomicAdd(p.stats + p.Cout + n0 + c0 + lane, tot2[cw]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();                                   // the peer may still be signalling this CTA's barriers / TMEM
  if (warp == 1) tc::tmem_dealloc_2sm(tmem_base, C::TMEM_COLS);
}

template <int BN, int MT>
static int launch_halo2(const void* x, int ldx, const void* w, ConvHalo2Params p, cudaStream_t st) {
  using C = Halo2Cfg<BN, MT>;
  p.blocks_x = (p.W + 7) / 8;
  p.blocks_y = (p.H + 16 * MT - 1) / (16 * MT);
  const long long items = (long long)p.blocks_x * p.blocks_y * p.B;
  if (items > 0x7fffffffLL || items < 2) return 1;
  p.items = (int)items;
  CUtensorMap tmX, tmW;
  {
    uint64_t dims[4] = {(uint64_t)p.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * p.W, (uint64_t)ldx * 2 * p.W * p.H};
    uint32_t box[4] = {64u, 10u, (uint32_t)(16 * MT + 2), 1u};
    if (tc::encode_tensor_map_bf16(&tmX, x, 4, dims, str, box, 128)) return -1;
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * p.Cin, (uint64_t)p.Cout}, str[1] = {(uint64_t)9 * p.Cin * 2};
    uint32_t box[2] = {64u, (uint32_t)C::BH};
    if (tc::encode_tensor_map_bf16(&tmW, w, 2, dims, str, box, 128)) return -1;
  }
  auto kern = conv3x3_halo2_kernel<BN, MT>;
  {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    EUNET_REQUIRE(e == cudaSuccess, "conv3x3_halo2: cudaFuncSetAttribute(%d): %s", C::SMEM, cudaGetErrorString(e));
  }
  const int ntiles = p.Cout / BN;
  int per = (kNumSMs / ntiles) & ~1;                         // CTA pairs: an even number of CTAs per N tile
  const int need = 2 * (int)((items + 1) / 2);
  if (per > need) per = need;
  if (per < 2) per = 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)per, (unsigned)ntiles);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;         // CTA pairs: (2k, y) and (2k + 1, y) share a TPC
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmX, tmW, p);
  if (e != cudaSuccess) {
    cudaFuncAttributes fa;
    memset(&fa, 0, sizeof(fa));
    (void)cudaGetLastError();
    (void)cudaFuncGetAttributes(&fa, kern);
    int nc = -1;
    cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
    set_error("conv3x3_halo2: launch (grid %d x %d, smem %d): %s [regs %d, static smem %zu, max dyn smem %d, max threads %d, "
              "max active clusters %d (%s)]", ntiles, per, C::SMEM, cudaGetErrorString(e), fa.numRegs, fa.sharedSizeBytes,
              fa.maxDynamicSharedSizeBytes, fa.maxThreadsPerBlock, nc, cudaGetErrorString(e2));
    (void)cudaGetLastError();
    return -1;
  }
  return check_launch("conv3x3_halo2");
}

// returns 0 = launched, 1 = shape not covered (caller uses conv_halo.cu), < 0 = error
int conv3x3_fwd_halo2(const void* x, int ldx, const void* w, void* y, int ldy, int B, int H, int W, int Cin, int Cout, double* stats,
                      const float* scale, const float* shift, int relu, int out_raw, int f16, float* amax, cudaStream_t st) {
  if (H < 8 || W < 8 || y == nullptr || Cin % 64 != 0) return 1;
  ConvHalo2Params p;
  p.y = y; p.ldy = ldy; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  p.blocks_x = p.blocks_y = p.items = 0;
  p.stats = stats; p.scale = scale; p.shift = shift; p.relu = relu; p.out_f16 = (out_raw || f16) ? 1 : 0;
  p.fmt16 = f16 ? 0u : 1u; p.amax = amax;
  if (Cout % 256 == 0) return 1;
  if (Cout % 128 == 0) return launch_halo2<128, 2>(x, ldx, w, p, st);
  if (Cout % 64 == 0) return launch_halo2<64, 4>(x, ldx, w, p, st);
  return 1;
}

}  // namespace eunet
