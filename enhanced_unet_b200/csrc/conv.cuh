// Internal declarations shared by the convolution translation units.
#pragma once
#include <cuda_runtime.h>

namespace eunet {

// fp32 CUDA-core path (conv_direct.cu)
int conv3x3_fwd_f32(const float* x, int ldx, const float* w, float* y, int ldy, int B, int H, int W, int Cin, int Cout,
                    double* stats, const float* scale, const float* shift, int relu, cudaStream_t st);
int conv3x3_wgrad_f32(const float* x, int ldx, const float* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                      cudaStream_t st);

// v2 halo-tile kernel (conv_halo.cu); f16: operands / outputs are fp16 instead of bf16: 0 = launched, 1 = shape not covered (use the per-tap kernel), < 0 = error
int conv3x3_fwd_halo_bf16(const void* x, int ldx, const void* w, void* y, int ldy, int B, int H, int W, int Cin, int Cout,
                          double* stats, const float* scale, const float* shift, int relu, int out_raw, int f16, float* amax,
                          cudaStream_t st);
int conv3x3_wgrad_halo_bf16(const void* x, int ldx, const void* dy, int lddy, float* dw, int B, int H, int W, int Cin, int Cout,
                            int f16, cudaStream_t st);
// CTA-pair (cta_group::2) halo kernel for the N <= 128 layers (conv_halo2.cu): same return convention
int conv3x3_fwd_halo2(const void* x, int ldx, const void* w, void* y, int ldy, int B, int H, int W, int Cin, int Cout, double* stats,
                      const float* scale, const float* shift, int relu, int out_raw, int f16, float* amax, cudaStream_t st);
extern int g_opt_cta_pair;    // 0 (default): never; 1: N = 128 tiles; 2: N = 128 and N = 64 tiles
extern int g_opt_conv_halo;   // 1 (default): use the halo kernel where it applies
extern int g_opt_bn192;       // 0 (default): never; 1: N = 192 tiles for Cout = 192 (dgrad of dec2.0); 2: also Cout = 384
extern int g_opt_a_ahead;     // 1 (default): halo kernel requests activation tiles ahead of the filter ring (A/B switch)
extern int g_opt_tma_store;   // 1 (default): Cout = 64 halo kernels write their output tile with a TMA tensor store

// Power-of-two (bw, bh, bb) with bw*bh*bb == pixels minimising the number of tiles over a [B,H,W] image batch.
struct PixelTile {
  int bw, bh, bb;
  int tiles_x, tiles_y, tiles_b;
  long long tiles() const { return (long long)tiles_x * tiles_y * tiles_b; }
};
PixelTile choose_pixel_tile(int B, int H, int W, int pixels);

}  // namespace eunet
