/* eunet.h - C ABI of libeunet_b200.so: the B200-native (sm_100a) kernels behind the Enhanced-UNet hot path.
 *
 * The reference (whh1747012859/Enhanced-UNet) is pure PyTorch and has NO FFI / plugin layer for this
 * path: the arithmetic is reached through torch.nn modules.  Each entry point below
 * therefore cites the reference Python interface whose arithmetic it replaces (file:line in
 * /root/reference); INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory owned by the caller (the PyTorch
 *    caching allocator in our host code).  The library never allocates, frees or retains pointers.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it asynchronously.
 *  - return 0 on success, negative on error; eunet_last_error() returns the (thread-local) message.
 *    There is no CPU fallback and no silent degradation.
 *  - activations are NHWC ("channels last"), dtype EUNET_F16 (default: fp16 tensors - 11 mantissa bits keep train-mode
 *    BatchNorm inside the 2e-2 logit tolerance where bf16's 8 do not), EUNET_BF16 or EUNET_F32 ("fp32 mode"),
 *    addressed as base pointer + `ld` = elements per pixel of the underlying buffer, so a channel
 *    slice of a wider (concat) buffer is a first-class operand.  Channel counts are multiples of 8
 *    (bandwidth kernels) / 16 (convolutions).
 *  - packed 3x3 filters: [Cout][9][Cin] (tap = ky*3+kx, Cin contiguous) in the activation dtype.
 *  - tcgen05 takes both operands of an MMA in ONE 16-bit format (fp16 x bf16 is an illegal instruction on B200), so
 *    in EUNET_F16 mode gradients are fp16 too.  Their range is kept by one power-of-two scale S per backward pass
 *    (eunet_grad_scale: max |dLoss/dlogits| -> 2^4): the entry points that read the caller's fp32 logit gradient
 *    multiply by S (`gscale[0]`), the entry points that write fp32 parameter gradients multiply by 1/S (`gscale[1]`);
 *    everything in between is linear.  `gscale` = NULL means S = 1 (bf16 / fp32 modes).
 */
#ifndef EUNET_H_
#define EUNET_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EUNET_ABI_VERSION 2
#define EUNET_F32 0
#define EUNET_BF16 1
#define EUNET_F16 2

const char* eunet_last_error(void);
int eunet_abi_version(void);
int eunet_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin);
/* tuning / A-B switches for benchmarking and tests (process-global; defaults are the product path):
 *   "conv_halo"    1 (default) = halo-tile persistent conv kernels where they apply; 0 = always the per-tap kernels;
 *                  2 = halo kernels for every shape they cover (tests); 5 = per-tap wgrad without row-halo X boxes
 *   "cta_pair"     0 (default) = off; 1 = CTA-pair (cta_group::2, M = 256) halo kernel for the N = 128 conv tiles; 2 = also N = 64
 *                  (correct, but measured 1.6x slower than the single-CTA kernels on B200 - csrc/conv_halo2.cu)
 *   "tma_store"    1 (default) = Cout = 64 halo kernels write their tile with a TMA tensor store; 0 = per-thread stores
 *   "tail_out_tma" 1 (default) = TMA-pipelined tail_out_fwd / tail_bwd_reduce / tail_dec1_*; 0 = the cp.async-ring kernels
 *   "bn_tma"       1 (default) = TMA load -> transform -> TMA store bn_apply_relu; 0 = the cp.async-ring kernel
 *   "a_ahead"      1 (default) = halo conv requests activation tiles as soon as their slot is free; 0 = after the previous
 *                  chunk's last filter tap (measured equal)
 *   "bn192"        0 (default) = off; 1 = N = 192 tiles for Cout = 192 (dgrad of dec2.0); 2 = also Cout = 384 (measured slower)
 *   "tail_dbg"     0 (default); timing experiments on tail_bwd_fused: bit 1 skip transform, 2 skip wgrad MMAs, 4 skip U^T MMAs,
 *                  8 skip drain (results invalid), 16 = fp32 transform arithmetic in fp16 mode (valid results) */
int eunet_set_option(const char* name, int value);

/* ---- metrics.py:12-58 (calculate_iou / calculate_dice / calculate_semantic_metrics) and the confusion
 * counts of visualization.py:294-311, 1484-1492.  Per image i: counts[i][g][p] = #pixels with gt class g
 * and predicted class p, classes {0,1,2} and 3 = "any other value" (e.g. ignore label 255).  Integer,
 * bit-exact.  elem_bytes in {1,4,8} (uint8 / int32 / int64 masks; the reference passes int64). */
int eunet_confusion4x4(const void* pred, const void* gt, int elem_bytes, long long n_images, long long px_per_image,
                       long long* counts /* [n_images][4][4], overwritten */, void* stream);

/* ---- instance matching support (metrics.py:61-194 calculate_instance_metrics; IoU of every prediction / ground-truth
 * pair, metrics.py:95-101): masks [n][hw] uint8 (non-zero = member) -> bit planes [n][(hw+31)/32] + areas [n] (int64);
 * inter[p][g] = |a_p AND b_g| for all pairs from the bit planes.  Integer, bit-exact. */
int eunet_pack_mask_bits(const unsigned char* masks, int n, long long hw, unsigned int* bits, long long* area, void* stream);
int eunet_pair_intersections(const unsigned int* a_bits, int na, const unsigned int* b_bits, int nb, long long words,
                             long long* inter /*[na][nb]*/, void* stream);

/* ---- layout / parameter packing (host glue of models.py:227-238: NCHW fp32 tensors at the boundary) ---- */
/* x [B,C,H,W] fp32 -> NHWC with Cpad channels (zero padded), dtype.
 * split_hilo (C == 3): channels {0-2, 3-5, 6-8} = {hi, lo, hi} with hi = dtype(x), lo = dtype(x - hi); together with
 * filters packed in mode 2 the first convolution then sees ~16 mantissa bits of the input and of its weights. */
int eunet_pack_input_nchw(const float* x, void* out, int dtype, int B, int C, int H, int W, int Cpad, int split_hilo,
                          void* stream);
/* w [Co,Ci,3,3] fp32 (nn.Conv2d.weight) -> packed [CoPad][9][CiPad] dtype.
 * transpose_flip = 0: forward form out[co][tap][ci] = w[co][ci][tap];
 * transpose_flip = 1: dgrad form   out[ci][8-tap][co] = w[co][ci][tap] (so dgrad is a forward conv over dY);
 * transpose_flip = 2: forward form for a hi/lo split 3-channel input: input channels {0-2, 3-5, 6-8} = {w_hi, w_hi, w_lo}. */
int eunet_pack_weight3x3(const float* w, void* out, int dtype, int Co, int Ci, int CoPad, int CiPad, int transpose_flip,
                         void* stream);
/* the same packing for up to 32 filter tensors in ONE launch (all arrays have `count` entries; mode = transpose_flip) */
int eunet_pack_weight3x3_multi(const void* const* w, void* const* out, const int* co, const int* ci, const int* copad,
                               const int* cipad, const int* mode, int count, int dtype, void* stream);
/* dw_packed [Co][9][CiPad] fp32 -> dw [Co,Ci,3,3] fp32 (layout of nn.Conv2d.weight.grad); hilo: sum the x_hi / x_lo channels */
int eunet_unpack_wgrad3x3(const float* dw_packed, float* dw, int Co, int Ci, int CiPad, int hilo, const float* gscale,
                          void* stream);
/* the same unpacking for up to 32 filter gradients in ONE launch (all arrays have `count` entries) */
int eunet_unpack_wgrad3x3_multi(const void* const* dw_packed, void* const* dw, const int* co, const int* ci, const int* cipad,
                                const int* hilo, int count, const float* gscale, void* stream);

/* ---- nn.Conv2d(k=3, padding=1) forward (models.py:219,222,309) and its autograd dgrad/wgrad
 * (loss.backward(), train_eval.py:338).  bf16: tcgen05/TMEM implicit GEMM with TMA-staged tiles;
 * fp32: CUDA-core direct convolution ("fp32 mode").
 * y[p,co] = sum_{tap,ci} x[p+tap,ci] * w[co][tap][ci], zero padding.  Epilogue options:
 *   stats != NULL: accumulate per-channel sum / sum of squares of the fp32 results into stats[0..Cout)
 *                  and stats[Cout..2Cout) (double, caller zeroes) - the BatchNorm batch statistics.
 *   scale/shift != NULL: y = y*scale[co] + shift[co] (folded eval-mode BN and/or bias); relu: max(y,0).
 *   out_raw: element type of y.  0 = the activation dtype (bf16 / fp32); 1 = the RAW dtype used for tensors
 *            that feed a BatchNorm (fp16 in bf16 mode - 8x finer than bf16 where the batch mean dominates -
 *            and fp32 in fp32 mode).  The bn_* / tail_* entry points read raw tensors in that dtype.
 *   y == NULL (16-bit modes, Cin = 16 -> Cout = 64 only): statistics-only pass, nothing is stored.
 *   amax (may be NULL): fp16 outputs saturate at +-65504 instead of overflowing; when a stored magnitude exceeded that
 *                  range the largest one is written here (atomic max of the fp32 bits; caller zeroes) so that the host
 *                  can raise instead of training on clamped values. */
int eunet_conv3x3_fwd(const void* x, int ldx, const void* w_packed, void* y, int ldy, int dtype, int B, int H, int W,
                      int Cin, int Cout, double* stats, const float* scale, const float* shift, int relu, int out_raw,
                      float* amax, void* stream);
/* Fused forward of the 2Hx2W tail (models.py:309-313 enhance head + 337 residual), bf16 tensor-core path:
 *   out[b,k,p] = d14[p][k] + b3[k] + sum_c w3[k][c] * relu(conv3x3(d1p16; w_packed)[p,c] * scale[c] + shift[c])
 * d1p16: [B*H2*W2, 16] bf16 (d1 padded to 16 channels); w_packed: eunet_pack_weight3x3 of enhance.0 ([64][9][16]);
 * scale/shift: BatchNorm affine (batch statistics in training, folded running statistics in eval); d14: fp32 [pixels][4];
 * out: fp32 NCHW [B,3,H2,W2].  mid_raw != NULL additionally stores the RAW fp16 convolution output [pixels][64]
 * (read by tail_bwd_*); NULL (inference) never materialises the 64-channel tensor. */
int eunet_conv3x3_tail_fwd(const void* d1p16, const void* w_packed, void* mid_raw, const float* scale, const float* shift,
                           const float* w3, const float* b3, const float* d14, float* out, int dtype, int B, int H2, int W2,
                           void* stream);
/* dw[co][tap][ci] += sum_p dy[p,co] * x[p+tap,ci]   (fp32 accumulate into caller-zeroed dw_packed) */
int eunet_conv3x3_wgrad(const void* x, int ldx, const void* dy, int lddy, float* dw_packed, int dtype, int B, int H, int W,
                        int Cin, int Cout, void* stream);

/* Data gradient of a 3x3 convolution with <= 3 real INPUT channels (enhance.0, models.py:309; autograd dgrad of
 * loss.backward(), train_eval.py:338), bf16 tensor-core path only:
 *   dx4[p][i] = sum_{tap,co} dy[p + tap - (1,1), co] * w_packed_flip[i][tap][co],  i < 3;  dx4[p][3] = 0
 * dy: [B*H*W, 64] bf16 (row stride lddy); w_packed_flip: eunet_pack_weight3x3(..., transpose_flip = 1) output
 * [cin_pad][9][64] bf16 (cin_pad >= 8); dx4: fp32 [B*H*W][4]. */
int eunet_conv3x3_dgrad_few(const void* dy, int lddy, const void* w_packed_flip, float* dx4, int dtype, int B, int H, int W,
                            int Cout, int cin_pad, void* stream);

/* ---- nn.BatchNorm2d (models.py:220,223,310): train-mode statistics -> affine, running stats ---- */
int eunet_bn_finalize(const double* stats, long long count, const float* gamma, const float* beta, const float* conv_bias,
                      float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                      float* scale, float* shift, float* mean, float* invstd, int C, void* stream);
int eunet_bn_fold_eval(const float* gamma, const float* beta, const float* conv_bias, const float* running_mean,
                       const float* running_var, float eps, float* scale, float* shift, int C, void* stream);
/* out = relu(y*scale+shift) (BN apply + nn.ReLU, models.py:220-221); optional fused nn.MaxPool2d(2)
 * (models.py:214) writing pooled [B,H/2,W/2,C] when pooled != NULL. */
int eunet_bn_apply_relu(const void* y, int ldy, void* out, int ldo, void* pooled, int ldp, int dtype, int B, int H, int W,
                        int C, const float* scale, const float* shift, void* stream);
/* BatchNorm+ReLU backward, pass 1: g = dact * [y*scale+shift > 0]; sums[0..C) += sum g, sums[C..2C) += sum g*xhat */
int eunet_bn_bwd_reduce(const void* dact, int ldd, const void* y, int ldy, int dtype, long long M, int C, const float* scale,
                        const float* shift, const float* mean, const float* invstd, double* sums, void* stream);
/* pass 2: dy = scale*(g - sum_g/M - xhat*sum_gx/M); also writes dgamma = sum_gx, dbeta = sum_g (fp32) */
int eunet_bn_bwd_apply(const void* dact, int ldd, const void* y, int ldy, void* dy, int lddy, int dtype, long long M, int C,
                       const float* scale, const float* shift, const float* mean, const float* invstd, const double* sums,
                       float* dgamma, float* dbeta, const float* gscale, void* stream);

/* The same two passes for an encoder output that also feeds nn.MaxPool2d(2) (models.py:229-231): the gradient of the
 * activation is dskip + route(dpool) (first maximum of every 2x2 window of the stored activations, ATen tie-break), formed on
 * the fly from the RAW conv output y - no separate max-pool backward pass, no read-modify-write of the skip slice.
 * dskip: [B,H,W,C] gradient from the decoder's concat slice; dpool: [B,H/2,W/2,C] gradient of the pooled tensor. */
int eunet_bn_bwd_reduce_pool(const void* dskip, int ldd, const void* dpool, int ldp, const void* y, int ldy, int dtype, int B, int H,
                             int W, int C, const float* scale, const float* shift, const float* mean, const float* invstd,
                             double* sums, void* stream);
int eunet_bn_bwd_apply_pool(const void* dskip, int ldd, const void* dpool, int ldp, const void* y, int ldy, void* dy, int lddy,
                            int dtype, int B, int H, int W, int C, const float* scale, const float* shift, const float* mean,
                            const float* invstd, const double* sums, float* dgamma, float* dbeta, const float* gscale, void* stream);

/* ---- nn.MaxPool2d(2) (models.py:214) and nn.Upsample(x2, bilinear, align_corners=False) (models.py:215) ---- */
int eunet_maxpool2_fwd(const void* x, int ldx, void* out, int ldo, int dtype, int B, int H, int W, int C, void* stream);
/* dx (+)= route(dpool) to the first maximum of each 2x2 window (ATen tie-break) */
int eunet_maxpool2_bwd(const void* dpool, int ldp, const void* x, int ldx, void* dx, int lddx, int accumulate, int dtype,
                       int B, int H, int W, int C, void* stream);
int eunet_upsample2_fwd(const void* x, int ldx, void* out, int ldo, int dtype, int B, int H, int W, int C, void* stream);
int eunet_upsample2_bwd(const void* dout, int ldo, void* dx, int ldx, int dtype, int B, int H, int W, int C, void* stream);
/* out = upsample2(relu(y * scale + shift)): BN apply + ReLU (models.py:220-224) + nn.Upsample (models.py:215, 233-235) in one
 * pass over the RAW conv output y [B,H,W,C] -> out [B,2H,2W,C]; for blocks whose activation feeds the upsample only. */
int eunet_bn_apply_relu_upsample2(const void* y, int ldy, void* out, int ldo, int dtype, int B, int H, int W, int C,
                                  const float* scale, const float* shift, void* stream);

/* ---- the 2Hx2W tail (models.py:212, 236, 308-313, 337): dec1 1x1, final upsample, enhance head, residual.
 * z = dec1(d2) is computed at HxW (a 1x1 conv commutes with bilinear interpolation), d1 = up(z). ---- */
int eunet_tail_dec1_fwd(const void* d2, int ldd2, int dtype, const float* w1 /*[3][64]*/, const float* b1, float* z4 /*[M][4]*/,
                        long long M, void* stream);
/* d1 = up(z): d1p = activation-dtype copy padded to 16 channels (conv input), d14 = fp32 [4M][4] copy (residual; may be NULL) */
int eunet_tail_up_fwd(const float* z4, void* d1p /*[B,2H,2W,16]*/, float* d14, int dtype, int B, int H, int W, void* stream);
/* NCHW fp32 [B,3,H,W] -> pixel-major [B*H*W][4] fp32 (the gradient of the logits as the tail backward kernels read it) */
int eunet_tail_pack3(const float* src, float* dst4, int B, int H, int W, const float* gscale, void* stream);
int eunet_tail_out_fwd(const float* d14, const void* mid /*[B,2H,2W,64]*/, int dtype, const float* scale, const float* shift,
                       const float* w3 /*[3][64]*/, const float* b3, float* out /*[B,3,2H,2W]*/, int B, int H, int W,
                       void* stream);
/* dout4: pixel-major [4M][4] fp32 gradient of the logits (eunet_tail_pack3) */
int eunet_tail_bwd_reduce(const float* dout4, const void* mid, int dtype, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const float* w3, double* acc /*[128 + 192 + 3]*/, int B,
                          int H, int W, void* stream);
int eunet_tail_bwd_dmid(const float* dout4, const void* mid, void* dmid, int dtype, const float* scale, const float* shift,
                        const float* mean, const float* invstd, const float* w3, const double* acc, int B, int H, int W,
                        void* stream);
/* Fused backward of the enhance head at 2Hx2W (reference models.py:309-311 under loss.backward(), train_eval.py:338;
 * bf16 tensor-core path): replaces eunet_tail_bwd_dmid + eunet_conv3x3_wgrad
 * + eunet_conv3x3_dgrad_few for enhance.0 - the 64-channel gradient dmid is formed on chip and never written:
 *   dmid = BN/ReLU backward of (mid_raw, dout4) with the sums `acc` from eunet_tail_bwd_reduce,
 *   dw_packed[co][tap][ci] += sum_p dmid[p,co] * d1p16[p+tap,ci]      (fp32 [64][9][16], caller-zeroed),
 *   dx4[p][i] = sum_{tap,co} dmid[p+tap-(1,1),co] * w_packed_flip[i][tap][co], i < 3   (fp32 [pixels][4]). */
int eunet_tail_bwd_fused(const float* dout4, const void* mid_raw, const void* d1p16, const void* w_packed_flip,
                         const float* scale, const float* shift, const float* mean, const float* invstd, const float* w3,
                         const double* acc, float* dx4, float* dw_packed, int dtype, int B, int H2, int W2, void* stream);
/* dd1p: gradient w.r.t. d1, `dd1_stride` elements of `dtype` per pixel of which the first 3 are read
 * (16 = the padded conv3x3 dgrad output; 4 with dtype fp32 = eunet_conv3x3_dgrad_few's dx4) */
int eunet_tail_up_bwd(const void* dd1p /*[B,2H,2W,dd1_stride]*/, int dtype, int dd1_stride, const float* dout, float* dz4,
                      int B, int H, int W, const float* gscale, void* stream);
int eunet_tail_dec1_bwd(const float* dz4, const void* d2, int ldd2, void* dd2, int lddd2, int dtype, const float* w1,
                        double* acc /*[192 + 3]*/, long long M, void* stream);
/* double accumulators -> fp32 parameter gradients (times gscale[1] when gscale != NULL) */
int eunet_cast_f64_f32(const double* src, float* dst, long long n, const float* gscale, void* stream);
/* fp16 mode: gscale[0] = S = 2^floor(log2(target_max / max|g|)), gscale[1] = 1/S (S = 1 for an all-zero or non-finite g);
 * gscale: 4 floats of device memory ([2] is scratch).  g: the fp32 gradient of the logits (loss.backward()). */
int eunet_grad_scale(const float* g, long long n, float target_max, float* gscale, void* stream);

/* ---- loss (train_eval.py:37-60 FocalLoss, 134-157 dice_loss, 159-181 tversky_loss, 183-197 combined,
 * 261-337 per-sample loop, 306-310 logit resize == 2x2 mean) ---- */
int eunet_loss_fwd(const float* logits /*[B,3,2H,2W]*/, const long long* target /*[B,H,W]*/, int B, int H, int W,
                   int logits_scale /*2: logits at 2Hx2W, 1: at HxW*/, double* partial /*[B][10], zeroed by callee*/,
                   float* loss /*scalar*/, float* per_sample /*[B] or NULL*/, double* coef /*[B][8]*/, void* stream);
int eunet_loss_bwd(const float* logits, const long long* target, int B, int H, int W, int logits_scale, const double* coef,
                   const float* grad_out /*scalar, device*/, float* dlogits, void* stream);

/* ---- inference post-processing ("next" row: Evaluator._run_model_single train_eval.py:411-412 and
 * Evaluator._convert_probs_to_mask train_eval.py:455-568) ---- */
/* probs[b,c,h,w] = softmax_c(resize(logits)); logits_scale 2: logits are [B,3,2H,2W] and are 2x2-averaged first */
int eunet_softmax_probs(const float* logits, float* probs /*[B,3,H,W]*/, int B, int H, int W, int logits_scale, void* stream);
/* F.interpolate(mode='bilinear', align_corners=False) on planar fp32 [planes][Hin][Win] -> [planes][Hout][Wout]: the
 * 0.75x / 1.25x views of Evaluator._run_tta_inference (train_eval.py:441-451).  ratio = source step per destination pixel
 * exactly as ATen forms it: (float)(1.0 / scale_factor) when a scale factor was given, (float)in / out for size=. */
int eunet_resize_bilinear(const float* src, float* dst, int planes, int Hin, int Win, int Hout, int Wout, float ratio_h,
                          float ratio_w, void* stream);
/* F.pad(x, (0, Wp - W, 0, Hp - H), mode='reflect') on planar fp32 [planes][H][W] -> [planes][Hp][Wp]: the pad to multiples
 * of 32 in front of the model (train_eval.py:249-253 training, 400-406 inference). */
int eunet_reflect_pad(const float* src, float* dst, int planes, int H, int W, int Hp, int Wp, void* stream);
/* mean of the five TTA views (train_eval.py:419-453) on planar fp32 [planes][H][W]: p_hflip / p_vflip are the
 * probabilities of the horizontally / vertically flipped INPUT as the model returned them (un-flipped here by index). */
int eunet_tta_combine(const float* p_base, const float* p_hflip, const float* p_vflip, const float* p_s075, const float* p_s125,
                      float* out, int planes, int H, int W, void* stream);
/* argmax + threshold cascade + the two global pixel-ratio filters, per image; mask uint8 [B,H,W];
 * counts int32 [B][2] = (live, dead) pixel counts after the cascade and before the ratio filters (overwritten) */
int eunet_probs_to_mask(const float* probs, unsigned char* mask, int* counts, int B, int H, int W, void* stream);

/* ---- fusion blocks of the smp body (models.py:276-302, 320-328), eval mode, standalone (secondary path a10).
 * gate_fwd: f = cat[out_main, out_aux]; f *= sigmoid(bn(conv1x1(gelu(bn(conv3x3(f)))))); writes the gated f as
 * NHWC16 (dtype) for the head convolutions and res4[M][4] = fusion_residual(f).  `params`: HOST pointer to 219 floats
 * {w0[3][6][9], s1[3], h1[3], w3[6][3], s4[6], h4[6], wr[3][6], br[3]} (BN folded to scale/shift).
 * out_fwd: out[B,3,H,W] = z4 + res4 where z4 = fusion_head.11 (1x1) of the head output (eunet_tail_dec1_fwd). */
int eunet_fusion_gate_fwd(const float* out_main, const float* out_aux, const float* params, void* fg16, int dtype,
                          float* res4, int B, int H, int W, void* stream);
int eunet_fusion_out_fwd(const float* z4, const float* res4, float* out, int B, int H, int W, void* stream);

/* ---- the same blocks in TRAINING mode (batch-statistics BatchNorm, Dropout2d) and their backward (models.py:276-302,
 * 320-328 under loss.backward()).  The head's 3x3 convolutions / BatchNorms / 1x1 use the generic entry points above; the
 * attention gate has its own passes because each of its two BatchNorms needs whole-batch statistics before it applies:
 *   gate_conv_fwd : a1 [M][4] fp32 = conv3x3(cat[main, aux]; w0); stats = {sum[3], sumsq[3]} (double, caller zeroes)
 *   gate_mid_fwd  : a2 [M][8] fp32 = conv1x1(gelu(a1 * scale1 + shift1); w3); stats = {sum[6], sumsq[6]}
 *   gate_apply_fwd: fg = cat[main, aux] * sigmoid(a2 * scale2 + shift2) -> fg16 (dtype, 16 channels), res4 = fusion_residual(fg)
 *   gate_bwd      : gradient of all of it.  Inputs: dfg16 (gradient w.r.t. the gated features from the head's first
 *                   convolution, dtype, first 6 of 16 channels) and dout4 (gradient of the block output, pixel-major fp32 x4:
 *                   the residual path).  bn1 / bn2: {scale, shift, mean, invstd} of the two gate BatchNorms (device).
 *                   acc (double[219], caller zeroes): [0,6) dbeta2, [6,12) dgamma2, [12,30) dWr[3][6], [30,33) dbr,
 *                   [33,36) dbeta1, [36,39) dgamma1, [39,57) dW3[6][3], [57,219) dW0[3][6][3][3].
 *                   dmain / daux: fp32 NCHW gradients w.r.t. the two inputs (times gscale[1] when gscale != NULL).
 * `gate_w`: HOST pointer to 201 floats {w0[3][6][9], w3[6][3], wr[3][6], br[3]}.
 * channel_scale: x[b, pixel, c] *= scale_bc[b][c] in place - Dropout2d with keep / (1 - p) factors, and its backward. */
int eunet_fusion_gate_conv_fwd(const float* out_main, const float* out_aux, const float* gate_w, float* a1, double* stats, int B,
                               int H, int W, void* stream);
int eunet_fusion_gate_mid_fwd(const float* a1, const float* scale1, const float* shift1, const float* gate_w, float* a2,
                              double* stats, long long M, void* stream);
int eunet_fusion_gate_apply_fwd(const float* out_main, const float* out_aux, const float* a2, const float* scale2,
                                const float* shift2, const float* gate_w, void* fg16, int dtype, float* res4, int B, int H, int W,
                                void* stream);
int eunet_fusion_gate_bwd(const float* out_main, const float* out_aux, const float* a1, const float* a2, const float* bn1,
                          const float* bn2, const void* dfg16, int dtype, const float* dout4, const float* gate_w, float* dz1,
                          float* dz2, double* acc, float* dmain, float* daux, const float* gscale, int B, int H, int W,
                          void* stream);
int eunet_channel_scale(void* x, int ld, const float* scale_bc, int dtype, int B, long long HW, int C, void* stream);

/* ---- optimiser step (train_eval.py:120, 341-343): global-norm clip + AdamW over flat fp32 buffers ---- */
int eunet_sumsq(const float* g, long long n, double* out /*scalar, accumulates; caller zeroes*/, void* stream);
int eunet_adamw_step(float* p, const float* g, float* m, float* v, long long n, const double* gradsq /*scalar*/,
                     float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                     float grad_scale, void* stream);

/* multi-tensor forms (one launch for up to 64 tensors): HOST arrays of `count` device pointers and element counts;
 * every parameter group shares the hyper-parameters and step count. */
int eunet_sumsq_multi(const void* const* g, const long long* n, int count, double* out, void* stream);
int eunet_adamw_multi(void* const* p, const void* const* g, void* const* m, void* const* v, const long long* n, int count,
                      const double* gradsq, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                      int step, float grad_scale, const float* hyper_dev /* NULL, or eunet_adamw_prepare's output */,
                      void* stream);
/* CUDA-graph friendly bookkeeping: *step_dev += 1; hyper_out = {*lr_dev, 1 - beta1^step, sqrt(1 - beta2^step), step}.
 * With hyper_dev != NULL eunet_adamw_multi takes lr and the bias corrections from there instead of its arguments, so a
 * captured training step (enhanced_unet_b200/graph.py) advances correctly on every replay. */
int eunet_adamw_prepare(int* step_dev, const float* lr_dev, float beta1, float beta2, float* hyper_out /*[4]*/, void* stream);

/* ---- bring-up / verification: UMMA + TMA probe (tests/test_gpu_probe.py) ---- */
int eunet_probe_umma(const void* a, int a_rows, int a_cols, int a_box_rows, int a_box_cols, int a_swizzle, const void* b,
                     int b_rows, int b_cols, int b_box_rows, int b_box_cols, int b_swizzle, const void* x, const int* x_dims,
                     const int* x_box, int x_swizzle, const void* params_blob, int params_bytes, float* out_tmem,
                     void* out_smem, int smem_bytes, void* stream);

/* cluster-launch probe: `ctas` CTAs in pairs (cluster 1 x 2 x 1), out[i] = rank of CTA i inside its pair */
int eunet_probe_cluster(int* out, int ctas, int smem_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EUNET_H_ */
