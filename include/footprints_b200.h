/* footprints_b200 — C ABI of the B200 (sm_100a) kernels behind the Footprints encoder-decoder.
 *
 * The reference (nianticlabs/footprints) has no FFI layer: its device work is PyTorch library calls
 * issued from footprints/network.py, footprints/training/losses.py and torch.optim.Adam.  Each entry
 * point below replaces the library calls of the cited reference lines (paths relative to the
 * reference checkout).  The Python host in footprints_b200/ binds these with ctypes
 * (footprints_b200/_lib.py); INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - POD arguments only; every pointer is a device pointer owned by the caller (PyTorch allocator);
 *     the library never allocates, frees or retains device memory.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises.
 *   - Return 0 on success, a negative FP_ERR_* otherwise; fp_last_error() gives a thread-local
 *     message.  Nothing throws or aborts.  There is no CPU fallback: without an sm_100 device the
 *     launch functions return FP_ERR_CUDA / FP_ERR_ARCH.
 *   - Activations are NHWC, 16-bit (FP_F16 or FP_BF16); accumulation is fp32 in TMEM.
 */
#ifndef FOOTPRINTS_B200_H
#define FOOTPRINTS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FP_VERSION 100

enum { FP_OK = 0, FP_ERR_ARG = -1, FP_ERR_CUDA = -2, FP_ERR_ARCH = -3, FP_ERR_UNSUPPORTED = -4 };
enum { FP_F16 = 0, FP_BF16 = 1 };
enum { FP_ACT_NONE = 0, FP_ACT_RELU = 1, FP_ACT_ELU = 2, FP_ACT_SIGMOID = 3 };
enum { FP_PAD_ZERO = 0, FP_PAD_REFLECT = 1 };

int fp_version(void);
const char* fp_last_error(void);
/* 0 if the current device is sm_100 (B200); FP_ERR_ARCH / FP_ERR_CUDA otherwise. */
int fp_check_device(void);
/* Re-read the FP_* tuning / triage environment variables.  The library reads them once, on first
 * use; a host that changes them afterwards (the variant-agreement tests do) calls this. */
int fp_reload_options(void);

/* Implicit-GEMM convolution on tcgen05 tensor cores: M = batch*ho*wo output pixels, N = cout,
 * K = kh*kw*(c0+c1).  The A operand is gathered on the fly from up to two NHWC sources, which
 * implements without materialising anything:
 *   - zero padding (torchvision resnet.py conv3x3/conv1x1, footprints/network.py:40-44),
 *   - ReflectionPad2d(1) (footprints/network.py:115,126,132,167,175),
 *   - F.interpolate(scale_factor=2, mode='nearest') of source 0 (network.py:98,154),
 *   - torch.cat([x, skip], 1) (network.py:155): channels [0,c0) from src0, [c0,c0+c1) from src1.
 * Epilogue: y = act(acc*scale[n] + shift[n] (+ residual[m][n])), i.e. folded eval-mode BatchNorm
 * + ReLU + residual add (resnet.py:92-104) or bias + ELU (network.py:127-136) or bias (+Sigmoid)
 * for the heads (network.py:176-178). */
typedef struct FpConvDesc {
  int32_t batch;
  int32_t c0, h0, w0; /* source 0: channels, stored height/width                                  */
  int32_t up0;        /* 1: source 0 is nearest-upsampled x2 (virtual input = 2*h0 x 2*w0);
                         2: source 0 is zero-inserted x2 — the data gradient of a stride-2 convolution
                            expressed as a stride-1 gather (c1 = 0, zero padding only)                 */
  int32_t c1;         /* source 1 channels (0 = no concat); stored at the virtual input size      */
  int32_t kh, kw, stride, pad_t, pad_l;
  int32_t pad_mode;   /* FP_PAD_ZERO / FP_PAD_REFLECT                                             */
  int32_t ho, wo;     /* output height/width                                                      */
  int32_t cout;       /* valid output channels                                                    */
  int32_t kpad;       /* packed K (multiple of 64, >= kh*kw*(c0+c1))                              */
  int32_t npad;       /* packed weight rows (multiple of bn, >= cout)                             */
  int32_t bn;         /* N tile: 16, 32, 64, 128 or 256                                           */
  int32_t act;        /* FP_ACT_*                                                                 */
  int32_t dtype;      /* FP_F16 / FP_BF16 for sources, weights, residual and 16-bit output        */
  int32_t out_f32;    /* 1: write fp32 rows instead of 16-bit                                     */
  int32_t out_ld;     /* output row stride in elements                                            */
  int32_t res_ld;     /* residual row stride in elements                                          */
  int32_t b_via_tma;  /* 1: weights by TMA (default); 0: weights by cp.async (debug)              */
  int32_t w_dtype_plus1; /* 0: weights have `dtype`; otherwise weight format + 1 (mixed f16 x bf16)       */
} FpConvDesc;

int fp_conv_fwd(const FpConvDesc* d, const void* src0, const void* src1, const void* w_packed,
                const float* scale, const float* shift, const void* residual, void* out, void* stream);

/* Split-precision variant ("precise" mode; closes the gap between 16-bit tensor-core operands and the reference's fp32
 * arithmetic on the layers upstream of the low-resolution heads).  Every activation is a PAIR of 16-bit tensors
 * hi + lo (hi = round16(y), lo = round16(y - hi): ~22 significant bits), weights likewise, and the product is the
 * three-term expansion  A*W ~ Ahi*Whi + Alo*Whi + Ahi*Wlo  accumulated in fp32 in ONE implicit GEMM whose K axis per
 * filter tap is [hi(src0,src1) | lo(src0,src1) | hi(src0,src1)] against weights [Whi | Whi | Wlo]
 * (fp_pack_conv_weight_split; kpad >= 3*kh*kw*(c0+c1)).  `low` carries the low halves; out_lo == NULL rounds the
 * output once (a plain 16-bit or fp32 tensor for a consumer outside the precise set).  Same gather semantics and epilogue
 * as fp_conv_fwd (general kernel only). */
typedef struct FpSplitLow {
  const void* src0_lo;
  const void* src1_lo;
  const void* residual_lo;
  void* out_lo;
} FpSplitLow;
int fp_conv_fwd_split(const FpConvDesc* d, const void* src0_hi, const void* src1_hi, const FpSplitLow* low,
                      const void* w_packed3, const float* scale, const float* shift, const void* residual_hi,
                      void* out_hi, void* stream);
/* [npad][kpad] with K index = tap*3*cin + part*cin + c: part 0,1 -> round16(w), part 2 -> round16(w - round16(w)). */
int fp_pack_conv_weight_split(const float* w_oihw, int cout, int cin, int kh, int kw, int npad, int kpad, int dtype,
                              void* w_packed3, void* stream);
/* stem (space-to-depth 4x4 taps x 16 channels): K index = tap*48 + part*16 + ch, kpad = 768. */
int fp_pack_stem_weight_split(const float* w_oihw, int cout, int npad, int dtype, void* w_packed3, void* stream);
/* fp_stem_prep / fp_maxpool3x3s2 on hi + lo pairs. */
int fp_stem_prep_split(const float* image_nchw, int batch, int h, int w, int dtype, void* out_hi, void* out_lo, void* stream);
int fp_maxpool3x3s2_split(const void* in_hi, const void* in_lo, int batch, int h, int w, int c, int dtype, void* out_hi,
                          void* out_lo, void* stream);

/* Train-mode BatchNorm of fp_bn_stats / fp_bn_apply on hi + lo pairs (precise mode's training forward). */
int fp_bn_stats_split(const void* z_hi, const void* z_lo, long long M, int C, int dtype, const float* gamma,
                      const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* scale,
                      float* shift, float* save_mean, float* save_invstd, float* workspace, void* stream);
int fp_bn_apply_split(const void* z_hi, const void* z_lo, const float* scale, const float* shift, const void* res_hi,
                      const void* res_lo, int relu, long long M, int C, int dtype, void* y_hi, void* y_lo, void* stream);

/* OIHW fp32 -> packed [npad][kpad] 16-bit, K index = (ky*kw + kx)*cpad + c (tap-major, channel-minor),
 * zero padded; cpad >= the channel-minor extent (cin, or cout when flipped) is the channel count of the
 * tensor the kernel will gather (a power of two >= 8 together with a concat partner).  flip=1 packs the
 * dgrad operand instead: rows = cin, K = (ky*kw+kx)*cpad + co with the filter rotated by 180 degrees.
 * Replaces nothing in the reference (weights there stay OIHW fp32 for cuDNN); it is the layout the UMMA
 * descriptors of fp_conv_fwd expect. */
int fp_pack_conv_weight(const float* w_oihw, int cout, int cin, int kh, int kw, int cpad, int npad, int kpad,
                        int dtype, int flip, void* w_packed, void* stream);

/* All weight operands of a network in ONE launch (training re-packs ~142 of them every step; §8 f1: this launch follows
 * fp_adam_step in the optimiser pass).  jobs_dev: device array sorted by `first`; job i produces packed elements
 * [first_i, first_{i+1}) of the concatenated output index space (`total` = end of the last job) with the semantics of
 * fp_pack_conv_weight (stem = 0) or fp_pack_stem_weight (stem = 1: kpad is 256, cpad / kh / kw / flip unused);
 * flip = 2 selects the split-precision layouts of fp_pack_conv_weight_split / fp_pack_stem_weight_split. */
typedef struct FpPackJob {
  const float* w;     /* OIHW fp32 parameter                      */
  void* out;          /* packed [npad][kpad] 16-bit               */
  int32_t cout, cin, kh, kw, cpad, npad, kpad, dtype, flip, stem;
  int64_t first;
} FpPackJob;
int fp_pack_weights_batched(const FpPackJob* jobs_dev, int njobs, long long total, void* stream);

/* Stem weight [64][3][7][7] -> packed [npad][256] for the space-to-depth formulation: the 7x7/2 conv
 * (torchvision resnet.py:197) equals a 4x4/1 conv over 2x2 space-to-depth blocks of 16 channels
 * ((dy*2+dx)*4 + c, c<3). */
int fp_pack_stem_weight(const float* w_oihw, int cout, int npad, int dtype, void* w_packed, void* stream);

/* image NCHW fp32 in [0,1] -> (x-0.45)/0.225 (footprints/network.py:50) -> space-to-depth NHWC
 * [batch][h/2][w/2][16] 16-bit. */
int fp_stem_prep(const float* image_nchw, int batch, int h, int w, int dtype, void* out_s2d, void* stream);

/* MaxPool2d(3, stride 2, pad 1) on NHWC 16-bit (torchvision resnet.py:200 via network.py:41). */
int fp_maxpool3x3s2(const void* in, int batch, int h, int w, int c, int dtype, void* out, void* stream);

/* Heads: bilinear upsample (align_corners=False, integer scale; network.py:180-181) of the two decoders'
 * [batch][h][w][2] fp32 head maps into one NCHW fp32 [batch][4][h*scale][w*scale] tensor
 * (torch.cat([mask, depth], 1), network.py:27-28).  scale 1 is a plain layout change. */
int fp_heads_to_nchw(const float* mask_head, const float* depth_head, int batch, int h, int w, int scale,
                     float* out_nchw, void* stream);

/* Output edge of the test-set inference (footprints/evaluation/inference.py:104-108 and
 * datasets/inference_dataset.py:35-39): the two full-resolution [batch][h][w][2] fp32 head maps -> ONE float16 NCHW
 * [batch][4][h][w] tensor with the sigmoid applied to the two mask channels -- exactly the array the reference saves with
 * np.save after `preds[:, 0:2] = sigmoid(preds[:, 0:2])`, `.cpu().numpy()`, `.astype(np.float16)`. */
int fp_heads_to_nchw_f16_post(const float* mask_head, const float* depth_head, int batch, int h, int w, void* out_f16,
                              void* stream);

/* Pyramid pooling block of the ground-segmentation network (preprocessing/segmentation/network.py:174-207,
 * Segmentor(use_PSP=True)): for s in (1, 2, 4, 6) AdaptiveAvgPool2d(s) -> 1x1 convolution 512 -> 128 without bias ->
 * bilinear upsample to h x w (align_corners=True), then torch.cat([x, x6, x4, x2, x1], 1).
 *   x   NHWC 16-bit [batch][h][w][512] (encoder layer4 output)        w_pool*  fp32 [128][512], the reduce.weight parameters
 *   out NHWC 16-bit [batch][h][w][1024], input of decoder.block1.pre_concat_conv.conv1 */
int fp_psp_fwd(const void* x, int batch, int h, int w, int dtype, const float* w_pool1, const float* w_pool2,
               const float* w_pool4, const float* w_pool6, void* out, void* stream);

/* Input edge of the inference callers (predict_simple.py:41-43,51-60 and datasets/inference_dataset.py:25-27,45-48):
 * `transforms.Resize((H, W), interpolation=Image.ANTIALIAS)` + `transforms.ToTensor()` on uint8 RGB frames, bit-exact with
 * Pillow's two-pass fixed-point Lanczos resampling (Resample.c: horizontal pass into a uint8 intermediate over the source
 * rows the vertical pass needs, then the vertical pass), fused with the /255 conversion to the float32 NCHW network input.
 *   src    uint8 [batch][hin][win][3] (RGB, packed)           tmp  uint8 [batch][ht][wout][4] (RGBX scratch, caller-owned)
 *   bounds int32 [out][2] = (first tap, tap count)             taps int32 [ksize][out] 22-bit fixed-point weights (tap-major),
 *                                                                   ZERO past an output's tap count (Pillow leaves them empty too)
 * as Pillow's precompute_coeffs + normalize_coeffs_8bpc produce them for the full-image box (the host builds them once per
 * geometry); the vertical bounds are relative to row y0 of the source, the first row the vertical pass reads. */
int fp_resize_u8_horizontal(const void* src, int batch, int hin, int win, int y0, int ht, int wout, const int32_t* bounds,
                            const int32_t* taps, int ksize, void* tmp, void* stream);
int fp_resize_u8_vertical_to_f32(const void* tmp, int batch, int ht, int wout, int hout, const int32_t* bounds,
                                 const int32_t* taps, int ksize, float* out_nchw, void* stream);

/* Heads (OutConvBlock, network.py:161-183): 3x3 reflection-padded convolution Cin -> 2 (+ Sigmoid) on CUDA cores
 * (N = 2 wastes the tensor-core tile).  x: NHWC 16-bit [batch][h][w][c], c a multiple of 8 <= 128; w_oihw [2][c][3][3]
 * and bias [2] are the fp32 parameters themselves; out: fp32 [batch][h][w][2] (input of fp_heads_to_nchw). */
int fp_head_conv_fwd(const void* x, const float* w_oihw, const float* bias, int batch, int h, int w, int c, int sigmoid,
                     int dtype, float* out, void* stream);
/* Data gradient of that convolution including the reflection-pad adjoint: dst[p][ci] (+)= sum W[co][ci][t]*dz8[q][co]
 * over every (q, t) whose reflected tap lands on p; then * act'(y_act) when y_act is given (ELU/ReLU derivative from
 * the producing layer's saved output).  dz8: [batch][h][w][8] 16-bit, channels 0..1 valid (fp_heads_bwd's output). */
int fp_head_conv_dgrad(const void* dz8, const float* w_oihw, int batch, int h, int w, int c, int accumulate,
                       const void* y_act, int act, int act_dtype, int grad_dtype, void* dst, void* stream);
/* Weight and bias gradient of that convolution (two-stage, deterministic). */
long long fp_head_conv_wgrad_workspace_floats(int c);
int fp_head_conv_wgrad(const void* x, const void* dz8, int batch, int h, int w, int c, int act_dtype, int grad_dtype,
                       int accumulate, float* dw_oihw, float* dbias, float* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training step.  Replaces autograd's backward of footprints/network.py (triggered at
 * footprints/training/train.py:155), train-mode BatchNorm (torchvision resnet.py:82-104 in .train()),
 * LossManager (footprints/training/losses.py:31-152) and Adam (footprints/model_manager.py:27).
 * Gradients flowing between layers are 16-bit (`grad_dtype`, bf16 by default for range); weight, bias
 * and BatchNorm gradients are fp32.  Every reduction is two-stage with a fixed summation order.
 * ------------------------------------------------------------------------------------------------ */

/* floats of scratch the column-reduction kernels (bn_stats / bn_bwd / bias_grad) need for C channels */
int fp_reduce_workspace_floats(int C);

/* Train-mode BatchNorm2d statistics of a conv output z [M][C]: batch mean / biased variance -> scale, shift
 * (y = z*scale + shift), saved mean / invstd for backward, and the running-stat update (momentum, unbiased
 * variance) exactly as torch.nn.functional.batch_norm(training=True).  running_* may be NULL. */
int fp_bn_stats(const void* z, long long M, int C, int dtype, const float* gamma, const float* beta,
                float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                float* save_mean, float* save_invstd, float* workspace, void* stream);

/* y = [relu](z*scale[c] + shift[c] (+ residual)) — BatchNorm apply + residual add + ReLU (resnet.py:92-104). */
int fp_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, int relu, long long M,
                int C, int dtype, void* y, void* stream);

/* BatchNorm backward with the following ReLU folded in: g = dy*(y>0) (relu=1) or dy; dgamma, dbeta (fp32);
 * dz = gamma*invstd*(g - mean(g) - xhat*mean(g*xhat)).  g_out (nullable) receives g, the gradient of the
 * residual branch. */
int fp_bn_bwd(const void* dy, const void* y, const void* z, const float* gamma, const float* save_mean,
              const float* save_invstd, int relu, long long M, int C, int act_dtype, int grad_dtype, void* dz,
              void* g_out, float* dgamma, float* dbeta, float* workspace, void* stream);

/* dbias[c] = sum_m dz[m][c] (Conv2d bias gradient). */
int fp_bias_grad(const void* dz, long long M, int C, int grad_dtype, float* dbias, float* workspace, void* stream);

/* LossManager.__call__ (losses.py:31-92) fused with its gradient.  pred[4]: the '1/8','1/4','1/2','1/1'
 * outputs [B][4][H][W] fp32; targets [B][H][W] fp32.  losses21: per scale {visible_ground, all_ground, depth,
 * ground_depth, their sum} then [20] = mean over scales.  grad (nullable) receives d(loss)/d(pred) for each scale.
 * workspace: 1184*16 floats. */
int fp_loss_fwd_bwd(const float* const* pred, float* const* grad, const float* depth, const float* visible,
                    const float* all_ground, const float* ground_depth, const float* moving, const float* depth_mask,
                    int batch, int h, int w, float min_depth, float max_depth, float prior_weight, float* losses21,
                    float* workspace, void* stream);

/* Adjoint of fp_heads_to_nchw for one decoder: channels [ch0, ch0+2) of dout [B][4][H][W] -> gradient of the head
 * convolution's pre-upsample output [B][h][w][8] (channels 2..7 zero), times sigmoid' when head_post (the saved
 * post-sigmoid head map) is given (network.py:177-181).  grad_scale multiplies the result: the loss scale that keeps
 * FP_F16 inter-layer gradients inside fp16's range (every later backward kernel is linear in this tensor; the optimiser
 * divides it out of the fp32 parameter gradients); 1 for FP_BF16 gradients. */
int fp_heads_bwd(const float* dout_nchw, int ch0, const float* head_post, int batch, int h, int w, int scale,
                 int grad_dtype, float grad_scale, void* dz8, void* stream);

/* Adjoint of [nearest x2] -> [concat] -> [ReflectionPad2d(1)] (network.py:154-155,115): folds the gradient of the
 * padded virtual input dvp [B][hv+2][wv+2][ccat] (produced by fp_conv_fwd run as a zero-padded full correlation
 * with the flipped weights) onto one source: channels [coff, coff+c); up=1 sums the 2x2 upsample replicas.
 * Optionally accumulates into dst and multiplies by the producing layer's activation derivative from its saved
 * output y_act (ELU: y>0 ? 1 : y+1, ReLU: y>0). */
int fp_fold_bwd(const void* dvp, int batch, int hv, int wv, int ccat, int coff, int c, int up, int accumulate,
                const void* y_act, int act, int act_dtype, int grad_dtype, void* dst, void* stream);

/* dz = dy * act'(y), n elements (multiple of 8). */
int fp_act_bwd(const void* dy, const void* y, int act, long long n, int act_dtype, int grad_dtype, void* dz,
               void* stream);
/* a += b on 16-bit gradient buffers, n elements (multiple of 8). */
int fp_grad_add(void* a, const void* b, long long n, int grad_dtype, void* stream);
/* MaxPool2d(3,2,1) backward (first maximum wins ties); h and w must be even (one thread per 2x2 input block). */
int fp_maxpool3x3s2_bwd(const void* x, const void* dy, int batch, int h, int w, int c, int act_dtype, int grad_dtype,
                        int accumulate, void* dx, void* stream);

/* Weight gradient of the convolution described by d (same gather as fp_conv_fwd) on tensor cores:
 * dw_oihw[co][ci][ky][kx] (+)= sum_m dz[m][co] * im2col(src)[m][(ky*kw+kx)*cin+ci].  stem=1: d is the 4x4
 * space-to-depth stem and dw is the original [64][3][7][7] filter. */
long long fp_conv_wgrad_workspace_bytes(const FpConvDesc* d);
int fp_conv_wgrad(const FpConvDesc* d, const void* src0, const void* src1, const void* dz, int dz_ld, int grad_dtype,
                  int stem, int accumulate, float* dw_oihw, float* workspace, void* stream);

/* torch.optim.Adam step (no weight decay / amsgrad) over a flat fp32 buffer; step is 1-based;
 * grad_scale multiplies the gradient first (1/world_size after a summing all-reduce). */
int fp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                 float beta2, float eps, int step, float grad_scale, void* stream);
/* The same update with the per-step scalars in DEVICE memory: hyper7_dev = {lr, beta1, beta2, eps, 1 - beta1^step,
 * sqrt(1 - beta2^step), grad_scale}.  For a CUDA-graph-captured training step (kernel arguments are baked into a graph;
 * the host refreshes these seven floats before each replay -- StepLR, bias correction). */
int fp_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                     const float* hyper7_dev, void* stream);

/* perf triage only: clock stamps recorded by CTA 0 of the last fp_conv_fwd when FP_DEBUG_FLAGS & 16 */
int fp_debug_read_trace(long long* out_3x2048);

#ifdef __cplusplus
}
#endif
#endif /* FOOTPRINTS_B200_H */
