/*
 * rgbd_b200 — C-ABI of the B200-native ELIC_united compress/decompress hot path.
 *
 * Plain pointers and sizes only (no torch / pybind types).  Every device pointer is a raw
 * CUDA device address, every `stream` a cudaStream_t passed as void*.  All functions return
 * 0 on success or a negative RGBD_E_* code and never throw; nothing allocates device
 * memory behind the caller's back (work buffers are passed in).
 *
 * Each entry point names the reference interface it replaces (paths relative to the
 * reference repository root):
 *
 *   CompressAI/compressai/cpp_exts/rans/rans_interface.cpp   (module compressai.ans)
 *   CompressAI/compressai/cpp_exts/ops/ops.cpp               (module compressai._CXX)
 *   CompressAI/compressai/entropy_models/entropy_models.py   (quantize / build_indexes / likelihood)
 *   utils/ckbd.py                                            (checkerboard squeeze / unsqueeze)
 *   modules/transform/*.py, modules/layers/*.py              (conv / deconv / SE / ESA blocks)
 *
 * Activation layout in HBM: NHWC ("pixels"), a tensor view is (base pointer, channel
 * stride of one pixel in elements = `cstride`, first channel = `coff`, channel count).
 * Concatenations of the reference (torch.cat on dim 1) are therefore views of one wider
 * buffer that several producers write at different `coff`.
 */
#ifndef RGBD_B200_H
#define RGBD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RGBD_OK 0
#define RGBD_E_INVALID (-1)   /* bad argument */
#define RGBD_E_CUDA (-2)      /* CUDA runtime error, see rgbd_last_error() */
#define RGBD_E_CAPACITY (-3)  /* output buffer too small */
#define RGBD_E_UNSUPPORTED (-4)

#define RGBD_DT_F32 0
#define RGBD_DT_BF16 1

#define RGBD_ACT_NONE 0
#define RGBD_ACT_RELU 1
#define RGBD_ACT_LEAKY 2 /* negative slope 0.01 (nn.LeakyReLU default) */
#define RGBD_ACT_GELU 3  /* exact GELU, 0.5 v (1 + erf(v / sqrt 2)) (nn.GELU default; the Swin MLPs of STF_united) */

/* epilogue modes of the conv kernels; v = acc + bias */
#define RGBD_EPI_LINEAR 0   /* y = act(v + res)                  res optional            */
#define RGBD_EPI_GATE 1     /* y = res + mul * sigmoid(v)        res optional, mul required */
#define RGBD_EPI_BILERP 2   /* y = act(v + bilinear_upsample(res)) res = small NHWC map     */
#define RGBD_EPI_SHUFFLE2 3 /* tensor-core path only: the 16 accumulator columns are 4 output parities x 4 channels
                               (column 4*(2*py+px)+c); site (sy, sx) stores act(v) at pixels (2*sy+py, 2*sx+px), c < Cout <= 4:
                               a ConvTranspose2d(k5, s2, p2, op1) with few output channels as ONE 3x3 launch */

#define RGBD_MAX_TAPS 25

const char *rgbd_last_error(void);
int rgbd_abi_version(void);
/* number of kernels this library has launched since the last reset (bench bookkeeping) */
int64_t rgbd_launch_count(int reset);
/* adds n to that counter: a CUDA-graph replay of a captured launch list reports its kernel nodes here */
void rgbd_count_launch(int n);

/* ------------------------------------------------------------------------------------------
 * Convolution family (implicit GEMM over NHWC pixels).
 * Replaces nn.Conv2d / nn.ConvTranspose2d as used by modules/layers/conv.py:7-34,
 * modules/layers/res_blk.py:7-27, modules/transform/{analysis,synthesis,attention,context,
 * entropy}.py and compressai/layers/layers.py:162-213.
 *
 * One launch computes, for every image n and every output site (oy, ox) of an Hs x Ws
 * sub-lattice,
 *     acc[co] = sum_t sum_ci  x[n, oy*i_step + dy[t], ox*i_step + dx[t], ci] * in_scale[n,ci]
 *                              * w[wtap[t]][ci][co]
 * (out-of-range input sites contribute zero = the reference's zero padding) and stores the
 * epilogue result at output pixel (oy*o_step + o_off_y, ox*o_step + o_off_x).
 * A strided Conv2d is i_step = stride, dy = ky - pad; a ConvTranspose2d(k5,s2,p2,op1) is four
 * launches (one per output parity) with o_step = 2; ConvTranspose2d(k3,s1,p1) is dy = 1 - ky.
 * ------------------------------------------------------------------------------------------ */
typedef struct rgbd_conv_desc {
    const void *x;          /* input pixels */
    void *y;                /* output pixels */
    void *y2;               /* optional second copy of the output (same dtype), may be NULL */
    const void *w;          /* packed weights, see rgbd_conv_weight_elems() */
    const float *bias;      /* [Cout] or NULL */
    const void *res;        /* optional, dtype = x_dtype */
    const void *mul;        /* optional, dtype = x_dtype */
    const float *in_scale;  /* optional [N][Cin] per-image input-channel scale (SE folding) */
    void *sched_ws;         /* tensor-core path: one zero-initialised int32 in device memory per plan, owned by the
                               caller (the persistent kernel's tile counter; the kernel leaves it at zero) */
    int32_t N, H, W;        /* input dims */
    int32_t Cin, x_cstride, x_coff;
    int32_t Ho, Wo;         /* full output dims */
    int32_t Cout, y_cstride, y_coff;
    int32_t y2_cstride, y2_coff;
    int32_t Hs, Ws, o_step, o_off_y, o_off_x, i_step;
    int32_t ntaps;
    int32_t res_cstride, res_coff, res_H, res_W; /* res_H/W only for RGBD_EPI_BILERP */
    int32_t mul_cstride, mul_coff;
    int32_t act, epi;
    int32_t x_dtype, y_dtype; /* RGBD_DT_* */
    int32_t cout_pad;         /* packed weight row length (multiple of 16) */
    int32_t w_image_stride;   /* tensor-core path: 0 = all images share w; else image n uses the weight
                                 set that starts w_image_stride taps after image n-1's (per-image weights
                                 produced by rgbd_scale_weights: the SE gate folded into the filter) */
    int8_t dy[RGBD_MAX_TAPS], dx[RGBD_MAX_TAPS];
    int8_t wtap[RGBD_MAX_TAPS];
    int8_t _pad[1];
} rgbd_conv_desc;

/* fp32-accumulate CUDA-core path: weights fp32 [ntaps_total][Cin][cout_pad]. Used for the
 * fp32 parity mode and for the 3/1-channel image-side layers. */
int rgbd_conv_simt(const rgbd_conv_desc *d, void *stream);
/* argument check shared by both conv paths; sizeof(rgbd_conv_desc) for binding self-checks */
int rgbd_conv_validate(const rgbd_conv_desc *d);
int rgbd_conv_desc_size(void);

/* tcgen05 tensor-core path (bf16 operands, fp32 TMEM accumulators, TMA-staged NHWC tiles).
 * Weights bf16 [ntaps_total][cout_pad][cin_pad] (K-major).  The plan object owns the TMA
 * descriptors for (x, w) so that a launch is a single kernel. */
typedef struct rgbd_conv_tc_plan rgbd_conv_tc_plan;
int rgbd_conv_tc_plan_create(const rgbd_conv_desc *d, int32_t cin_pad, rgbd_conv_tc_plan **out);
int rgbd_conv_tc_run(const rgbd_conv_tc_plan *p, void *stream);
void rgbd_conv_tc_plan_destroy(rgbd_conv_tc_plan *p);

/* Fused bottleneck block on the tensor cores (bf16 NHWC in / out):
 *     y = act( res + W3 * relu( W2 (*) relu( W1 * x + b1 ) + b2 ) + b3 ),   1x1 (Cin -> 96), 3x3 (96 -> 96, pad 1), 1x1 (96 -> Cout)
 * = ResidualBottleneck (modules/layers/res_blk.py:7-27; res = x, or the output of its 1x1 skip conv) and
 * AttentionBlock.ResidualUnit (CompressAI/compressai/layers/layers.py:178-197; res = x, final_relu = 1) in ONE launch,
 * with both 96-channel intermediates kept in shared memory / TMEM.  Weights are bf16, K-major:
 *   w1 [96][Cin], w3 [Cout][128] (K padded 96 -> 128 with zeros),
 *   w2 [14 planes][96][64]: plane t < 9 = tap t (ky * 3 + kx), input channels 0-63; plane 9 + i = input channels 64-95 of
 *   tap 2 i in elements [0, 32) and of tap 2 i + 1 in elements [32, 64) of each row (zeros for the missing tap 9).
 * Cin % 64 == 0, Cmid == 96, Cout == 192; every view 16-byte aligned. */
typedef struct rgbd_rb_desc {
    const void *x, *res;
    void *y;
    const void *w1, *w2, *w3;
    const float *b1, *b2, *b3;     /* [96], [96], [Cout] or NULL */
    int32_t N, H, W;
    int32_t Cin, x_cstride, x_coff;
    int32_t Cmid, Cout;
    int32_t res_cstride, res_coff, y_cstride, y_coff;
    int32_t final_relu;
    int32_t _pad;
    void *sched_ws;                /* one zero-initialised int32 in device memory per plan: tile counter of the persistent
                                    * kernel's dynamic scheduler (left at zero by every launch), as rgbd_conv_desc.sched_ws */
} rgbd_rb_desc;
typedef struct rgbd_rb_plan rgbd_rb_plan;
int rgbd_rb_plan_create(const rgbd_rb_desc *d, rgbd_rb_plan **out);
int rgbd_rb_run(const rgbd_rb_plan *p, void *stream);
void rgbd_rb_plan_destroy(rgbd_rb_plan *p);

/* ------------------------------------------------------------------------------------------
 * Small spatial / channel ops of the transforms.
 * ------------------------------------------------------------------------------------------ */
/* SE_Block (modules/transform/attention.py:52-67): s[n,c] = sigmoid(W2 relu(W1 mean_hw x)).
 * `partial` is a work buffer of N*(nchunk*C + C + Cr) floats (partial sums, means, hidden); out scale[n,c] = s (+1 if plus_one, which
 * folds EntropyParametersEX's `x + se(x)`, modules/transform/entropy.py:75). Deterministic
 * fixed-order reduction, independent of N. */
int rgbd_se_scale(const void *x, int32_t dtype, int32_t N, int32_t HW, int32_t C, int32_t cstride,
                  int32_t coff, const float *w1, const float *w2, int32_t Cr, int32_t plus_one,
                  float *partial, int32_t nchunk, float *scale, void *stream);
/* The same in two halves, for callers that keep the table of partial sums between calls and refresh only the channels
 * that were rewritten (the Bi-CEE context buffer: the hyper-prior channels stay fixed over the 20 stages of
 * models/elic_united.py:265-348).  partial[n][chunk][pstride]: rgbd_se_partial fills the channels [c0, c1) of the view;
 * rgbd_se_gate turns the first C channels into scale[n][c]; work = N * (C + Cr) floats.  Bit-identical to rgbd_se_scale. */
int rgbd_se_partial(const void *x, int32_t dtype, int32_t N, int32_t HW, int32_t cstride, int32_t coff, int32_t c0,
                    int32_t c1, int32_t nchunk, float *partial, int32_t pstride, void *stream);
int rgbd_se_gate(const float *partial, int32_t pstride, int32_t nchunk, int32_t N, int32_t HW, int32_t C,
                 const float *w1, const float *w2, int32_t Cr, int32_t plus_one, float *work, float *scale, void *stream);
/* y = x * scale[n, c]: applies the SE gate as a separate pass for the tensor-core conv path,
 * whose A operand goes HBM -> smem by TMA without passing through registers. */
int rgbd_scale_channels(const void *x, void *y, int32_t dtype, const float *scale, int32_t N, int64_t HW,
                        int32_t C, int32_t x_cstride, int32_t x_coff, int32_t y_cstride, int32_t y_coff,
                        void *stream);
/* w_out[n][t][co][ci] = bf16(w[t][co][ci] * scale[n][ci]) for ci < Cin (0 beyond): folds a per-image
 * input-channel gate (SE_Block; EntropyParametersEX's `x + se(x)`, modules/transform/entropy.py:75) into
 * per-image copies of a tensor-core filter [taps][cout_pad][cin_pad] instead of rescaling the (much
 * larger) activation tensor; used with rgbd_conv_desc.w_image_stride = taps. */
int rgbd_scale_weights(const float *w, const float *scale, void *w_out, int32_t N, int32_t taps,
                       int32_t cout_pad, int32_t cin_pad, int32_t Cin, void *stream);
/* F.max_pool2d(kernel 7, stride 3) of ESA (attention.py:88) */
int rgbd_maxpool7s3(const void *x, void *y, int32_t dtype, int32_t N, int32_t H, int32_t W,
                    int32_t C, void *stream);
/* NCHW fp32 image -> NHWC (dtype), and back with optional clamp to [0,1] (elic_united.py:452).
 * split3 != 0 writes 3*C channels [hi | lo | hi] (hi = dtype(x), lo = dtype(x - hi)): the two-term
 * bf16 expansion the tensor-core first layer consumes against weights packed [w_hi | w_hi | w_lo].
 * split3 == 2 additionally folds 2x2 pixel blocks into channels (space-to-depth: y is [N, H/2, W/2, 4*3*C], channel
 * block 2*ry+rx = input pixel (2Y+ry, 2X+rx)), which turns the 5x5 stride-2 first conv into a 3x3 stride-1 conv with one
 * tap group (H, W even). */
int rgbd_nchw_to_nhwc(const float *x, void *y, int32_t dtype, int32_t N, int32_t C, int32_t H,
                      int32_t W, int32_t y_cstride, int32_t y_coff, int32_t split3, void *stream);
int rgbd_nhwc_to_nchw(const void *x, int32_t dtype, float *y, int32_t N, int32_t C, int32_t H,
                      int32_t W, int32_t x_cstride, int32_t x_coff, int32_t clamp01, void *stream);

/* torch.cat plumbing: copy the channel slice of one NHWC view into another (same dtype), and an
 * async memset for the zero-initialised y_hat buffers (utils/ckbd.py:37-48 `torch.zeros_like`). */
int rgbd_copy_view(const void *x, void *y, int32_t dtype, int64_t npix, int32_t C, int32_t x_cstride,
                   int32_t x_coff, int32_t y_cstride, int32_t y_coff, void *stream);
/* fp32 NHWC view -> bf16 NHWC view (the latents y stay fp32 for the quantiser; the hyper-analysis
 * h_a reads this bf16 copy on the tensor cores) */
int rgbd_cast_view_bf16(const float *x, void *y, int64_t npix, int32_t C, int32_t x_cstride, int32_t x_coff,
                        int32_t y_cstride, int32_t y_coff, void *stream);
int rgbd_zero(void *p, int64_t bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Token-side ops of SymmetricalTransFormerUnited (models/stf_united.py); tokens = pixels of NHWC views, the Linear layers
 * run as 1x1 convs (rgbd_conv_*).
 * ------------------------------------------------------------------------------------------ */
/* nn.LayerNorm over the C channels of every pixel (biased variance, eps inside the root; stf_united.py:157,193).  With
 * gather2x2 the token is PatchMerging's concatenation [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)] of four pixels of
 * C / 4 channels of the [N, Hin, Win] input (stf_united.py:236-244) and npix = N * Hin/2 * Win/2.  C <= 768. */
int rgbd_layernorm(const void *x, int32_t x_dtype, void *y, int32_t y_dtype, int64_t npix, int32_t C, int32_t x_cstride,
                   int32_t x_coff, int32_t y_cstride, int32_t y_coff, const float *gamma, const float *beta, float eps,
                   int32_t gather2x2, int32_t Hin, int32_t Win, void *stream);
/* nn.PixelShuffle(2) on NHWC: y[n, 2h+i, 2w+j, c] = x[n, h, w, 4c + 2i + j] (PatchSplit :270-273, end convs :574-582). */
int rgbd_pixel_shuffle2(const void *x, void *y, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t Cout, int32_t x_cstride,
                        int32_t x_coff, int32_t y_cstride, int32_t y_coff, void *stream);
/* (Shifted-)window multi-head attention (WindowAttention.forward :83-115 inside SwinTransformerBlock.forward :162-212):
 * qkv holds [q | k | v] (C channels each, head-major) per pixel; windows of window x window tokens on the map rolled by
 * -shift; scores q k^T * scale + relative-position bias (bias_table [(2 window - 1)^2][heads], fp32) + the -100 mask between
 * the regions of BasicLayer.forward :334-352 when shift > 0; softmax; times v; written back at the unrolled position. */
int rgbd_window_attention(const void *qkv, void *out, int32_t dtype, int32_t N, int32_t H, int32_t W, int32_t C, int32_t heads,
                          int32_t window, int32_t shift, const float *bias_table, float scale, int32_t qkv_cstride,
                          int32_t qkv_coff, int32_t out_cstride, int32_t out_coff, void *stream);

/* ------------------------------------------------------------------------------------------
 * Metrics and image export next to the path (testing/tester_united.py:92-123, utils/metrics.py:8-14).
 * ------------------------------------------------------------------------------------------ */
/* out[n] = sum over one image's n_per_image elements of (a - b)^2 (double; inputs clamped to [0, 1] first when clamp01),
 * the MSE of compute_metrics (utils/metrics.py:9-12).  work: N * n_part doubles; fixed summation order. */
int rgbd_sq_error_sums(const float *a, const float *b, int32_t N, int64_t n_per_image, int32_t clamp01, double *work,
                       int32_t n_part, double *out, void *stream);
/* One level of pytorch_msssim's _ssim (v1.0.0, the package utils/metrics.py:5 imports; not vendored by the reference):
 * 11-tap Gaussian window (sigma 1.5) applied separably without padding to x, y, x^2, y^2, xy of every [H, W] plane,
 * K = (0.01, 0.03); sums[plane] = {sum of ssim_map, sum of cs_map} over the (H - 10) x (W - 10) valid positions.
 * work: rgbd_ssim_work_elems(planes, H, W) doubles. */
int rgbd_ssim_level(const float *x, const float *y, int32_t planes, int32_t H, int32_t W, float data_range, int32_t clamp01,
                    double *work, double *sums, void *stream);
int64_t rgbd_ssim_work_elems(int32_t planes, int32_t H, int32_t W);
/* F.avg_pool2d(x, 2, padding = (H % 2, W % 2)) between MS-SSIM levels: y is [planes, (H + 1) / 2, (W + 1) / 2]. */
int rgbd_avgpool2(const float *x, float *y, int32_t planes, int32_t H, int32_t W, int32_t clamp01, void *stream);
/* saveImg (utils/IOutils.py:100-102): clamp to [0, 1], * 255, truncate -> interleaved u8 [N, crop_h, crop_w, C] of the
 * top-left crop (crop0, dataset/utils.py:84-85) of an NCHW fp32 image. */
int rgbd_quantize_u8(const float *x, uint8_t *y, int32_t N, int32_t C, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w,
                     void *stream);
/* 16-bit depth export (testing/tester_united.py:101-108): (x * scale).astype(uint16) of a [N, 1, H, W] image, cropped. */
int rgbd_quantize_u16(const float *x, uint16_t *y, int32_t N, int32_t H, int32_t W, int32_t crop_h, int32_t crop_w,
                      float scale, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gaussian-conditional entropy model, checkerboard-fused.
 * ------------------------------------------------------------------------------------------ */
/* One anchor / non-anchor coding step of the encoder: utils/ckbd.py:83-105 fused with
 * GaussianConditional.build_indexes (entropy_models.py:561-568) and
 * EntropyModel.quantize("symbols") (entropy_models.py:118-146).
 *   y      : fp32 NHWC latent, slice = channels [y_coff, y_coff+g)
 *   params : fp32 NHWC [N,H,W,2g] = (scales | means) (elic_united.py:289)
 *   parity : 0 = anchor ((h+w) odd), 1 = non-anchor ((h+w) even)
 * For each image n the g*H*(W/2) selected sites, in (c, h, w/2) row-major order, get
 *   sym[n*stream_stride + chunk_off + j] = int32(rint(y - mean)),  idx[...] = scale index,
 * and yhat (dtype) at that site = float(sym) + mean. */
int rgbd_ckbd_quantize_index(const float *y, int32_t y_cstride, int32_t y_coff,
                             const float *params, const float *scale_table, int32_t n_scales,
                             float scale_bound, int32_t N, int32_t H, int32_t W, int32_t g,
                             int32_t parity, int32_t *sym, uint8_t *idx, int64_t stream_stride,
                             int64_t chunk_off, void *yhat, int32_t yhat_dtype,
                             int32_t yhat_cstride, int32_t yhat_coff, void *stream);
/* Decoder pre-pass: indexes only (utils/ckbd.py:108-112). */
int rgbd_ckbd_index(const float *params, const float *scale_table, int32_t n_scales,
                    float scale_bound, int32_t N, int32_t H, int32_t W, int32_t g, int32_t parity,
                    uint8_t *idx, int64_t stream_stride, int64_t chunk_off, void *stream);
/* Decoder post-pass: yhat = float(sym) + mean scattered back (utils/ckbd.py:113-114). */
int rgbd_ckbd_dequant_scatter(const int32_t *sym, int64_t stream_stride, int64_t chunk_off,
                              const float *params, int32_t N, int32_t H, int32_t W, int32_t g,
                              int32_t parity, void *yhat, int32_t yhat_dtype,
                              int32_t yhat_cstride, int32_t yhat_coff, void *stream);
/* forward(): one coding step of codeOnePart (elic_united.py:94-115) without the coder:
 * yhat at the parity sites = ste_round(y - mean) + mean, and the likelihood of
 * GaussianConditional.forward (entropy_models.py:534-558) of the *dequantised* value
 * written in NCHW fp32 at lik[n, lik_coff + c, h, w]. */
int rgbd_ckbd_ste_likelihood(const float *y, int32_t y_cstride, int32_t y_coff,
                             const float *params, float scale_bound, float lik_bound, int32_t N,
                             int32_t H, int32_t W, int32_t g, int32_t parity, void *yhat,
                             int32_t yhat_dtype, int32_t yhat_cstride, int32_t yhat_coff,
                             float *lik, int32_t lik_C, int32_t lik_coff, void *stream);

/* ------------------------------------------------------------------------------------------
 * Factorised-prior entropy bottleneck (z).
 * ------------------------------------------------------------------------------------------ */
/* EntropyBottleneck.compress symbols (entropy_models.py:437-440): sym[n][c*HW + p] =
 * int32(rint(z - median[c])), idx = c; also zhat = sym + median in NHWC (dtype). */
int rgbd_eb_quantize(const float *z, int32_t z_cstride, int32_t N, int32_t HW, int32_t C,
                     const float *medians, int32_t *sym, uint8_t *idx, void *zhat,
                     int32_t zhat_dtype, int32_t zhat_cstride, int32_t zhat_coff, void *stream);
/* EntropyBottleneck.decompress dequantise (entropy_models.py:442-446, 262-263) */
int rgbd_eb_dequantize(const int32_t *sym, int32_t N, int32_t HW, int32_t C, const float *medians,
                       void *zhat, int32_t zhat_dtype, int32_t zhat_cstride, int32_t zhat_coff,
                       void *stream);
/* EntropyBottleneck.forward likelihood (entropy_models.py:369-428) at zhat = ste-rounded z,
 * likelihood written NCHW fp32; eb_params = packed per-channel parameters [C][59] (see entropy_models.py). */
int rgbd_eb_likelihood(const float *z, int32_t z_cstride, int32_t N, int32_t HW, int32_t C,
                       const float *eb_params, float lik_bound, void *zhat, int32_t zhat_dtype,
                       int32_t zhat_cstride, int32_t zhat_coff, float *lik, void *stream);

/* ------------------------------------------------------------------------------------------
 * rANS coder — byte-identical replacement for compressai.ans.
 * Tables: the int32[n_tables][stride] CDF matrix of the reference is compacted on the host
 * into one uint16 array (values mod 2^16, like the reference's uint16 casts in
 * rans_interface.cpp:131-133) plus per-table (base, length, offset).
 * ------------------------------------------------------------------------------------------ */
typedef struct rgbd_rans_tables {
    const uint16_t *cdf;    /* device, sum(length) entries */
    const int32_t *base;    /* device [n_tables] first entry of table i */
    const int32_t *length;  /* device [n_tables] = _cdf_length */
    const int32_t *offset;  /* device [n_tables] = _offset */
    int32_t n_tables;
    int32_t total;          /* sum(length) */
    /* device [total] x 16 B, encoder only: per CDF bin {u64 rcp, u32 bias, u32 range | shift << 16},
     * the exact-reciprocal form of Rans64EncSymbolInit (ryg_rans rans64.h:167-247), so that
     * mulhi(x, rcp) >> shift == x / range on the serial chain.  Built once per table set on the host. */
    const void *enc_rec;
} rgbd_rans_tables;

/* BufferedRansEncoder.encode_with_indexes + flush (rans_interface.cpp:99-192), and
 * RansEncoder.encode_with_indexes (:194-205), batched: stream s encodes symbols
 * sym[s*stream_stride .. +n_sym) with the same tables.  Each stream owns cap_words 32-bit
 * words of `out`; the finished stream is the TAIL of its region: bytes
 * out + 4*(s*cap_words + cap_words - nwords[s]) .. end.  nwords[s] < 0 flags overflow. */
int rgbd_rans_encode(const int32_t *sym, const uint8_t *idx, int64_t stream_stride, int32_t n_sym,
                     int32_t n_streams, const rgbd_rans_tables *t, uint32_t *out,
                     int64_t cap_words, int32_t *nwords, void *stream);

/* The `bytes` objects compress() returns (models/elic_united.py:423-427: encoder.flush() / RansEncoder strings): packs
 * the tails of two groups of finished streams (group a: n_a streams of cap_a words in out_a, then group b; nwords in
 * that order, as written by rgbd_rans_encode) into dst = [n_a + n_b counts | words of stream 0 | stream 1 | ...].
 * dst may be pinned host memory (zero-copy).  If the words do not fit dst_cap_words only the counts are written. */
int rgbd_gather_streams(const uint32_t *out_a, int64_t cap_a, int32_t n_a, const uint32_t *out_b, int64_t cap_b,
                        int32_t n_b, const int32_t *nwords, uint32_t *dst, int64_t dst_cap_words, void *stream);

/* RansDecoder.set_stream (rans_interface.cpp:278-284): state[s] = {x, next word}. */
typedef struct rgbd_rans_dec_state {
    uint64_t x;
    int64_t pos;
} rgbd_rans_dec_state;
int rgbd_rans_decode_init(const uint32_t *words, const int64_t *word_off, int32_t n_streams,
                          rgbd_rans_dec_state *state, void *stream);
/* RansDecoder.decode_stream / decode_with_indexes (rans_interface.cpp:207-276, 286-351):
 * continue every stream by n_sym symbols: indexes idx[s*stream_stride + chunk_off ..),
 * symbols to sym[same].  State persists in device memory between calls. */
int rgbd_rans_decode_chunk(const uint32_t *words, const int64_t *word_off, const int64_t *word_len,
                           int32_t n_streams, rgbd_rans_dec_state *state, const uint8_t *idx,
                           int32_t *sym, int64_t stream_stride, int64_t chunk_off, int32_t n_sym,
                           const rgbd_rans_tables *t, void *stream);

/* Multi-stream layout (opt-in: the reference decoder reads one y stream): the y symbols of an image are cut
 * into equal sub-streams of n_sym symbols (a whole number of channels of one checkerboard half), each of them a complete
 * RansEncoder.encode_with_indexes string of its own (rgbd_rans_encode with stream_stride = n_sym).  One coding step of
 * the decoder then decodes `per_group` whole sub-streams of each of the n_groups images concurrently — one warp each,
 * fresh state from the stream's first two words (RansDecoder.set_stream + decode_stream, rans_interface.cpp:278-351):
 *   stream (g, m), m < per_group: symbols / indexes at g * group_stride + m * stream_stride + chunk_off,
 *   its words described by word_off / word_len [g * slot_group_stride + slot_base + m]; the final decoder state goes to
 *   state[same slot] (pos == word_len there means the stream was consumed exactly). */
int rgbd_rans_decode_streams(const uint32_t *words, const int64_t *word_off, const int64_t *word_len,
                             int32_t n_groups, int32_t per_group, int32_t slot_base, int32_t slot_group_stride,
                             rgbd_rans_dec_state *state, const uint8_t *idx, int32_t *sym, int64_t group_stride,
                             int64_t stream_stride, int64_t chunk_off, int32_t n_sym, const rgbd_rans_tables *t,
                             void *stream);

/* pmf_to_quantized_cdf (ops.cpp:24-81). Host function. cdf has n+1 entries. */
int rgbd_pmf_to_quantized_cdf(const float *pmf, int32_t n, int32_t precision, uint32_t *cdf);

#ifdef __cplusplus
}
#endif
#endif /* RGBD_B200_H */
