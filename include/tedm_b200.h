/* tedm_b200 -- C ABI of the B200-native TEDM hot path (libtedm_b200.so).
 *
 * The reference (mmr12/TEDM) is pure Python: its boundary for this path is the nn.Module surface
 * of models/unet_model.py, models/diffusion_model.py and models/datasetDM_model.py, and every
 * arithmetic step is a PyTorch library call.  This header is what replaces those library calls:
 * one entry point per fused op.  Each comment cites the reference lines the entry point replaces.
 *
 * Conventions
 *   - plain pointers + sizes; no torch types.  All pointers are DEVICE pointers unless stated.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t) and never synchronises, so call
 *     sequences can be captured in a CUDA graph.
 *   - activations: NHWC bf16.  conv weights: [Cout][kh][kw][Cin] bf16 ("KRSC").  Statistics, loss,
 *     schedule tables, time embeddings: fp32.  Timesteps: int64.  Public image tensors (x_0, x_t,
 *     noise, UNet output, logits): NCHW fp32 exactly as the reference.
 *   - return 0 on success, <0 on error (TEDM_ERR_*); tedm_last_error() gives the message of the
 *     last failing call on this thread.  Unsupported shapes are hard errors: there is no fallback.
 */
#ifndef TEDM_B200_H
#define TEDM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEDM_ABI_VERSION 1
#define TEDM_ERR_ARG (-1)
#define TEDM_ERR_CUDA (-2)
#define TEDM_ERR_UNSUPPORTED (-3)

typedef void* tedm_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define TEDM_API __attribute__((visibility("default")))
#else
#define TEDM_API
#endif

TEDM_API int tedm_version(void);
TEDM_API const char* tedm_last_error(void);

/* ---- DDPM arithmetic ------------------------------------------------------------------- */

/* x_t = sa[t_b] * x0' + sb[t_b] * noise, x0' = normalize ? x0*2-1 : x0.  fp32, bit-exact with the
 * reference's unfused mul/mul/add.  Replaces DiffusionModel.forward_diffusion_model
 * (models/diffusion_model.py:176-203), get_index_from_list (trainers/utils.py:48-59) and
 * normalize_to_neg_one_to_one (trainers/utils.py:28-29, models/diffusion_model.py:169-170). */
TEDM_API int tedm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac,
                  const float* sqrt_1m_ac, float* x_t, int batch, int chw, int T, int normalize,
                  tedm_stream_t stream);

/* per-image mean |pred-target| * w[t_b] -> loss_per_image[b]; mean over b -> loss[0].
 * grad (nullable) = d loss / d pred = sign(pred-target) * w[t_b] / (chw * batch).
 * Replaces F.l1_loss + reduce + p2 weighting (models/diffusion_model.py:138-143). */
TEDM_API int tedm_l1_loss(const float* pred, const float* target, const int64_t* t, const float* p2_weight,
                 float* loss_per_image, float* loss, float* grad, int batch, int chw, int T,
                 tedm_stream_t stream);

/* One reverse-diffusion update after the UNet call: x0_hat = sqrt_recip_ac*x_t - sqrt_recipm1_ac*eps;
 * dynamic thresholding by the EXACT per-image quantile of |x0_hat| (radix select of the order
 * statistics k_lo and k_lo+1, then torch.quantile's lerp with weight q_weight; the host derives
 * k_lo/q_weight from the percentile exactly as torch does: rank = float32(q) * (chw-1));
 * s = max(s,1); x0_hat = clip(x0_hat,-s,s)/s; x_prev = coef1*x0_hat + coef2*x_t (+ sigma*z when z is
 * not NULL; pass NULL at t==0).  All images share the timestep, so the schedule values are scalars.
 * x0_hat and s_out (per-image threshold) are optional outputs.
 * Replaces sample_timestep / p_mean_variance / predict_x_0_from_noise / q_posterior
 * (models/diffusion_model.py:205-286). */
TEDM_API int tedm_sampler_step(const float* x_t, const float* eps, const float* z, float* x_prev,
                      float* x0_hat /*nullable*/, float* s_out /*nullable*/, float sqrt_recip_ac,
                      float sqrt_recipm1_ac, float coef1, float coef2, float sigma, int k_lo,
                      float q_weight, int batch, int chw, tedm_stream_t stream);

/* The same update with the five per-step schedule values (sqrt_recip_ac, sqrt_recipm1_ac, coef1, coef2, sigma) read from
 * DEVICE memory `coefs` and z always given (sigma = 0 at t == 0): one captured launch serves all 1000 reverse steps of
 * sample_plot_image (trainers/utils.py:62-98) when the step is replayed from a CUDA graph. */
TEDM_API int tedm_sampler_step_dev(const float* x_t, const float* eps, const float* z, float* x_prev,
                          float* x0_hat /*nullable*/, float* s_out /*nullable*/, const float* coefs, int k_lo,
                          float q_weight, int batch, int chw, tedm_stream_t stream);

/* ---- UNet pieces ------------------------------------------------------------------------- */

/* SinusoidalPosEmb(dim) -> Linear(dim,tdim) -> GELU(erf) -> Linear(tdim,tdim); fp32.
 * freq: [dim/2] fp32, the reference's exp(arange(dim/2) * -log(1e4)/(dim/2-1)) table.
 * Replaces models/unet_model.py:76-93 and Unet.time_mlp (:287-292). */
TEDM_API int tedm_time_embed(const int64_t* t, const float* freq, const float* w1, const float* b1,
                    const float* w2, const float* b2, float* temb, int batch, int dim, int tdim,
                    tedm_stream_t stream);

/* out[b][j] = sum_k W[j][k] * silu(temb[b][k]) + bias[j] for the concatenation of every
 * ResnetBlock.time_mlp of the net (models/unet_model.py:150-152,168-171). */
TEDM_API int tedm_time_proj(const float* temb, const float* w_cat, const float* b_cat, float* out, int batch,
                   int tdim, int total, tedm_stream_t stream);

/* 7x7 pad-3 stem conv on the fp32 NCHW network input -> NHWC bf16 (Unet.init_conv,
 * models/unet_model.py:267,334).  weight fp32 [Cout][Cin][7][7]. */
TEDM_API int tedm_stem_conv7x7(const float* x, const float* weight, const float* bias, void* out, int batch,
                      int cin, int height, int width, int cout, tedm_stream_t stream);

/* Implicit-GEMM convolution on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *   mode 0: 1x1            (res_conv, to_qkv, to_out, head layer 1: unet_model.py:157,185,188,226,227)
 *   mode 1: 3x3 pad 1      (Block.proj, last-level down/up conv: :122,307-308,323-324)
 *   mode 2: 4x4 stride 2 pad 1 (Downsample: :47-49)
 *   mode 3: nearest-x2 upsample folded into 3x3 pad 1 (Upsample: :39-44); weight is the
 *           parity-combined [4][Cout][2][2][Cin] tensor made by tedm_fold_upsample_weight.
 * K runs over src0's channels then src1's (the skip concat of Unet.forward, :356,359,365, is
 * never materialised), then over the optional extra sources.  Epilogue: + bias, + residual, bf16 store, optional per-(image, group)
 * partial sum / sum-of-squares of the fp32 accumulators for the GroupNorm that follows. */
typedef struct {
  const void* src0;      /* [B][H][W][c0] bf16 */
  const void* src1;      /* [B][H][W][c1] bf16 or NULL */
  const void* weight;    /* bf16 KRSC, K = taps*(c0+c1 [+ extra channels]) [+ centre-only extra channels] */
  const float* bias;     /* [cout] or NULL */
  const void* residual;  /* [B][Ho][Wo][cout] bf16 or NULL */
  void* out;             /* [B][Ho][Wo][cout] bf16 */
  float* gn_partial;     /* NULL or [B][gn_parts][gn_groups][2] fp32 (sum, sum of squares) */
  int batch, height, width; /* input extent */
  int c0, c1, cout;
  int mode;
  int gn_groups;         /* number of GroupNorm groups (0 when gn_partial is NULL) */
  int out_dtype;         /* 0: out is bf16 (TMA store); 1: out is fp32 (head layer 1 keeps full precision) */
  int64_t src0_image_stride, src1_image_stride, out_image_stride; /* elements; 0 = dense */
  /* split output (0 = off): output channels [0, split) go to out (+ residual), channels [split, cout) to out2
   * (+ residual2, nullable) as a dense [B][Ho][Wo][cout - split] tensor.  This is the data gradient of a
   * skip-concat convolution landing in its two sources (backward of models/unet_model.py:356,359,365). */
  int split;
  void* out2;
  const void* residual2;
  /* further A sources, walked after src0 / src1 inside every tap (K order: tap, then source, then channel): the same
   * [B][H][W][c] bf16 layout.  Uses: (1) a 1x1 branch folded into a 3x3 conv as extra K at the centre tap
   * (extra_center = 1: ResnetBlock's res_conv inside block2's conv, models/unet_model.py:157,175); (2) the fp32
   * precision mode, where every operand is a bf16 (hi, lo) pair and the product is hi*hi + lo*hi + hi*lo, i.e. a conv
   * over the sources [x_hi, x_lo, x_hi] against the weights [w_hi, w_hi, w_lo]. */
  int n_extra;                      /* 0..4 */
  const void* extra_src[4];
  int extra_c[4];
  int64_t extra_image_stride[4];    /* elements; 0 = dense */
  int extra_center[4];
  /* NULL, or fp32 [B][cout][2] from tedm_gn_affine: the residual then enters as SiLU(a * residual + b), i.e. `residual` is
   * the raw output of block2's conv and this call is ResnetBlock's `block2(h) + res_conv(x)` (models/unet_model.py:174-175)
   * with GroupNorm, SiLU and the add done in the 1x1 conv's epilogue (needs tiles inside one image: Ho * Wo >= 128). */
  const float* residual_affine;
  /* NULL, or fp32 [B][c0][2] from tedm_gn_affine: the conv's input is SiLU(a * src0 + b), i.e. src0 is the raw output of
   * block1's conv and this call is Block.forward's GroupNorm + scale/shift + SiLU (models/unet_model.py:128-134) fused into
   * block2's conv: the halo boxes are normalised in shared memory, the activation never reaches HBM.  Single-source 3x3 convs
   * on the halo-tile path only (tedm_conv_src_affine_supported). */
  const float* src0_affine;
} tedm_conv_args;
TEDM_API int tedm_conv_igemm_fwd(const tedm_conv_args* args, tedm_stream_t stream);
/* Weight gradient of the same convolution (backward of models/unet_model.py:43,49,122,157,185,188,226,227,308,324):
 * dw[co][tap][ci] = sum over pixels of dy[pixel][co] * src[pixel + tap offset][ci], fp32, overwritten.
 * `args` describes the forward call (src0/src1/extent/mode/cout; weight, bias, residual, out, gn_* are ignored;
 * out_image_stride, if non-zero, is dy's image stride); dy is the NHWC bf16 output gradient.  taps = 1 (mode 0),
 * 9 (mode 1), 16 (mode 2: ky*4+kx; mode 3: parity*4 + a*2 + b of the folded 2x2 kernels).
 * With oihw_accumulate != 0, dw is instead the fp32 OIHW parameter gradient [cout][c0+c1][kh][kw] and the result is
 * ACCUMULATED into it (mode 3: the folded taps are scattered back onto the 3x3 kernel of Upsample's conv). */
TEDM_API int tedm_conv_igemm_wgrad(const tedm_conv_args* args, const void* dy, float* dw, int oihw_accumulate,
                          float* workspace, tedm_stream_t stream);
/* fp32 elements of the REQUIRED `workspace` above.  The 3x3 halo-tile kernel stores its split-K partial tiles there and a
 * second kernel adds them in slice order (always bit-reproducible).  The generic kernel (1x1, 4x4-s2, folded upsample, the
 * widest 3x3) adds its slices straight into dw with fp32 reductions: in arrival order by default, or -- after
 * tedm_conv_set_deterministic(1) -- slice after slice, the order enforced by per-tile turn counters kept at the end of
 * the workspace (split-K capped at 4; measured cost in DESIGN.md).  The workspace must be ZERO when first used (the
 * counters reset themselves); calls that share one workspace must be ordered on one stream. */
TEDM_API int64_t tedm_conv_igemm_wgrad_workspace(void);
/* 1: every convolution weight gradient is bit-reproducible from run to run (torch.use_deterministic_algorithms-style
 * opt-in); 0 (default): the generic kernel's split-K slices add in arrival order. */
TEDM_API int tedm_conv_set_deterministic(int enable);
/* 1 (default): 3x3 / 4x4 / folded-upsample convolutions over >= 128 input channels whose N tile is <= 128 run as CTA pairs
 * (tcgen05 cta_group::2, clusters of two SMs: M = 256, each CTA stages its own pixels and half of the weight tile);
 * 0: one CTA per tile everywhere; 2: pairs wherever the geometry allows (tests, A/B runs). */
TEDM_API int tedm_conv_set_cta_pairs(int enable);
/* 1 if a 3x3 conv of this shape accepts tedm_conv_args.src0_affine */
TEDM_API int tedm_conv_src_affine_supported(int height, int width, int c0, int cout);
/* number of partial-statistics slots per image that tedm_conv_igemm_fwd writes for this output extent */
TEDM_API int tedm_conv_gn_parts(int out_height, int out_width);
/* tuning/debug: force the N tile (64/128/256; 0 = automatic) of tedm_conv_igemm_fwd */
TEDM_API int tedm_conv_set_tile_n(int bn);
/* tuning/debug: halo-tile kernel of the 3x3 weight gradient: 0 off, 1 automatic (default), 2 wherever the geometry allows */
TEDM_API int tedm_conv_set_wgrad_halo(int enable);
/* tuning/debug: enable (default) / disable the weight-stationary row path of the 3x3 conv */
TEDM_API int tedm_conv_set_ws(int enable);
/* 3x3 convolutions on images of >= 16 rows run on 8 x 16-pixel tiles whose 10 x 18 halo box serves all nine taps (1, default);
 * 0 = one activation box per tap everywhere; 2 = 1 with the residual tiles of the 1x1 convs fetched through registers instead
 * of TMA; 3 = 1 plus halo tiles for the folded upsample conv (measured slower) -- A/B runs and tests. */
TEDM_API int tedm_conv_set_halo(int enable);

/* fp32 OIHW [Cout][Cin][kh][kw] -> bf16 KRSC [Cout][kh][kw][Cin] (derived weight cache). */
TEDM_API int tedm_weight_to_krsc(const float* w_oihw, void* w_krsc, int cout, int cin, int kh, int kw,
                        tedm_stream_t stream);
/* fp32 OIHW 3x3 -> bf16 [4 parities][Cout][2][2][Cin] for mode 3. */
TEDM_API int tedm_fold_upsample_weight(const float* w_oihw, void* w_folded, int cout, int cin, tedm_stream_t stream);
/* All conv weights of a net in ONE launch (a training step changes every weight): table_dev is a DEVICE array of
 * entries; entry i covers CTAs [cta_begin, cta_begin + (cout/32)*(cin/32)); cout, cin multiples of 32.
 * fwd: KRSC (modes 0-2) or the folded layout (mode 3); dgrad: the tedm_weight_to_dgrad layout; either may be NULL. */
typedef struct {
  const void* w;   /* fp32 OIHW parameter */
  void* fwd;       /* bf16 forward operand or NULL */
  void* dgrad;     /* bf16 data-gradient operand or NULL */
  int cout, cin, mode, cta_begin;
} tedm_weight_entry;
TEDM_API int tedm_prepare_weights(const tedm_weight_entry* table_dev, int n_entries, int total_ctas, tedm_stream_t stream);

/* GroupNorm finalise + affine + optional (scale+1)/shift + SiLU (+ residual), one pass.
 * Replaces Block.forward after the conv (models/unet_model.py:126-135) and the residual add of
 * ResnetBlock.forward (:175).  scale_shift: fp32 [B][ss_stride], scale at ss_offset, shift at
 * ss_offset + C (the chunk(2) of :171), or NULL. */
TEDM_API int tedm_gn_silu_fwd(const void* x, const float* gn_partial, int gn_parts, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int ss_offset,
                     const void* residual, void* out, int batch, int hw, int channels, int groups,
                     float eps, tedm_stream_t stream);

/* The per-(image, channel) affine of tedm_gn_silu_fwd on its own: affine[b][c] = (a / 2, b / 2) with
 * GroupNorm(x)[c] * (scale + 1) + shift = a * x + b (models/unet_model.py:128-133), for consumers that apply the
 * normalisation themselves (tedm_conv_args.residual_affine).  fp32 [B][C][2]. */
TEDM_API int tedm_gn_affine(const float* gn_partial, int gn_parts, const float* gamma, const float* beta,
                   const float* scale_shift, int ss_stride, int ss_offset, float* affine, int batch, int hw,
                   int channels, int groups, float eps, tedm_stream_t stream);

/* Per-pixel channel LayerNorm with gain only (+ residual): models/unet_model.py:52-61, and the
 * Residual wrapper (:29-36) when `residual` is given. */
TEDM_API int tedm_layernorm_fwd(const void* x, const float* g, const void* residual, void* out, int64_t npix,
                       int channels, float eps, tedm_stream_t stream);

/* LinearAttention core between to_qkv and to_out (models/unet_model.py:197-209):
 * qkv [B][n][3*heads*dh] bf16 -> out [B][n][heads*dh] bf16.  workspace: fp32,
 * tedm_linear_attention_workspace(batch, n, heads, dh) elements. */
TEDM_API int64_t tedm_linear_attention_workspace(int batch, int n, int heads, int dim_head);
TEDM_API int tedm_linear_attention_fwd(const void* qkv, void* out, float* workspace, int batch, int n, int heads,
                              int dim_head, float scale, tedm_stream_t stream);

/* Attention core of the mid block (models/unet_model.py:229-240): q,k L2-normalised along n,
 * sim*scale, softmax, @v.  n <= 256. */
TEDM_API int tedm_attention_fwd(const void* qkv, void* out, int batch, int n, int heads, int dim_head, float scale,
                       tedm_stream_t stream);

/* nearest x2 (nn.Upsample, models/unet_model.py:42) on NHWC bf16. */
TEDM_API int tedm_upsample2x(const void* x, void* out, int batch, int height, int width, int channels,
                    tedm_stream_t stream);

/* final 1x1 conv C -> out_dim, NHWC bf16 -> NCHW fp32 (Unet.final_conv, models/unet_model.py:331,368). */
TEDM_API int tedm_final_conv1x1(const void* x, const float* weight, const float* bias, float* out, int batch, int hw,
                       int channels, int out_dim, tedm_stream_t stream);

/* layout conversion at the module boundary */
TEDM_API int tedm_nchw_f32_to_nhwc_bf16(const float* x, void* out, int batch, int channels, int hw, tedm_stream_t stream);
TEDM_API int tedm_nhwc_bf16_to_nchw_f32(const void* x, float* out, int batch, int channels, int hw, tedm_stream_t stream);

/* ---- training step: backward of the UNet pieces -------------------------------------------
 * The reference gets these from torch autograd over models/unet_model.py (loss.backward() in
 * trainers/train_CXR14.py:30-40).  Activation gradients are NHWC bf16; every PARAMETER gradient is
 * fp32 and is accumulated (+=) into its destination, like torch's .grad.
 * The data gradient of a convolution runs on tedm_conv_igemm_fwd itself with the weights re-laid
 * out by tedm_weight_to_dgrad (a 3x3 becomes the flipped/transposed 3x3, the stride-2 4x4 becomes a
 * mode-3 parity conv, the folded upsample conv becomes a mode-2 stride-2 conv). */

/* fp32 OIHW parameter -> bf16 operand of the data-gradient conv; `mode` is the FORWARD mode.
 *   0/1: [Cin][kh][kw][Cout] flipped (run as mode 0/1);  2: [4][Cin][2][2][Cout] (run as mode 3);
 *   3 (w is the 3x3 of Upsample): [Cin][4][4][Cout] (run as mode 2 on the 2x-resolution gradient). */
TEDM_API int tedm_weight_to_dgrad(const float* w_oihw, void* w_dgrad, int cout, int cin, int mode, tedm_stream_t stream);
/* fp32 [Cout][taps][Cin] from tedm_conv_igemm_wgrad -> grad_oihw += (mode 3 un-folds the parity kernels to 3x3). */
TEDM_API int tedm_wgrad_to_oihw(const float* dw, float* grad_oihw, int cout, int cin, int mode, tedm_stream_t stream);

/* Backward of tedm_gn_silu_fwd (Block.forward, models/unet_model.py:126-135): dx = d loss / d (conv output).
 * dgamma/dbeta [C] +=; dbias [C] += sum of dx over batch and pixels (the conv's bias gradient; nullable);
 * dscale_shift (nullable): same layout as scale_shift, the (b, c) entries of this block are overwritten.
 * workspace: fp32 [3*batch*channels].  The gradient of the fused residual input is dy itself. */
TEDM_API int tedm_gn_silu_bwd(const void* x, const void* dy, const float* gn_partial, int gn_parts, const float* gamma,
                     const float* beta, const float* scale_shift, int ss_stride, int ss_offset, void* dx,
                     float* workspace, float* dgamma, float* dbeta, float* dbias, float* dscale_shift,
                     int batch, int hw, int channels, int groups, float eps, tedm_stream_t stream);

/* Backward of tedm_layernorm_fwd (models/unet_model.py:52-61): dx = LN'(x)·dy (+ add, nullable: the gradient
 * arriving over the Residual branch); dg [C] +=. */
TEDM_API int tedm_layernorm_bwd(const void* x, const float* g, const void* dy, const void* add, void* dx, float* dg,
                       int64_t npix, int channels, float eps, tedm_stream_t stream);

/* dbias[c] += sum over pixels of dy[pixel][c] (bias gradient of a conv not followed by GroupNorm). */
TEDM_API int tedm_bias_grad(const void* dy, float* dbias, int64_t npix, int channels, tedm_stream_t stream);
/* out = a + b, bf16: the two gradient streams meeting at a skip connection (models/unet_model.py:339,344). */
TEDM_API int tedm_add_bf16(const void* a, const void* b, void* out, int64_t n, tedm_stream_t stream);

/* Backward of tedm_final_conv1x1 (models/unet_model.py:331,368): dh NHWC bf16; dweight [out_dim][C], dbias += . */
TEDM_API int tedm_final_conv1x1_bwd(const void* h, const float* weight, const float* dout, void* dh, float* dweight,
                           float* dbias, int batch, int hw, int channels, int out_dim, tedm_stream_t stream);
/* Weight/bias gradient of the 7x7 stem (models/unet_model.py:267,334): x fp32 NCHW, dy NHWC bf16. */
TEDM_API int tedm_stem_conv7x7_wgrad(const float* x, const void* dy, float* dweight, float* dbias, int batch, int cin,
                            int height, int width, int cout, tedm_stream_t stream);

/* tedm_time_embed that also returns what its backward needs: emb [B][dim], hidden_pre [B][tdim] (pre-GELU). */
TEDM_API int tedm_time_embed_train(const int64_t* t, const float* freq, const float* w1, const float* b1, const float* w2,
                          const float* b2, float* emb, float* hidden_pre, float* temb, int batch, int dim, int tdim,
                          tedm_stream_t stream);
/* Backward of a small fp32 Linear of the time path (models/unet_model.py:150-152,287-292):
 *   dY = dy_raw * act'(y_pre) ; dw [n_out][n_in] += dY^T act(x) ; db += sum_b dY ; dx_raw = dY w (overwritten, nullable)
 * act codes: 0 identity, 1 SiLU, 2 GELU(erf).  y_pre may be NULL when act_y == 0. */
TEDM_API int tedm_linear_bwd(const float* dy_raw, const float* y_pre, int act_y, const float* x, int act_x, const float* w,
                    float* dw, float* db, float* dx_raw, int batch, int n_out, int n_in, tedm_stream_t stream);

/* Backward of tedm_linear_attention_fwd; fwd_workspace is the forward call's workspace (column maxima and
 * per-chunk context partials).  workspace: fp32, tedm_linear_attention_bwd_workspace() elements. */
TEDM_API int64_t tedm_linear_attention_bwd_workspace(int batch, int n, int heads, int dim_head);
TEDM_API int tedm_linear_attention_bwd(const void* qkv, const void* dout, const float* fwd_workspace, void* dqkv,
                              float* workspace, int batch, int n, int heads, int dim_head, float scale,
                              tedm_stream_t stream);
/* Backward of tedm_attention_fwd. */
TEDM_API int tedm_attention_bwd(const void* qkv, const void* dout, void* dqkv, int batch, int n, int heads, int dim_head,
                       float scale, tedm_stream_t stream);

/* The same backward for any token count, flash style on tensor cores (4 launches; nothing n x n leaves the SM).  o is the
 * forward output [B][n][heads*dim_head] bf16 (delta_i = <dO_i, O_i>); workspace: tedm_attention_bwd_flash_workspace floats. */
TEDM_API int64_t tedm_attention_bwd_flash_workspace(int batch, int n, int heads);
TEDM_API int tedm_attention_bwd_flash(const void* qkv, const void* o, const void* dout, void* dqkv, float* workspace, int batch,
                             int n, int heads, int dim_head, float scale, tedm_stream_t stream);

/* torch.optim.Adam update (trainers/train_CXR14.py:139) over a flat fp32 arena; n % 4 == 0; grad is multiplied by
 * grad_scale first.  The 1-based step count for the bias correction is `step`, or, when step_counter (a DEVICE int)
 * is given, the counter's value after this call has incremented it -- so that a step replayed from a CUDA graph
 * still advances. */
TEDM_API int tedm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, int* step_counter,
                   float grad_scale, tedm_stream_t stream);

/* ---- TEDM / LEDM head ---------------------------------------------------------------------- */

/* Per-pixel MLP tail after the per-level layer-1 GEMMs (commuted form of the reference's
 * upsample+concat+conv1x1): z1 = b1 + sum_{s<n_sum} sum_l g_l[img*n_sum+s][y>>sh_l][x>>sh_l];
 * ReLU; BN1 (folded affine a1,c1); W2,b2; ReLU; BN2 (a2,c2); w3,b3 -> logits fp32 [n_img][H][W].
 * Replaces F.interpolate + cat + classifier (models/datasetDM_model.py:57-64,80-88;
 * trainers/train_datasetDM.py:30-42), eval-mode BatchNorm. */
typedef struct {
  const void* g[4];      /* level l: [n_img*n_sum][H>>shift[l]][W>>shift[l]][c1], bf16 or fp32 (g_dtype) */
  int g_dtype;           /* 0 = bf16, 1 = fp32 */
  int shift[4];
  int n_levels;
  int n_sum;             /* 1 for TEDM (shared head), S for LEDM/LEDMe */
  int n_img, height, width;
  int c1, c2;            /* 128, 32 */
  const float* b1; const float* bn1_a; const float* bn1_c;   /* [c1] */
  const float* w2;       /* [c2][c1] fp32 */
  const float* b2; const float* bn2_a; const float* bn2_c;   /* [c2] */
  const float* w3;       /* [c2] */
  float b3;
  float* logits;
  /* optional (tedm_head_infer only): the full-resolution level given as its FEATURE map instead of a g map --
   * f_full bf16 [n_img][H][W][c_full] and that level's layer-1 weight slice w1_full bf16 [c1][c_full]; its layer 1 then
   * runs inside the tail kernel.  Needs fp32 g maps for the other levels, c_full == 64, n_sum == 1. */
  const void* f_full;
  const void* w1_full;
  int c_full;
  int exact;             /* 1 (fp32 precision mode): fp32 g maps through the plain-fp32 tail kernel, no tensor-core tail */
} tedm_head_args;
TEDM_API int tedm_head_infer(const tedm_head_args* args, tedm_stream_t stream);

/* ---- head training (BatchNorm batch statistics; gradients into the head parameters only) --------------------
 * Replaces model(x) in training mode + loss.backward() of trainers/train_datasetDM.py:30-42,88-99 for the
 * classifier of models/datasetDM_model.py:57-64.  The GEMMs (layer 1 per level, layer 2 on relu(z1) with BatchNorm-1
 * folded into its weights, d h1 = W2^T d z2, every weight gradient) are tedm_conv_igemm_fwd / _wgrad calls; these entry
 * points are the memory-bound passes between them.  Per-pixel tensors: a1 bf16 [N][H][W][128]; z2 fp32 [N][H][W][64]
 * (channels 32..63 padding); dz2 bf16 [N][H][W][64].  Reduction buffers are fp32 and must be zeroed by the caller. */
/* a1 = relu(b1 + sum_s sum_l g_l[...]) as bf16; sums[0][c] += sum a1, sums[1][c] += sum a1^2 (args: g, shift, n_levels,
 * n_sum, n_img, height, width, c1, b1 are read). */
TEDM_API int tedm_head_train_z1(const tedm_head_args* args, void* a1, float* sums, tedm_stream_t stream);
/* BatchNorm2d training-mode statistics from (sum, sum of squares) over `count` samples: stats[0]=mean, [1]=rstd,
 * [2]=A=gamma*rstd, [3]=C=beta-mean*A ([4][channels]); running_mean/var (nullable pair) updated like torch. */
TEDM_API int tedm_bn_finalize(const float* sums, double count, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* stats, int channels,
                     tedm_stream_t stream);
/* w2_folded bf16 [64][128] = W2 diag(A1) (rows >= 32 zero); b2_folded fp32 [64] = b2 + W2 C1; w2_t bf16 [128][64] = W2^T. */
TEDM_API int tedm_head_fold_w2(const float* w2, const float* b2, const float* stats1, void* w2_folded, float* b2_folded,
                      void* w2_t, tedm_stream_t stream);
/* sums[0][j] += sum relu(z2_j), sums[1][j] += sum relu(z2_j)^2, j < 32. */
TEDM_API int tedm_head_z2_stats(const float* z2, float* sums, int64_t npix, tedm_stream_t stream);
/* mode 0: logits = w3 . BN2(relu(z2)) + b3.   mode 1: S[0][j] += sum dh2_j, S[1][j] += sum dh2_j*a2hat_j,
 * S[2][j] += sum dlogit*h2_j, S[3][0] += sum dlogit.   mode 2: dz2 (bf16) and S[4][j] += sum dz2_j.   S is fp32 [5][32]. */
TEDM_API int tedm_head_train_tail(int mode, const float* z2, const float* stats2, const float* w3, const float* b3,
                         const float* dlogit, float* logits, float* S, void* dz2, double count, int64_t npix,
                         tedm_stream_t stream);
/* dh1: fp32 [N][H][W][128].  mode 0: T[0][k] += sum dh1_k, T[1][k] += sum dh1_k*a1hat_k.   mode 1: dz1 through BatchNorm-1 and ReLU-1; db1 += sum dz1;
 * pooled[l] (bf16 [N][H>>s][W>>s][128], HOST array of device pointers; shifts: HOST array, 0..3) = 2^s x 2^s block sums
 * of dz1, the output gradient of layer 1 at level l's native resolution. */
TEDM_API int tedm_head_bn1_bwd(int mode, const void* dh1, const void* a1, const float* stats1, float* T, float* db1,
                      void* const* pooled, const int* shifts, int n_levels, int n_img, int height, int width,
                      double count, tedm_stream_t stream);
/* dW2 += dW2_folded diag(A1) + db2 C1^T; db2, dgamma/dbeta of both norms, dw3, db3 += their reduction buffers. */
TEDM_API int tedm_head_param_grads(const float* dw2_folded, const float* stats1, const float* S, const float* T, float* dw2,
                          float* db2, float* dgamma1, float* dbeta1, float* dgamma2, float* dbeta2, float* dw3,
                          float* db3, tedm_stream_t stream);

/* prob[b] = mean_s sigmoid(logits[b*S+s]); mask = prob > 0.5
 * (auxiliary/postprocessing/testing_shared_weights.py:113,120,133-138; app.py:79). */
TEDM_API int tedm_ensemble_mask(const float* logits, float* prob, uint8_t* mask, int batch, int n_steps, int hw,
                       tedm_stream_t stream);

/* Whole Residual(PreNorm(LinearAttention)) block, inference forward, in three launches (models/unet_model.py:29-36,
 * 64-73,178-210): out = LayerNorm(W_out . linattn(W_qkv . LayerNorm(x; g_pre)) + b_out; g_out) + x.
 * x, out: NHWC bf16 [B][n][channels]; wqkv bf16 [3*heads*dim_head][channels]; wout bf16 [channels][heads*dim_head].
 * q, k, v and the attention output stay on chip (softmax over n is computed online).  Supported: 4 heads x 32,
 * channels 64 or 128, n a multiple of 64 (tedm_linear_attention_fused_supported); anything else is TEDM_ERR_UNSUPPORTED
 * and the caller runs the unfused sequence.  workspace: tedm_linear_attention_fused_workspace(batch, n) floats. */
TEDM_API int tedm_linear_attention_fused_supported(int n, int channels, int heads, int dim_head);
TEDM_API int64_t tedm_linear_attention_fused_workspace(int batch, int n);
TEDM_API int tedm_linear_attention_fused_fwd(const void* x, const void* wqkv, const float* g_pre, const void* wout,
                                    const float* b_out, const float* g_out, void* out, float* workspace, int batch,
                                    int n, int channels, int heads, int dim_head, float scale, float eps,
                                    tedm_stream_t stream);

/* ---- supervised segmentation: loss, metrics, input transport -------------------------------- */

/* Rows are the (b, c) planes of an NCHW fp32 logit tensor (n_rows = B*C, row_len = H*W); row r is compared with target
 * row r / target_repeat, i.e. repeat(y, 'b c h w -> (b step) c h w') (trainers/train_baseline.py:30-31) is never built.
 * row_mean[r] (nullable) = mean_i bce(logits[r][i], target[.][i]); loss[0] = mean_r row_mean[r]: replaces
 * reduce(binary_cross_entropy_with_logits(pred, y, reduction='none'), 'b c h w -> b c', 'mean').mean()
 * (trainers/train_baseline.py:44-45).  grad (nullable) = grad_scale * (sigmoid(x) - y) / (row_len * n_rows).
 * workspace: tedm_bce_workspace_floats(n_rows) floats.  Summation order is fixed (deterministic). */
TEDM_API int tedm_bce_logits(const float* logits, const float* target, float* row_mean, float* loss, float* grad,
                    float* workspace, long long n_rows, long long row_len, int target_repeat, float grad_scale,
                    tedm_stream_t stream);
TEDM_API int tedm_bce_workspace_floats(long long n_rows);

/* dice / precision / recall per (b, c) row (trainers/train_baseline.py:146-161).  pred: uint8 mask (nonzero = True) or,
 * with pred_is_logits, fp32 logits thresholded as sigmoid(x) > .5 (:122).  out fp32 [n_rows][8] =
 * {dice = 2TP/(sum pred + sum target), precision = TP/(TP+FP), recall = TP/(TP+FN), TP, FP, FN, sum pred, sum target};
 * an empty denominator gives NaN like the reference (its callers nanmean). */
TEDM_API int tedm_seg_metrics(const void* pred, int pred_is_logits, const float* target, float* out, long long n_rows,
                     long long row_len, int target_repeat, tedm_stream_t stream);

/* dst = src / 255 (fp32 true division: bit-exact torchvision ToTensor, dataloaders/CXR14.py:67-70, JSRT.py:62-65);
 * images cross PCIe as uint8, a quarter of the reference's fp32 bytes. */
TEDM_API int tedm_u8_to_unit(const uint8_t* src, float* dst, long long n, tedm_stream_t stream);
/* src uint8 [n_img][n_masks][hw] -> dst fp32 [n_img][hw] = min(sum_k (src_k / 255 > .5), 1)  (dataloaders/JSRT.py:67-82). */
TEDM_API int tedm_u8_masks_to_label(const uint8_t* src, float* dst, long long n_img, long long hw, int n_masks,
                           tedm_stream_t stream);

/* ---- test-only ----------------------------------------------------------------------------- */

/* Hardware probe used by tests/test_umma_probe.py: runs 128x64x64 UMMAs whose A descriptor start
 * is shifted by shifts[v] 128-byte rows (descriptor base_offset field = base_offsets[v]) over a
 * TMA-written 384-row buffer.  shifts/base_offsets are HOST arrays; A [384][64] bf16, Bm [64][64]
 * bf16 and out [nvar][128][64] fp32 are device pointers. */
TEDM_API int tedm_debug_umma_probe(const void* A, const void* Bm, const int* shifts, const int* base_offsets,
                                   int nvar, float* out, tedm_stream_t stream);

/* The same block with every GEMM on tcgen05 (csrc/attention_tc.cu): LayerNorm -> k / q projections as 128-pixel UMMA tiles
 * -> softmaxes by the thread that owns the pixel -> context / to_out as UMMAs again; v is never materialised and the to_out
 * conv is folded into a per-image matrix.
 *   wqkv_g     = to_qkv weight with the pre-norm gain folded in: bf16 [384][C], row r = W[r] * g_pre (the kernel's
 *                LayerNorm then has no gain); the 128 q rows additionally times log2(e) (the softmax over d is an exp2)
 *   shift_log2 = log2(e) * B, B >= every |q| and |k| logit; a WEIGHT-ONLY bound does (max_r ||wqkv_g[r]||_2 * sqrt(C), by
 *                Cauchy-Schwarz, since ||LayerNorm(x)||_2 <= sqrt(C)).  It replaces the running maximum of the softmax over
 *                pixels (a per-column constant cancels) and lets the softmax over d skip its maximum; the caller routes
 *                blocks with B > ~40 (exp(-2B) must stay a normal fp32 number) to tedm_linear_attention_fused_fwd.
 * n % 512 == 0, C = 64 / 128, 4 heads x 32.  workspace: tedm_linear_attention_tc_workspace() fp32 elements; it starts with
 * the folded per-image matrices M [batch][C][128] bf16. */
TEDM_API int tedm_linear_attention_tc_supported(int n, int channels, int heads, int dim_head);
TEDM_API int64_t tedm_linear_attention_tc_workspace(int batch, int n, int channels);   /* fp32 elements */
TEDM_API int tedm_linear_attention_tc_fwd(const void* x, const void* wqkv_g, float shift_log2, const void* wout, const float* b_out,
                                 const float* g_out, void* out, float* workspace, int batch, int n, int channels, int heads,
                                 int dim_head, float scale, float eps, tedm_stream_t stream);

/* ---- fp32 precision mode (north star: "1e-4 in fp32 mode"; the reference's default arithmetic, config.py:15) --------
 * Inference only.  Activations are fp32 NHWC between kernels.  Convolutions run on tedm_conv_igemm_fwd with every operand
 * split into a bf16 (hi, lo) pair (tedm_f32_split for activations; weights likewise) and the product taken as
 * hi*hi + lo*hi + hi*lo with fp32 accumulation: sources [x_hi, x_lo, x_hi] (tedm_conv_args.extra_src) against the
 * channel-concatenated weights [w_hi, w_hi, w_lo], out_dtype = 1.  The functions below restate the rest of
 * models/unet_model.py in fp32 with exact exp / division: same reference lines as their bf16 counterparts above. */
TEDM_API int tedm_f32_split(const float* x, void* hi_bf16, void* lo_bf16, int64_t n, tedm_stream_t stream);
TEDM_API int tedm_f32_stem_conv7x7(const float* x, const float* weight, const float* bias, float* out_nhwc, int batch, int cin,
                          int height, int width, int cout, tedm_stream_t stream);                     /* :267,334 */
TEDM_API int tedm_f32_gn_silu(const float* x, const float* gn_partial, int gn_parts, const float* gamma, const float* beta,
                     const float* scale_shift, int ss_stride, int ss_offset, const float* residual, float* out, int batch,
                     int hw, int channels, int groups, float eps, tedm_stream_t stream);              /* :126-135,175 */
TEDM_API int tedm_f32_layernorm(const float* x, const float* g, const float* residual, float* out, int64_t npix, int channels,
                       float eps, tedm_stream_t stream);                                               /* :52-61 */
TEDM_API int64_t tedm_f32_linear_attention_workspace(int batch, int n, int heads);
TEDM_API int tedm_f32_linear_attention(const float* qkv, float* out, float* workspace, int batch, int n, int heads, int dim_head,
                              float scale, tedm_stream_t stream);                                      /* :196-210 */
/* rnorm: scratch of batch * 2 * heads * dim_head floats (reciprocal L2 norms of q and k over the token axis) */
TEDM_API int tedm_f32_attention(const float* qkv, float* out, float* rnorm, int batch, int n, int heads, int dim_head, float scale,
                       tedm_stream_t stream);                                                          /* :229-241 */
TEDM_API int tedm_f32_final_conv1x1(const float* x, const float* weight, const float* bias, float* out_nchw, int batch, int hw,
                           int channels, int out_dim, tedm_stream_t stream);                          /* :331,368 */
TEDM_API int tedm_f32_add(const float* a, const float* b, float* out, int64_t n, tedm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TEDM_B200_H */
