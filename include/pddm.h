/*
 * pddm.h -- C ABI of libpddm_b200.so: the sm_100a (B200) kernels under the improved-diffusion hot path of
 * ArturPrzybysz/ProbabilisticDeepDiffusionModels.
 *
 * The reference has no FFI of its own (it is pure Python on stock torch ops); each entry point
 * below names the reference call site (file:line under /root/reference) whose arithmetic it replaces.  The
 * Python host (probabilisticdeepdiffusionmodels_b200/_lib.py) binds these with ctypes and registers them as
 * torch.library ops; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host.
 *   - the caller owns all memory (incl. workspaces); the library never allocates device memory and keeps no
 *     pointer after returning.  Every call only enqueues work on `stream` (no sync, CUDA-graph capturable).
 *   - return value: 0 on success, a negative pddm_status otherwise; nothing throws across the ABI.
 *   - activations inside the network are NHWC ("channels last") bf16 or fp32; the model boundary is NCHW fp32.
 *   - timesteps are 1-indexed (tables are read at [t-1]) exactly like the reference.
 */
#ifndef PDDM_H_
#define PDDM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* pddm_stream_t; /* cudaStream_t */

enum pddm_status {
  PDDM_OK = 0,
  PDDM_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, misaligned pointer */
  PDDM_ERR_UNSUPPORTED = -2, /* shape outside what the kernel supports (see each op) */
  PDDM_ERR_WORKSPACE = -3,   /* workspace_bytes smaller than pddm_*_workspace() */
  PDDM_ERR_CUDA = -4,        /* launch failed (cudaPeekAtLastError) */
  PDDM_ERR_ARCH = -5,        /* device is not sm_100 (no fallback path exists) */
  PDDM_ERR_TMA = -6          /* cuTensorMapEncodeTiled rejected the descriptor */
};

enum pddm_dtype { PDDM_F32 = 0, PDDM_BF16 = 1 };

int pddm_version(void);
const char* pddm_strerror(int status);
/* 0 if the current device can run the library (compute capability 10.x), PDDM_ERR_ARCH otherwise. */
int pddm_check_device(void);
int pddm_sm_count(void);
/* Leave n SMs free in every subsequent launch of the persistent kernels (tap-GEMMs, weight gradients, pipelined
 * GroupNorm size their grids to the SM count): room for a collective (NCCL) running concurrently on another stream,
 * whose CTAs would otherwise have to wait for a GEMM CTA to finish -- or make the GEMM's last CTAs wait for them.
 * Process-wide launch configuration (0 = whole device, the default); it never changes results beyond the summation
 * order of split-K weight gradients.  Grids captured into a CUDA graph keep the value they were captured with. */
int pddm_set_sm_reserve(int32_t n);
int pddm_get_sm_reserve(void);

/* ------------------------------------------------------------------------------------------------------
 * Diffusion math (fp32, NCHW, HBM-bound elementwise kernels)
 * ---------------------------------------------------------------------------------------------------- */

/* q_sample: x_t = sqrt_ab[t-1]*x0 + sqrt_1mab[t-1]*noise.  Replaces Engine.q_mean_std/get_q_t
 * (src/engine.py:251-261).  t: int64[B] (device) or NULL -> t_const for every sample. */
typedef struct {
  const float* x0;
  const float* noise;
  float* x_t;
  const int64_t* t;
  int32_t t_const;
  const float* alphas_hat_sqrt;         /* [T] */
  const float* one_min_alphas_hat_sqrt; /* [T] */
  int32_t B;
  int32_t chw; /* elements per sample */
} pddm_q_sample_params;
int pddm_q_sample(const pddm_q_sample_params* p, pddm_stream_t stream);

/* per-sample mean squared error  L_b = mean_chw (noise - pred)^2  (+ optional gradient).
 * Replaces get_loss / mean_flat (src/engine.py:263-277, src/utils.py:13-17).
 * pred may be the first `C` channels of a [B, c_total, HW] tensor (learned-variance models): pred_c_total,
 * c = channels compared, hw = pixels.  If grad_pred != NULL it receives
 *   grad_pred[b, c, :] = gscale[b] * 2 * (pred - noise) / (c*hw)   for c < C  (channels >= C untouched). */
typedef struct {
  const float* pred;
  const float* noise;
  float* per_sample; /* [B] or NULL */
  float* grad_pred;  /* same layout as pred, or NULL */
  const float* gscale; /* [B] upstream dL/dL_b, needed iff grad_pred */
  const float* grad_v_unit; /* [B, C, HW] or NULL: if set (c_total == 2C) the variance half of grad_pred becomes
                               grad_pred[b, C + c, :] = gscale[b] * v_scale * grad_v_unit[b, c, :]  (L_hybrid) */
  float v_scale;
  int32_t B, C, c_total, hw;
} pddm_sq_err_params;
int pddm_sq_err(const pddm_sq_err_params* p, pddm_stream_t stream);

/* The 12 coefficient tables of Engine.__init__ (src/engine.py:121-150), device-resident fp32 [T]. */
typedef struct {
  const float* betas;
  const float* alphas_sqrt;
  const float* posterior_variance;
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* denoising_coef;
  const float* alphas_hat_sqrt;
  const float* one_min_alphas_hat_sqrt;
  const float* posterior_log_variance_clipped; /* log(cat(pv[1:2], pv[1:]))   */
  const float* log_betas;
  int32_t T;
} pddm_tables;

/* One reverse step x_{t-1} = mean(x_t, eps) - sigma * z.  Replaces model_mean_from_epsilon /
 * xstart_from_epsilon / q_posterior / get_sigma / denoising_step (src/engine.py:354-397, 477-490).
 *   model_out: [B, c_out, HW] with c_out = C (eps) or 2C (eps | v, learned variance)
 *   z: noise or NULL (mean only); z is ignored at t == 1 (src/engine.py:389-394)
 *   t_step_dev: if non-NULL the step index is read from device memory (CUDA-graph replay), else t_step
 *   clip: clamp the x0 estimate to [-1,1] and go through the posterior mean (src/engine.py:366-367)
 *   sigma_mode: 0 = sqrt(beta), 1 = sqrt(beta_tilde), 2 = learned (needs c_out == 2C) */
typedef struct {
  const float* x_t;
  const float* model_out;
  const float* z;
  float* x_prev;
  pddm_tables tab;
  const int32_t* t_step_dev;
  int32_t t_step;
  int32_t B, C, c_out, hw;
  int32_t clip, sigma_mode;
} pddm_p_sample_params;
int pddm_p_sample_step(const pddm_p_sample_params* p, pddm_stream_t stream);

/* Chain bookkeeping for graph replay: t_dev[0] -= 1; t_vec[b] = (float)t_dev[0] for all b. */
int pddm_step_advance(int32_t* t_dev, float* t_vec, int32_t B, pddm_stream_t stream);

/* Variational-bound terms in bits/dim per sample.  Replaces normal_kl / discretized_gaussian_log_likelihood
 * (src/utils.py:50-115) and the drivers _calculate_L_T/_L_intermediate/_L_0 (src/engine.py:437-506).
 *   mode 0 (fixed variance, reference NLL eval): out[b] = t==1 ? -dll(x0; mean, log sigma)/ln2
 *                                                            : KL(q_post || N(mean, sigma^2))/ln2
 *          with mean = no-clip model mean (src/engine.py:375-381) and sigma per sigma_mode (0/1)
 *   mode 1 (learned variance, App. C.5): model_out = [eps | v]; mean through x0-estimate without clip;
 *          if grad_v != NULL also writes d out[b] / d v  ([B, C, HW])
 *   mode 2 (prior term L_T, src/engine.py:437-444): out[b] = KL(q(x_T|x0) || N(0,1))/ln2, uses x0 only. */
typedef struct {
  const float* x0;
  const float* x_t;
  const float* model_out;
  const int64_t* t; /* [B] */
  float* out;       /* [B] */
  float* grad_v;    /* mode 1 only, may be NULL */
  pddm_tables tab;
  int32_t B, C, c_out, hw;
  int32_t mode, sigma_mode;
} pddm_vlb_params;
int pddm_vlb_terms(const pddm_vlb_params* p, pddm_stream_t stream);

/* timestep_embedding (src/modules/nn.py:104-122): out[b] = [cos(t*f) | sin(t*f) | 0-pad], f_i = exp(-ln(max_period)*i/half).
 * t is int64 (t_is_float = 0) or float32 (sampling passes float timesteps, src/engine.py:386). */
int pddm_timestep_embedding(const void* t, int32_t t_is_float, void* out, int32_t out_dtype, int32_t B, int32_t dim,
                            float max_period, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Layout / elementwise helpers (NHWC activations)
 * ---------------------------------------------------------------------------------------------------- */
/* [B,C,HW] fp32 -> [B,HW,C] (dst dtype) and back. */
int pddm_nchw_to_nhwc(const float* src, void* dst, int32_t dst_dtype, int32_t B, int32_t C, int32_t HW,
                      pddm_stream_t stream);
int pddm_nhwc_to_nchw(const void* src, int32_t src_dtype, float* dst, int32_t B, int32_t C, int32_t HW,
                      pddm_stream_t stream);
/* dst[m, dst_off + c] = src[m, src_off + c] for c < C ; rows have ld_src / ld_dst elements (bf16). th.cat
 * (src/modules/unet.py:492) and its backward are expressed with this. */
int pddm_copy_channels(const void* src, int32_t ld_src, int32_t src_off, void* dst, int32_t ld_dst, int32_t dst_off,
                       int64_t M, int32_t C, pddm_stream_t stream);
/* nearest x2 upsample (src/modules/unet.py:79) of NHWC bf16 and its adjoint (2x2 sum). */
int pddm_upsample2x(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t stream);
int pddm_upsample2x_bwd(const void* grad_dst, void* grad_src, int32_t B, int32_t H, int32_t W, int32_t C,
                        pddm_stream_t stream);
/* Stride-2 phase split: dst[p, b, i, j, :] = src[b, 2i + (p>>1), 2j + (p&1), :]  (H, W even) and its inverse.
 * Lets the stride-2 Downsample conv (src/modules/unet.py:102) run as a tap-GEMM on unit-stride TMA boxes. */
int pddm_phase_split(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t stream);
int pddm_phase_merge(const void* src, void* dst, int32_t B, int32_t H, int32_t W, int32_t C, pddm_stream_t stream);
/* y = a + b (bf16, n elements; n % 8 == 0): gradient fan-in. */
int pddm_add_bf16(const void* a, const void* b, void* y, int64_t n, pddm_stream_t stream);
/* elementwise dtype conversion fp32 <-> bf16 (n elements). */
int pddm_convert(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, pddm_stream_t stream);
/* SiLU on an fp32 vector -> out dtype; and its backward (dx = dy * silu'(x)).  (src/modules/nn.py:13-15) */
int pddm_silu(const float* x, void* y, int32_t y_dtype, int64_t n, pddm_stream_t stream);
int pddm_silu_bwd(const float* x, const float* dy, float* dx, int64_t n, pddm_stream_t stream);
/* SiLU on a bf16 feature map (n elements, n % 8 == 0, 16-byte aligned) and its backward -- the reference's stand-alone
 * nn.SiLU between GroupNorm and conv (src/modules/unet.py:146-150, 163-166) for models assembled from the nn.py seam. */
int pddm_silu_map(const void* x, void* y, int64_t n, pddm_stream_t stream);
int pddm_silu_map_bwd(const void* x, const void* dy, void* dx, int64_t n, pddm_stream_t stream);
/* Column sums, deterministic (two-stage, fixed summation order; no atomics).  workspace: pddm_colsum_workspace bytes
 * for (rows per segment, segments, C) -- (M, 1, C) for pddm_colsum, (HW, B, C) for the per-sample variant.
 * out[c] (+)= sum_m x[m, c]  (x bf16 [M, ld], fp32 out [C]); used for conv bias gradients. */
size_t pddm_colsum_workspace(int64_t rows_per_segment, int32_t segments, int32_t C);
int pddm_colsum(const void* x, int32_t ld, int64_t M, int32_t C, float* out, int32_t accumulate, void* workspace,
                size_t workspace_bytes, pddm_stream_t stream);
/* out[b, c] = sum_{hw} x[b, hw, c] : per-sample column sums (gradient of the broadcast emb add). */
int pddm_colsum_per_sample(const void* x, int32_t B, int32_t HW, int32_t C, float* out, void* workspace,
                           size_t workspace_bytes, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * GroupNorm(32 groups, eps) [+ per-sample scale/shift] [+ SiLU]   (src/modules/nn.py:13-20,94-101;
 * ResBlock order src/modules/unet.py:188-200).  x: [B, HW, C] bf16 or fp32, y: bf16.
 *   z = gn(x)*gamma + beta ; if scale_shift: z = z*(1+scale[b,c]) + shift[b,c] ; y = silu ? z*sigmoid(z) : z
 * mean/rstd [B, G] fp32 are saved for the backward.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* x;
  int32_t x_dtype;
  const float* gamma;
  const float* beta;
  const float* scale; /* [B, ld_ss] or NULL */
  const float* shift; /* [B, ld_ss] or NULL */
  int32_t ld_ss;
  void* y; /* bf16 [B, HW, C] */
  float* mean;
  float* rstd;
  int32_t B, HW, C, G;
  float eps;
  int32_t silu;
  /* -- optional extensions (all zero = the plain contiguous single-tensor case) ----------------------------
   * Row strides in elements (0 = C): x is [B, HW, ldx], y is [B, HW, ldy] -- channel-slice views are allowed.
   * Two-source input (the concat-free th.cat([h, skip], 1) of src/modules/unet.py:492): channels [0, C_a) are read
   * from x, channels [C_a, C) from x2[..., 0 : C - C_a] (row stride ldx2).  bf16, no scale/shift only. */
  const void* x2;
  int32_t C_a;
  int32_t ldx, ldx2, ldy;
} pddm_gn_fwd_params;
size_t pddm_gn_silu_fwd_workspace(int32_t B, int32_t G);
int pddm_gn_silu_fwd(const pddm_gn_fwd_params* p, void* workspace, size_t workspace_bytes, pddm_stream_t stream);

/* Backward.  dy: bf16 [B,HW,C].  Outputs: dx (dx_dtype), dgamma/dbeta [C] fp32 (overwritten),
 * dx_colsum [B, C] fp32 or NULL (= sum_hw dx: gradient of a per-sample broadcast add in front of the norm,
 * i.e. of `h + emb_out` src/modules/unet.py:199 and of the producing conv's bias),
 * dscale/dshift [B, ld_ss] or NULL.  workspace: pddm_gn_silu_bwd_workspace(B, C) bytes. */
typedef struct {
  const void* x;
  int32_t x_dtype;
  const void* dy;
  const float* gamma;
  const float* beta;
  const float* scale;
  const float* shift;
  int32_t ld_ss;
  const float* mean;
  const float* rstd;
  void* dx;
  int32_t dx_dtype;
  float* dgamma;
  float* dbeta;
  float* dx_colsum;
  float* dscale;
  float* dshift;
  int32_t B, HW, C, G;
  int32_t silu;
  /* -- optional extensions (all zero = the plain contiguous single-tensor case; bf16, no scale/shift only) ----
   * x2/C_a/ldx/ldx2: two-source input as in the forward; lddy: row stride of dy.
   * gres [B, HW, ld_gres] bf16: a gradient that reaches x by another route (the residual / skip branch,
   *   src/modules/unet.py:201,234) and is added to dx before it is stored: dx_total = dx_norm + gres.
   * dx / dx2 receive channels [0, C_a) / [C_a, C) (dx2 = NULL: dx is one [B, HW, ld_dx] tensor of all C channels);
   *   dx_accumulate / dx2_accumulate: add into the destination (bulk-tensor reduce) instead of overwriting it --
   *   the fan-in of a tensor that also feeds a later th.cat.
   * part_dgamma / part_dbeta [B, ld_part] fp32: per-SAMPLE partial sums (the caller reduces them over the batch,
   *   one launch for a whole network); dgamma / dbeta may then be NULL.
   * dx_colsum [B, ld_colsum], dx_colsum2 [B, ld_colsum2]: sum_hw dx_total per segment; *_accumulate adds to it. */
  const void* x2;
  int32_t C_a;
  int32_t ldx, ldx2, lddy;
  const void* gres;
  int32_t ld_gres;
  void* dx2;
  int32_t ld_dx, ld_dx2;
  int32_t dx_accumulate, dx2_accumulate;
  float* part_dgamma;
  float* part_dbeta;
  int32_t ld_part;
  float* dx_colsum2;
  int32_t ld_colsum, ld_colsum2;
  int32_t colsum_accumulate, colsum2_accumulate;
} pddm_gn_bwd_params;
/* > 0 (the number of pipeline slots) if the persistent bulk-tensor GroupNorm kernels take this shape with `ntens`
 * tensors resident per item (forward: 1; backward: 2, or 3 with gres), 0 if the call would use the older
 * cluster kernels (which support none of the extensions above). */
int pddm_gn_pipe_slots(int32_t B, int32_t HW, int32_t C, int32_t G, int32_t C_a, int32_t ntens);
size_t pddm_gn_silu_bwd_workspace(int32_t B, int32_t C);
int pddm_gn_silu_bwd(const pddm_gn_bwd_params* p, void* workspace, size_t workspace_bytes, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Convolution / linear as a tcgen05 "tap GEMM"   (conv_nd / linear, src/modules/nn.py:23-40; call sites
 * src/modules/unet.py:70,102,149,163,174,219,221,342,344,153)
 *
 *   y[b, oh, ow, n] = bias[n] + bcast[b, n] + residual[b, oh, ow, n]
 *                     + sum_{tap} sum_{c} x[b + tap_db, h + tap_dh, w + tap_dw, c] * w[n, tap, c]
 *   with (oh, ow) = (h*out_sh + out_oh, w*out_sw + out_ow) and (b, h, w) ranging over [B, H, W].
 *
 * x is a 4-D NHWC bf16 tensor [x_NB, H, W, ldx] (ldx >= Cin: channel-slice views allowed, ldx % 8 == 0); reads
 * outside [0,H)x[0,W) are zero (TMA out-of-bounds fill) which implements padding=1.  w is bf16
 * [Cout, w_ntaps*Cin] (tap-major, channel-minor rows); tap i multiplies weight slot tap_w[i] (a stride-2 data
 * gradient uses a subset of the 9 slots per output phase).  3x3/s1: 9 taps (dh,dw in -1..1); 1x1 and linear: 1 tap;
 * stride-2: taps over the phase-split input (tap_db = phase*B); dgrad: the same kernel on the flipped/transposed
 * weight pack.  Cin % 32 == 0, Cout % 8 == 0.  Accumulation is fp32 in tensor memory.
 * ---------------------------------------------------------------------------------------------------- */
#define PDDM_MAX_TAPS 9
typedef struct {
  const void* x;
  const void* w;
  const float* bias;     /* [Cout] or NULL */
  const float* bcast;    /* [B, ld_bcast] or NULL */
  const void* residual;  /* [B, out_H, out_W, Cout] or NULL */
  void* y;               /* [B, out_H, out_W, Cout] */
  int32_t ld_bcast;
  int32_t res_dtype, y_dtype;
  int32_t x_NB, B, H, W, Cin, ldx, Cout;
  int32_t ntaps;
  int32_t tap_db[PDDM_MAX_TAPS], tap_dh[PDDM_MAX_TAPS], tap_dw[PDDM_MAX_TAPS];
  int32_t tap_w[PDDM_MAX_TAPS]; /* weight slot read by each tap: w[n, tap_w[tap], c]; w has w_ntaps slots per row */
  int32_t w_ntaps;
  int32_t out_H, out_W, out_sh, out_sw, out_oh, out_ow;
  /* optional second source (concat-free th.cat([h, skip], 1), src/modules/unet.py:492): input channels [0, Cin_a) are
   * read from x, channels [Cin_a, Cin) from x2[..., 0 : Cin - Cin_a] (same [x_NB, H, W] extents, row stride ldx2).
   * Cin_a must be a multiple of the K-block (64, or 32 when Cin % 64 != 0).  NULL = single source. */
  const void* x2;
  int32_t Cin_a, ldx2;
} pddm_conv_params;
int pddm_conv2d_fwd(const pddm_conv_params* p, pddm_stream_t stream);

/* Weight gradient:  dw[n, tap, c] = sum_{b,h,w} dy[b, h, w, n] * x[b + tap_db, h + tap_dh, w + tap_dw, c]
 * (dy indexed in the conv's tile space [B, H, W]; for strided outputs pass the phase view).  Split-K partial
 * sums go to the workspace and are reduced deterministically into dw_out, fp32, laid out either packed
 * [Cout, ntaps, Cin] (layout 0) or as the parameter [Cout, Cin, ntaps] (layout 1 = torch Conv2d weight). */
typedef struct {
  const void* x;
  const void* dy;
  float* dw;
  int32_t x_NB, B, H, W, Cin, ldx, Cout, lddy;
  int32_t ntaps;
  int32_t tap_db[PDDM_MAX_TAPS], tap_dh[PDDM_MAX_TAPS], tap_dw[PDDM_MAX_TAPS];
  int32_t dw_layout;
  int32_t accumulate; /* dw += instead of dw = */
  /* layout 1 only: dw is the channel slice [dw_c0, dw_c0 + Cin) of a parameter with dw_ldc input channels
   * (the two halves of a conv whose input is a two-source concat); 0 = the whole parameter (dw_ldc = Cin). */
  int32_t dw_ldc, dw_c0;
} pddm_wgrad_params;
size_t pddm_conv2d_wgrad_workspace(const pddm_wgrad_params* p);
int pddm_conv2d_wgrad(const pddm_wgrad_params* p, void* workspace, size_t workspace_bytes, pddm_stream_t stream);

/* fp32 parameter [Cout, Cin, kh*kw] -> bf16 GEMM operand, zero-padded to [Cout_pad, ., Cin_pad].
 *   mode 0 (forward): dst[n, tap, c]            = w[n, c, tap]     dst is [Cout_pad, ntaps, Cin_pad]
 *   mode 1 (dgrad):   dst[c, ntaps-1-tap, n]    = w[n, c, tap]     dst is [Cin_pad, ntaps, Cout_pad]
 *                     (flipped taps, transposed channels) */
int pddm_pack_conv_weight(const float* w, void* dst, int32_t Cout, int32_t Cin, int32_t ntaps, int32_t mode,
                          int32_t Cout_pad, int32_t Cin_pad, pddm_stream_t stream);

/* Re-pack every weight of a model in ONE launch (after each optimiser step).  descs: device array of
 * pddm_pack_desc; blocks: device array of int32 pairs (descriptor index, tile index) -- one thread block packs one
 * tile of PDDM_PACK_TILE output channels x PDDM_PACK_TILE input channels x all taps of one tensor; tiles are numbered
 * row-major over (ceil(Cout_pad/TILE), ceil(Cin_pad/TILE)).  Cout_pad and Cin_pad must be multiples of 8.
 * max_ntaps: the largest ntaps among the descriptors (sizes the shared-memory tile). */
#define PDDM_PACK_TILE 32
typedef struct {
  const float* src;
  void* dst; /* bf16 */
  int32_t Cout, Cin, ntaps, mode, Cout_pad, Cin_pad;
  int32_t ld_dst;   /* elements between consecutive destination rows; 0 = dense (ntaps * Cin_pad | ntaps * Cout_pad):
                       lets several parameters be packed side by side into one wide GEMM operand */
  int32_t reserved;
  void* dst2;       /* optional: the pack of the OTHER mode (same paddings, dense rows), written from the same tile
                       of the fp32 source -- a conv needs both the forward and the data-gradient operand */
} pddm_pack_desc;
int pddm_pack_weights_multi(const void* descs, const void* blocks, int32_t nblocks, int32_t max_ntaps,
                            pddm_stream_t stream);
/* out[c] = sum_m x[m, c] for a small fp32 matrix [M, C] (deterministic). */
int pddm_colsum_f32(const float* x, int32_t M, int32_t C, float* out, pddm_stream_t stream);
/* Per-sample column sums with strides: out[b, c] (+)= sum_hw x[b, hw, c] for x bf16 [B, HW, ldx] (c < C),
 * out fp32 [B, ld_out]; one launch, fixed summation order.  (bias gradients, src/modules/unet.py:102,219) */
int pddm_colsum_rows(const void* x, int32_t ldx, int32_t B, int32_t HW, int32_t C, float* out, int32_t ld_out,
                     int32_t accumulate, pddm_stream_t stream);
/* Batch fold of a matrix of per-sample partial sums: dst[j] = sum_b ps[b, src_of[j]] for j < n (src_of = NULL:
 * identity), fixed order.  ONE launch turns every per-sample partial of a backward pass (GroupNorm dgamma / dbeta,
 * bias column sums) into parameter gradients. */
int pddm_batch_fold(const float* ps, int32_t ld, int32_t B, const int32_t* src_of, int32_t n, float* dst,
                    pddm_stream_t stream);
/* dst[r, c] = (bf16) src[r, c] for strided row-major matrices (rows x cols, cols % 8 == 0, 16-byte aligned rows). */
int pddm_convert_rows(const float* src, int32_t ld_src, void* dst, int32_t ld_dst, int32_t rows, int32_t cols,
                      pddm_stream_t stream);

/* Helpers that put the two "thin" convolutions on the tensor cores as well: the stem conv (Cin <= 4, reads the
 * NCHW fp32 model input, src/modules/unet.py:353) becomes a K=32 GEMM over an im2col patch matrix, and the head conv
 * (Cout <= 8, writes the NCHW fp32 model output, src/modules/unet.py:440) runs with zero-padded output channels.
 *   im2col3x3: out[b,h,w, ci*9+tap] = x[b,ci,h+dh,w+dw] (bf16 [B,H,W,Kp], zero for k >= Cin*9 and outside the image)
 *   nchw_to_nhwc_padded: NCHW fp32 [B,C,HW] -> NHWC bf16 [B,HW,Cp], channels >= C zero
 *   nhwc_slice_to_nchw: first C channels of NHWC fp32 [B,HW,ld] -> NCHW fp32 [B,C,HW] */
int pddm_im2col3x3(const float* x_nchw, void* out, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Kp,
                   pddm_stream_t stream);
int pddm_nchw_to_nhwc_padded(const float* src, void* dst, int32_t B, int32_t C, int32_t HW, int32_t Cp,
                             pddm_stream_t stream);
int pddm_nhwc_slice_to_nchw(const float* src, float* dst, int32_t B, int32_t C, int32_t HW, int32_t ld,
                            pddm_stream_t stream);
/* Sample post-processing in one pass (replaces src/modules/fid_score.py:15-27 + src/datasets/data.py:108-128 on the
 * host): out[b, hw, c] = uint8(255 * clip(x[b, c, hw] * std[c] + mean[c], 0, 1)); mean/std (device, double [C]) may both
 * be NULL (= the reference's unnormalize(normalize=None, clip=True)).  Double arithmetic and truncation reproduce the
 * reference's numpy expression bit for bit. */
int pddm_images_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int32_t B, int32_t C, int32_t HW, const double* mean,
                         const double* stdv, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * QKVAttention (src/modules/unet.py:237-256): per (sample, head), softmax_fp32((q*s)(k*s)^T) v, s = d^-1/4.
 * qkv: bf16 [B, T, 3*heads*d] with the reference's head-major [q|k|v] channel packing; out: bf16 [B, T, heads*d].
 * lse [B, heads, T] fp32 is saved for the backward.  d in {32, 64, 96, 128}, T <= 256.
 * The backward of a head with more than 128 keys runs as two cooperating CTAs (one per 128-key tile) that exchange
 * their fp32 partial dQ products through the caller-owned workspace `ws` (pddm_attn_bwd_workspace_bytes, 0 when
 * T <= 128; contents are scratch).
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  const void* qkv;
  void* out;
  float* lse;
  int32_t B, T, heads, d;
} pddm_attn_fwd_params;
int pddm_attn_fwd(const pddm_attn_fwd_params* p, pddm_stream_t stream);
typedef struct {
  const void* qkv;
  const void* out;
  const void* dout;
  const float* lse;
  void* dqkv; /* bf16 [B, T, 3*heads*d] */
  int32_t B, T, heads, d;
  void* ws;   /* >= pddm_attn_bwd_workspace_bytes(B, T, heads, d) bytes, 16-byte aligned; may be NULL when T <= 128 */
  int64_t ws_bytes;
} pddm_attn_bwd_params;
int64_t pddm_attn_bwd_workspace_bytes(int32_t B, int32_t T, int32_t heads, int32_t d);
int pddm_attn_bwd(const pddm_attn_bwd_params* p, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Opt-in high-precision forward (highprec.py): fp32 activations, every GEMM operand split into bf16 hi + lo terms
 * (x = hi + lo, hi = bf16(x), lo = bf16(x - hi)) so that W x = W_hi x_hi + W_hi x_lo + W_lo x_hi runs on the bf16
 * tap-GEMM with fp32 accumulation.  pddm_split_bf16: n elements.  pddm_gn_split_f32: GroupNorm (+SiLU) of an fp32
 * [B, HW, C] tensor, statistics written to mean / rstd [B, G], result written as the split pair (and, if y32 != NULL,
 * as fp32).  pddm_attn_fwd_f32: QKVAttention (src/modules/unet.py:237-256) in plain fp32 on [B, T, 3*heads*d].
 * ---------------------------------------------------------------------------------------------------- */
int pddm_split_bf16(const float* x, void* hi, void* lo, int64_t n, pddm_stream_t stream);
int pddm_gn_split_f32(const float* x, const float* gamma, const float* beta, float* mean, float* rstd, void* hi, void* lo,
                      float* y32, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t silu,
                      pddm_stream_t stream);
int pddm_attn_fwd_f32(const float* qkv, float* out, int32_t B, int32_t T, int32_t heads, int32_t d, pddm_stream_t stream);

/* ------------------------------------------------------------------------------------------------------
 * Fused Adam (+ EMA) over a flat fp32 parameter arena (torch.optim.Adam defaults, src/engine.py:238-248;
 * Ema.update src/modules/ema.py:21-33).  ema may be NULL.  step is the 1-based step count.
 * ---------------------------------------------------------------------------------------------------- */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* ema;
  int64_t n;
  float lr, beta1, beta2, eps, weight_decay, ema_decay, grad_scale;
  int32_t step;
  const int32_t* step_dev; /* if non-NULL the step count is read from device memory (CUDA-graph replay) */
  const float* lr_dev;     /* if non-NULL overrides lr (LR schedulers under graph replay) */
} pddm_adam_params;
int pddm_adam_ema_step(const pddm_adam_params* p, pddm_stream_t stream);
/* The same update over MANY separately allocated tensors in ONE launch (a model's parameter list as autograd
 * leaves it: one gradient tensor per parameter).  descs: device array of pddm_adam_tensor; blocks: device array of
 * int32 pairs (tensor index, first element) -- one thread block updates PDDM_ADAM_CHUNK consecutive elements of one
 * tensor.  The scalar fields of hyper (lr ... lr_dev) apply to every tensor; its pointer fields and n are ignored. */
#define PDDM_ADAM_CHUNK 8192
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* ema; /* may be NULL */
  int64_t n;
} pddm_adam_tensor;
int pddm_adam_ema_multi(const void* descs, const void* blocks, int32_t nblocks, const pddm_adam_params* hyper,
                        pddm_stream_t stream);
/* *counter += delta (single thread); advances device-side step counters between graph replays. */
int pddm_counter_add(int32_t* counter, int32_t delta, pddm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PDDM_H_ */
