/*
 * mmdyn_b200 — C ABI of the B200-native (sm_100a) kernels behind the cnn-vae / cnn-mvae
 * training + inference step of SAIC-MONTREAL/multimodal-dynamics.
 *
 * The reference has no FFI of its own (it is 100 % Python calling torch ops), so
 * every entry point below cites the torch call site in the reference that it replaces
 * (paths relative to the reference root, mmdyn/pytorch/...).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers + sizes only; all data pointers are DEVICE pointers unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - no allocation, no synchronisation, no ownership transfer inside any call;
 *   - return 0 on success, <0 on error (mmdyn_last_error() gives the message); never throws;
 *   - "op" tensors are IEEE fp16 (10-bit mantissa, = TF32 operand precision) in NHWC layout,
 *     accumulators / statistics / losses / optimizer state are fp32.
 */
#ifndef MMDYN_B200_H_
#define MMDYN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMDYN_MAX_TAPS 16
#define MMDYN_MAX_PHASES 4
#define MMDYN_MAX_GROUPS 8

/* --- library ------------------------------------------------------------------------------- */
const char* mmdyn_last_error(void);
int mmdyn_version(void);
/* number of kernels launched by this library in this process since load (bench gpu_launches) */
long long mmdyn_launch_count(void);
/* one-time per device: raises the dynamic shared memory limits of the tcgen05 kernels */
int mmdyn_init(int device);

/* --- implicit-GEMM convolution / transposed convolution / linear (tcgen05 + TMEM + TMA) ------
 * replaces nn.Conv2d / nn.ConvTranspose2d / nn.Linear forward and their input-gradients:
 *   models/vae.py:198-216 (encoder convs, fc, heads), :263-277 (decoder upsample + deconvs),
 *   and the autograd backward of the same (problems/problems.py:153 loss.backward()).
 *
 *   out[row(m), n] = sum_{tap t, channel c} A[gather(m, t), c] * W[phase][n][t*Cin + c] (+ bias[n])
 *
 * rows m enumerate a "virtual grid" (image, yv, xv); tap t reads input pixel
 * (yv*s_in + dy[t], xv*s_in + dx[t]) (zero outside the image); the result goes to output pixel
 * (yv*s_out + off_y[phase], xv*s_out + off_x[phase]).  This one form covers stride-2 convs,
 * sub-pixel phases of stride-2 transposed convs, stride-1 k4 (5<->8) layers, linears (one tap)
 * and all of their dgrads.
 */
typedef struct mmdyn_igemm_desc {
  const void* A;        /* fp16 NHWC gathered operand                                        */
  const void* W;        /* fp16 packed weights [n_phases*N][ntaps*Cin], K contiguous (TMA)    */
  void* out;            /* see out_mode                                                        */
  const float* bias;    /* [N] or NULL                                                         */
  int32_t n_img;        /* images (rows of the virtual grid = n_img*P)                         */
  int32_t P;            /* virtual pixels per image = OYv*OXv                                  */
  int32_t OXv;          /* virtual grid width                                                  */
  int32_t IH, IW;       /* input spatial size                                                  */
  int32_t a_pix_stride; /* elements between consecutive input pixels (>= Cin)                  */
  int32_t Cin;          /* channels per tap (multiple of 8); ntaps*Cin multiple of 64          */
  int32_t s_in;         /* input stride                                                        */
  int32_t ntaps;
  int32_t n_phases;
  int8_t tap_dy[MMDYN_MAX_PHASES][MMDYN_MAX_TAPS];
  int8_t tap_dx[MMDYN_MAX_PHASES][MMDYN_MAX_TAPS];
  int32_t N;            /* output channels (multiple of block_n)                               */
  int32_t block_n;      /* 16, 32, 64, 128 or 256                                              */
  int32_t ksplit;       /* >1: split K across CTAs, out_mode must be 2 (atomic fp32)           */
  int32_t row_mode;     /* 0: tile = 128 consecutive rows (image-major);
                           1: tile = one virtual pixel x 128 images, invalid taps skipped      */
  int32_t out_mode;     /* 0: fp16 rows, 1: fp32 rows, 2: fp32 rows atomicAdd,
                           3: fp32 NCHW planes from merged 2x2 sub-pixel phases (N = 16, 12 used),
                           4: fp16 NHWC from merged 2x2 sub-pixel phases: n = (ph*2+pw)*ldc + c goes
                              to pixel (2*yv+ph, 2*xv+pw), channel c (ldc = channels, multiple of 16),
                           5: as 3 with the BCE loss + logit gradient fused (bce_* fields below)        */
  int32_t OH, OW;       /* output spatial size                                                 */
  int32_t s_out;
  int32_t off_y[MMDYN_MAX_PHASES], off_x[MMDYN_MAX_PHASES];
  int32_t ldc;          /* elements between consecutive output pixels                          */
  int32_t a_row_stride; /* elements between input rows; 0 = dense (a_pix_stride*IW)             */
  int32_t a_img_stride; /* elements between images of A; 0 = dense (row stride * IH).  With
                           a_pix_stride < Cin the "pixels" overlap: a window of Cin/a_pix_stride
                           physical pixels is one tap (used by the 3-channel logits layer, whose
                           4 x-taps of 8 padded channels are one 64-byte window; TMA path only)   */
  /* out_mode 5 only — the logits layer with the reconstruction loss fused into its epilogue
   * (vae.py:277 + problems.py:409-413, 431-449): per output pixel BCE-with-logits against
   * bce_target (and bce_mask), summed into bce_loss[bce_slot[group]], gradient
   * gscale*(sigmoid(x*m) - t*m)*m written as fp16 NHWC4 with a one-pixel border (see mmdyn_bce_logits,
   * pad = 1).  group = image / bce_rows_per_group; every group is compared with the same
   * bce_rows_per_group target images.  The fp32 NCHW logits themselves are only stored for images
   * in [logit_row_lo, logit_row_hi). */
  const float* bce_target;
  const float* bce_mask;       /* optional */
  void* bce_dlogits;           /* optional (forward-only evaluation) */
  float* bce_loss;
  float bce_gscale;
  int32_t bce_rows_per_group;
  int32_t bce_slot[MMDYN_MAX_GROUPS]; /* < 0: group carries no loss (its gradient rows are left untouched) */
  int32_t logit_row_lo, logit_row_hi;
  /* patch_mode != 0 (merged 3x3-tap layers only: s_in = 1, the 9 taps dy,dx in {-1,0,1} row-major first,
   * out_mode 3 / 4 / 5, OXv = IW in {8,16,32}): shared-memory patch reuse — one activation box per filter
   * column serves its three row taps, and with out_mode 4 only the structurally non-zero (tap, sub-pixel
   * phase) weight blocks are fetched and multiplied (vae.py:271-277 forward; dgrad of vae.py:200-203). */
  int32_t patch_mode;
  /* optional BatchNorm statistics of the raw output fused into the epilogue (out_mode 4 + patch_mode: channels =
   * ldc <= 64; out_mode 0 without patch_mode: channels = N in [64, 256], one phase, no K split) —
   * bn_sums[group][channels][2] += {sum x, sum x^2} of the fp16-rounded outputs, group = image / bn_rows_per_group,
   * which must be a multiple of the images per tile (replaces mmdyn_bn_stats for this layer's output; caller
   * zeroes bn_sums). */
  int32_t bn_rows_per_group;
  float* bn_sums;
  int32_t s_in_x;       /* input stride along x when it differs from s_in (0 = s_in): the 4-channel logit
                           gradient is addressed in 16-byte pixel PAIRS along x (TMA strides are multiples of
                           16 bytes), so its window tap advances 1 pair per virtual pixel but 2 rows per row  */
} mmdyn_igemm_desc;
int mmdyn_igemm(const mmdyn_igemm_desc* d, void* stream);

/* --- weight gradient (tcgen05, both operands MN-major) ---------------------------------------
 * replaces the weight-gradient half of autograd for the same layers (problems/problems.py:153).
 *   dW[n][t*Cg + c] += scale * sum_m Nat[m][n] * G[gather(m, t), c]
 * `Nat` is read in natural row order (rows = virtual grid), `G` is gathered with the taps.
 */
typedef struct mmdyn_wgrad_desc {
  const void* G;        /* fp16 NHWC gathered operand, Cg channels per tap                     */
  const void* Nat;      /* fp16 [n_img*P][nat_stride] natural-order operand                    */
  float* dW;            /* fp32 [Cn][ntaps*Cg], accumulated with atomics (caller zeroes)       */
  int32_t n_img, P, OXv, IH, IW;
  int32_t g_pix_stride; /* elements between consecutive pixels of G                            */
  int32_t Cg;           /* channels per tap of G (multiple of 8), ntaps*Cg multiple of 128
                           (or Cg = 16, ntaps = 4: 64 columns)                                 */
  int32_t s_in, ntaps;
  int8_t tap_dy[MMDYN_MAX_TAPS];
  int8_t tap_dx[MMDYN_MAX_TAPS];
  int32_t Cn;           /* channels of Nat used (16, 32, or multiple of 64; <=256 per launch)  */
  int32_t nat_stride;   /* elements between consecutive rows of Nat                            */
  int32_t ldw;          /* row pitch of dW in elements (>= ntaps*Cg)                           */
  int32_t row_splits;   /* CTAs along the reduction                                            */
  float scale;
  int32_t g_row_stride; /* elements between rows / images of G; 0 = dense.  Same overlapping-   */
  int32_t g_img_stride; /* window convention as mmdyn_igemm_desc (TMA path only)                */
  int32_t s_in_x;       /* as mmdyn_igemm_desc.s_in_x                                           */
} mmdyn_wgrad_desc;
int mmdyn_wgrad(const mmdyn_wgrad_desc* d, void* stream);

/* --- first encoder layer: Conv2d(3,32,4,2,1) on the fp32 NCHW input ---------------------------
 * replaces vae.py:198 forward and its weight gradient (the input needs no gradient).
 * W is the fp16 packed [32][64] matrix (K order (ci,kh,kw), 48 used).  out: fp16 NHWC (B,32,32,32).
 */
/* act_out (nullable): Swish of the fp16 output, same layout (vae.py:199) — saves the stand-alone activation pass */
int mmdyn_conv1_fwd(const float* x_nchw, const void* Wp, void* out, void* act_out, int n_img, void* stream);
int mmdyn_conv1_wgrad(const float* x_nchw, const void* dRaw, float* dW /*[32][48]*/, int n_img,
                      float scale, int row_splits, void* stream);

/* --- grouped BatchNorm2d (training statistics) + Swish ----------------------------------------
 * replaces nn.BatchNorm2d + Swish, vae.py:201-208, 269-276, 331-334, forward and backward.
 * Tensors are [G groups][rows_per_group][C] fp16 (NHWC flattened); every group (= one
 * sub-sampled MVAE pass, problems.py:478-529) has its own batch statistics.
 */
int mmdyn_bn_stats(const void* x, float* sums /*[G][C][2] zeroed*/, int G, int rows_per_group,
                   int C, void* stream);
/* mean/invstd -> scale/shift (a = gamma*invstd, b = beta - mean*a); running stats updated
 * `stat_repeat` times per group, in group order, with `momentum` and the unbiased variance (torch
 * semantics; stat_repeat > 1 replays the update for passes that share one trunk evaluation). */
int mmdyn_bn_finalize(const float* sums, const float* gamma, const float* beta, float* ab /*[G][C][2]*/,
                      float* mean_invstd /*[G][C][2]*/, float* running_mean, float* running_var,
                      int G, int rows_per_group, int C, float eps, float momentum, int stat_repeat,
                      long long* num_batches_tracked /* nullable int64 counter += G*stat_repeat */, void* stream);
/* mmdyn_bn_finalize + mmdyn_bn_swish_fwd in one launch (the streaming kernel derives scale/shift from
 * `sums` itself, publishes ab / mean_invstd for the backward and updates the running statistics). */
int mmdyn_bn_finalize_swish_fwd(const void* x, const float* sums, const float* gamma, const float* beta, float* ab,
                                float* mean_invstd, float* running_mean, float* running_var,
                                long long* num_batches_tracked, void* y, int G, int rows_per_group, int C, float eps,
                                float momentum, int stat_repeat, void* stream);
/* y = swish(a*x + b); ab == NULL means identity affine (plain Swish) */
int mmdyn_bn_swish_fwd(const void* x, const float* ab, void* y, int G, int rows_per_group, int C,
                       void* stream);
/* pass 1 of the BN+Swish backward (read-only): with dU = dY * swish'(a*x+b),
 * sums2[g][c] = {sum dU, sum dU*xhat} (zeroed by caller); dY is left untouched.
 * ab == NULL: identity affine (plain Swish): dY := dY * swish'(x) in place, no sums. */
int mmdyn_bn_swish_bwd_reduce(const void* x, const float* ab, const float* mean_invstd, void* dY,
                              float* sums2, int G, int rows_per_group, int C, void* stream);
/* pass 2: dX = a*(dU - mean(dU) - xhat*mean(dU*xhat)) written over dY (dU recomputed from dY and x);
 * dgamma += sum dU*xhat, dbeta += sum dU (summed over groups, times grad_unscale);
 * coef_scratch: [G][C][4] floats */
int mmdyn_bn_bwd_apply(const void* x, const float* ab, const float* mean_invstd, const float* sums2,
                       void* dU, float* dgamma, float* dbeta, float* coef_scratch, int G,
                       int rows_per_group, int C, float grad_unscale, void* stream);
/* the same pass with dX written to a second buffer instead of over dU: rows are image rows of 2^row_w_log2 pixels, stored
 * with one extra pixel on each side, [rows / 2^w][2^w + 2][C] (the caller zeroes the buffer once; the border is never
 * written) — the layout in which the k4/s2 data and weight gradients of the producing layer read their 4 x-taps as two
 * 128-byte pixel pairs (plan.deconv_s2_plan, mmdyn_igemm_desc.s_in_x) */
int mmdyn_bn_bwd_apply_padded(const void* x, const float* ab, const float* mean_invstd, const float* sums2,
                              const void* dU, void* dX_padded, int row_w_log2, float* dgamma, float* dbeta,
                              int G, int rows_per_group, int C, float grad_unscale, void* stream);

/* --- fc tail: bias + Swish + Dropout mask (vae.py:210-214) ------------------------------------
 * raw [B][C] fp32 (igemm output incl. bias); for each of n_masks masks (fp32, values 0 or 1/(1-p),
 * or NULL = no dropout) writes h[m][B][C] fp16 = swish(raw) * mask_m.  Backward: dRaw (fp16) =
 * swish'(raw) * sum_m mask_m * dH[m]. */
int mmdyn_swish_dropout_fwd(const float* raw, const float* const* masks, void* h, int n_masks,
                            int B, int C, void* stream);
int mmdyn_swish_dropout_bwd(const float* raw, const float* const* masks, const float* dH,
                            void* dRaw, int n_masks, int B, int C, void* stream);

/* --- ProductOfExperts + reparametrisation + KL (vae.py:52-61, 139-157, 311-328; problems.py:406,429)
 * experts: up to 4 (mu_e, logvar_e) pairs [B][D] fp32 with row stride `ld`; the prior expert
 * N(0, I) is implicit when use_prior != 0 (MVAE); with use_prior == 0 and one expert this is the
 * plain VAE posterior (vae.py:84-85).  Outputs: mu, logvar (posterior), z = eps*exp(0.5 logvar)+mu
 * (fp32 [B][D]) and zh / zh2 (optional fp16 copies of z: the operand rows of up to two decoders); kl_sum += -0.5*sum(1+lv-mu^2-e^lv).
 */
int mmdyn_poe_fwd(const float* const* mu_e, const float* const* lv_e, int n_experts, int use_prior,
                  int ld, const float* eps, float* mu, float* lv, float* z, void* zh, void* zh2,
                  float* kl_sum, int B, int D, void* stream);
/* backward: dmu_e/dlv_e[e] (+)= grads through z, KL and (optionally) direct upstream gradients of
 * the posterior.  kl_coef = kl_weight * grad scale; dz[0..2]: up to three (nullable) fp32 [B][D]
 * gradients w.r.t. z, one per decoder, summed here; dmu_in / dlv_in (nullable, [B][D]) are added
 * to dL/dmu and dL/dlogvar (loss terms computed outside the library on the posterior);
 * accumulate != 0 adds into dmu_e/dlv_e (row stride ld_out) */
int mmdyn_poe_bwd(const float* const* mu_e, const float* const* lv_e, int n_experts, int use_prior,
                  int ld, const float* eps, const float* const* dz, const float* dmu_in,
                  const float* dlv_in, float kl_coef, float* const* dmu_e, float* const* dlv_e,
                  int ld_out, int accumulate, int B, int D, void* stream);
/* All sub-sampled passes of a step (problems.py:478-529: 3, or 7 with --use-pose) in ONE launch each way: pass k
 * reads / writes the pointers of passes[k] (same meaning as the arguments of mmdyn_poe_fwd / mmdyn_poe_bwd).  In the
 * backward several passes may accumulate into the same expert-gradient rows (the pose expert serves 4 passes):
 * accumulation is atomic there. */
#define MMDYN_MAX_POE_PASSES 8
typedef struct mmdyn_poe_pass {
  const float* mu_e[4];
  const float* lv_e[4];
  int32_t n_experts;
  const float* eps;
  /* forward outputs */
  float* mu;
  float* lv;
  float* z;
  void* zh;
  void* zh2;
  float* kl_sum;
  /* backward */
  const float* dz[3];
  const float* dmu_in;
  const float* dlv_in;
  float* dmu_e[4];
  float* dlv_e[4];
} mmdyn_poe_pass;
int mmdyn_poe_fwd_multi(const mmdyn_poe_pass* passes, int n_passes, int use_prior, int ld, int B, int D, void* stream);
int mmdyn_poe_bwd_multi(const mmdyn_poe_pass* passes, int n_passes, int use_prior, int ld, float kl_coef, int ld_out,
                        int accumulate, int B, int D, void* stream);

/* --- reconstruction losses (problems.py:409-413, 431-449, 499-503, 535) -----------------------
 * BCE-with-logits, reduction 'sum' into loss_sum[0]; dlogits (fp16 NHWC, 4 channels per pixel,
 * 3 used) = gscale*(sigmoid(x) - t)*m.  mask (optional, NCHW fp32) multiplies logits and targets
 * as the reference does.  logits/target: NCHW fp32 (n,3,H,W).  pad > 0: dlogits images are
 * (H+2*pad) x (W+2*pad) with a border that is never written (the caller zeroes it once): the
 * layout the decoder backward reads as overlapping 4-pixel windows (mmdyn_igemm_desc.a_row_stride). */
int mmdyn_bce_logits(const float* logits, const float* target, const float* mask, float* loss_sum,
                     void* dlogits_nhwc4, float gscale, int n, int H, int W, int pad, void* stream);
/* Same loss for callers outside the fused step (Reconstruction._elbo_loss / _mvae_elbo_loss called
 * on their own, problems.py:401-458, and the per-sample scoring path reduce=False, :415-417, :451-456):
 * logits / target / mask are flat fp32 [n][per_sample] in any (identical) layout; loss_sum[0] += total,
 * per_sample_sum[i] += loss of sample i (nullable), dlogits (nullable, fp32, same layout) =
 * gscale * (sigmoid(x*m) - t*m) * m. */
int mmdyn_bce_logits_flat(const float* logits, const float* target, const float* mask, float* loss_sum,
                          float* per_sample_sum, float* dlogits, float gscale, int n, int per_sample,
                          void* stream);
/* row-wise squared error of [n][d] fp32 matrices: row_sum[i] += mult * sum_j (r-t)^2 */
int mmdyn_mse_rows(const float* recon, const float* target, float* row_sum, float mult, int n, int d,
                   void* stream);
/* MSE 'sum' * multiplier for the pose vectors: loss_sum += mult*sum (r-t)^2; dr = gscale*2*mult*(r-t) */
int mmdyn_mse(const float* recon, const float* target, float* loss_sum, float* drecon, float mult,
              float gscale, int n, void* stream);

/* --- fp32 linear layers of the pose MLP expert (vae.py:14-19, 118-123, 219-222, 282-283) -------
 * y = act(x W^T + b), W [N][K] row-major (torch layout), act: 0 identity, 1 ReLU. */
int mmdyn_linear_f32_fwd(const float* x, const float* W, const float* b, float* y, int M, int N,
                         int K, int ldx, int ldy, int act, void* stream);
/* dx (=|+=) (dy * act'(y)) W ; dW += scale * (dy*act')^T x ; db += scale * colsum(dy*act').
 * dy_act (scratch, [M][N]) receives dy*act'(y); any of dx / dW / db may be NULL;
 * dx_accumulate != 0 adds into dx (several decoders feed the same latent). */
int mmdyn_linear_f32_bwd(const float* x, const float* W, const float* y, const float* dy,
                         float* dy_act, float* dx, float* dW, float* db, int M, int N, int K,
                         int ldx, int ldy, int lddx, int act, int dx_accumulate, float scale,
                         void* stream);

/* --- misc reductions / packing ----------------------------------------------------------------
 * column sums of an fp32 [M][N] matrix (bias gradients): out[n] += scale*sum_m x[m][n] */
int mmdyn_colsum_f32(const float* x, float* out, int M, int N, int ld, float scale, void* stream);
/* same for an fp16 matrix (N multiple of 8, 16-byte aligned rows) */
int mmdyn_colsum_f16(const void* x, float* out, int M, int N, int ld, float scale, void* stream);
/* gather-pack fp32 parameters into an fp16 operand matrix: dst[i] = idx[i] < 0 ? 0 : src[idx[i]] */
int mmdyn_pack_f16(const float* src, const int32_t* idx, void* dst, long long n, void* stream);
/* fp32 gather (packed bias copies): dst[i] = idx[i] < 0 ? 0 : src[idx[i]] */
int mmdyn_gather_f32(const float* src, const int32_t* idx, float* dst, long long n, void* stream);
/* scatter-add a packed fp32 gradient back to parameter layout: dst[idx[i]] += src[i] (idx>=0);
 * each parameter element appears at most once in idx */
int mmdyn_unpack_add_f32(const float* src, const int32_t* idx, float* dst, long long n,
                         void* stream);
/* dst[k] += src[inv[k]] for inv[k] >= 0, k < n: the same scatter driven by the inverse map (every
 * arena element has at most one packed source), so the arena is touched with coalesced accesses */
int mmdyn_gather_add_f32(const float* src, const int32_t* inv, float* dst, long long n, void* stream);
/* dst = fp16(scale * src) */
int mmdyn_f32_to_f16(const float* src, void* dst, long long n, float scale, void* stream);
/* x *= s in place */
int mmdyn_scale_f32(float* x, long long n, float s, void* stream);
/* fp32 NCHW (n,3,H,W) logit gradients -> fp16 NHWC with cp = 4 or 8 channels per pixel (3 used) times
 * scale: cp = 4 is the layout the decoder backward reads (used when the loss is computed outside the
 * library, e.g. by torch autograd) */
int mmdyn_logit_grad_pack(const float* dlogits_nchw, void* out_nhwc, float scale, int n, int H, int W,
                          int pad, int cp, void* stream);

/* --- CVAE conditioning (vae.py:231-237, 286-291) ---------------------------
 * torch.cat((x, c), -1) followed by Linear(K0 + cd, N) = Linear on x (tensor cores, weight columns
 * 0..K0-1) + the rank-cd term c . W[:, K0:]^T, computed here in fp32 with the weight columns
 * addressed in place (row pitch ldw = K0 + cd).
 *   linear_f32_acc   : y[M][N] += x[M][K] . W[N][K]^T                 (pitches ldx, ldw, ldy)
 *   linear_f32_wgrad : dW[N][K] += scale * dy[M][N]^T . x[M][K]       (pitches lddy, ldx, ldw)
 *   cond_add_f16     : raw[r][n'] (fp16) += sum_j c[r][j] * W[n_idx[n']*ldw + col0 + j]   (decoder upsample,
 *                      whose fp16 output columns are permuted: n_idx = torch row of packed column n')
 *   cond_wgrad_f16   : dW[n_idx[n']*ldw + col0 + j] += scale * sum_r g[r][n'] * c[r][j]
 */
int mmdyn_linear_f32_acc(const float* x, const float* W, float* y, int M, int N, int K, int ldx, int ldw, int ldy,
                         void* stream);
int mmdyn_linear_f32_wgrad(const float* x, const float* dy, float* dW, int M, int N, int K, int ldx, int lddy,
                           int ldw, float scale, void* stream);
int mmdyn_cond_add_f16(void* raw, const float* c, const float* W, const int32_t* n_idx, int R, int N, int ldw,
                       int col0, int cd, void* stream);
int mmdyn_cond_wgrad_f16(const void* g, const float* c, float* dW, const int32_t* n_idx, int R, int N, int ldw,
                         int col0, int cd, float scale, void* stream);

/* --- pose MLP on the tensor cores at fp32 accuracy (vae.py:14-19, 118-123, 219-222, 282-283) -----------------
 * The 512-wide Linear layers of the pose expert run on mmdyn_igemm / mmdyn_wgrad with every fp32 operand split into
 * two fp16 terms (x = hi + lo) and the three significant products hi*hi + lo*hi + hi*lo concatenated along the
 * contraction dimension: out[m] = [hi | lo | hi] (mode 0, activations) or [hi | hi | lo] (mode 1: gradients; weights
 * are packed like mode 1 / mode 0 by mmdyn_pack_f16 with index bit 30 marking a low part).  fp32 accumulation in TMEM
 * keeps the result within ~1e-6 of the fp32 product (north_star: 1e-5 for the fp32 parts).
 *   x [M][N] fp32 -> out [M][3N] fp16; relu != 0: x := max(x, 0); mask_y: x := x * (mask_y > 0) (ReLU backward by its
 *   output); x_out (nullable, may alias x): the fp32 value after relu / mask; colsum0 / colsum1 (nullable): column
 *   sums * colsum_scale of that value accumulated into colsum0[n] for n < n_split and colsum1[n - n_split] beyond
 *   (bias gradients of one or two concatenated Linear layers). */
int mmdyn_split_f16(const float* x, const float* mask_y, float* x_out, void* out, int M, int N, int mode, int relu,
                    float* colsum0, float* colsum1, int n_split, float colsum_scale, void* stream);

/* --- Regressor tail (models.py:56-62, 64-77) ---------------------------------
 * out_net = Linear(512(+cd), 256) -> ReLU -> Linear(256, 256) -> ReLU -> Linear(256, out_dim): the first Linear
 * runs on the tensor cores (mmdyn_igemm, fp32 out), the ReLU behind it here, the rest on mmdyn_linear_f32_*.
 *   relu_f32     : y = max(x, 0)                      (n values, in place allowed)
 *   act_grad_f32 : dx = dy * act'(y), act 1 = ReLU judged by its output y, 0 = identity */
int mmdyn_relu_f32(const float* x, float* y, long long n, void* stream);
int mmdyn_act_grad_f32(const float* y, const float* dy, float* dx, int M, int N, int ldy, int act, void* stream);

/* --- device-side input pipeline --------------------------------------------
 * replaces transforms.Compose([Resize(input_size), ToTensor()]) per frame (utils/datasets.py:23-31,
 * 382-392) and the batch assembly of seq_collate_fn (:395-404) for uint8 frames kept in HBM.
 * Bit-identical to Pillow's Image.resize(BILINEAR) (antialiased triangle filter, 22-bit fixed point,
 * horizontal then vertical pass, each rounded to uint8) followed by uint8 / 255.
 *   mmdyn_resize_table_ints : number of int32 entries of the coefficient table for these sizes
 *   mmdyn_resize_table      : fills the table on the HOST (double precision, as precompute_coeffs +
 *                             normalize_coeffs_8bpc); the caller copies it to the device once
 *   mmdyn_frames_u8_to_f32  : out[i] (fp32, 3 x out_h x out_w, planar) = ToTensor(Resize(frames[index[i]]));
 *                             frames: uint8 [N][in_h][in_w][3]; index: int64 device array or NULL (= i)
 */
int mmdyn_resize_table_ints(int in_h, int in_w, int out_h, int out_w);
int mmdyn_resize_table(int in_h, int in_w, int out_h, int out_w, int32_t* table_host, int n_ints);
int mmdyn_frames_u8_to_f32(const void* frames_u8, const long long* index, const int32_t* table_dev,
                           float* out_nchw, int n, int in_h, int in_w, int out_h, int out_w, void* stream);

/* --- fused Adam over a flat fp32 arena (problems.py:138,155; torch.optim.Adam defaults) --------
 * p, g, m, v: [n] fp32.  step_count is the 1-based step AFTER this update (bias correction).
 * g is multiplied by gscale first (e.g. 1/world_size after a sum-allreduce). */
int mmdyn_adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step_count,
                    float gscale, void* stream);
/* same update with the step count read from device memory (*step_dev, already incremented for
 * this update): lets a captured CUDA graph replay the optimizer with correct bias correction */
int mmdyn_adam_flat_devstep(float* p, const float* g, float* m, float* v, long long n, float lr,
                            float beta1, float beta2, float eps, float weight_decay,
                            const uint64_t* step_dev, float gscale, void* stream);
/* SGD with momentum (problems.py:132-136): buf = mom*buf + (g + wd*p); p -= lr*buf */
int mmdyn_sgd_flat(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                   float weight_decay, int first_step, float gscale, void* stream);
/* Guarded forms (what optimizer.step() of problems.py:155 runs here): a gradient entry that is inf / NaN
 * (fp16 overflow in the backward, a poisoned input) leaves its parameter and moments untouched and ORs 1
 * into *nonfinite_flag (device, may be NULL) — read on the host together with the loss, so a step never
 * trains on a non-finite number silently. */
int mmdyn_adam_flat_guarded(float* p, const float* g, float* m, float* v, long long n, float lr,
                            float beta1, float beta2, float eps, float weight_decay,
                            const uint64_t* step_dev, float gscale, unsigned int* nonfinite_flag, void* stream);
int mmdyn_sgd_flat_guarded(float* p, const float* g, float* buf, long long n, float lr, float momentum,
                           float weight_decay, int first_step, float gscale, unsigned int* nonfinite_flag,
                           void* stream);

/* --- data-parallel exchange fused with the optimizer over NVLink peer memory -------------------------
 * (north_star: gradients of the sharded batch are combined across the GPUs of one box; the step it completes is
 * problems.py:150-155.)  One process per GPU; every rank maps the other ranks' gradient arena, parameter arena and
 * flag block (CUDA IPC) and passes the N pointers of each (own rank included) as HOST arrays.
 *   mmdyn_enable_peer_access : cudaDeviceEnablePeerAccess from the current device to `peer_device` (idempotent)
 *   mmdyn_peer_rs_adam_ag    : reduce-scatter + Adam + all-gather in one kernel.  Rank r sums the gradients of its
 *       contiguous shard of the n-float arena over all ranks (peer loads), applies the mmdyn_adam_flat_guarded update
 *       with its local moments, and stores the new parameters into every rank's arena (peer stores).  flag_ptrs[p]:
 *       2*world zero-initialised uint32 of rank p (ready / done epochs); epoch_dev: device counter, incremented by
 *       the caller before every launch on every rank; block_counter: one zeroed uint32.  The kernel returns on a
 *       rank only after all peers finished reading its gradients and writing its parameters. */
/*   mmdyn_ipc_export / mmdyn_ipc_import : CUDA IPC handle (64 bytes) of the allocation containing ptr + the
 *       offset of ptr inside it; import maps it into the CURRENT device address space with peer access enabled
 *       (the mapping stays for the life of the process) and returns the peer pointer. */
int mmdyn_enable_peer_access(int peer_device);
int mmdyn_ipc_export(const void* ptr, void* handle_out /* 64 bytes */, long long* offset_out);
int mmdyn_ipc_import(const void* handle /* 64 bytes */, long long offset, void** ptr_out);
int mmdyn_peer_rs_adam_ag(float* const* grad_ptrs, float* const* param_ptrs, unsigned int* const* flag_ptrs, float* m,
                          float* v, long long n, int rank, int world, float lr, float beta1, float beta2, float eps,
                          float weight_decay, const uint64_t* step_dev, const uint64_t* epoch_dev, float gscale,
                          unsigned int* nonfinite_flag, unsigned int* block_counter, void* stream);

/* --- deterministic device RNG (Philox4x32-10) for eps / dropout masks --------------------------
 * replaces torch.randn (vae.py:58) and nn.Dropout's mask (vae.py:213) on the fast path */
/* the Philox counter of element block i is (*ctr_dev if ctr_dev else 0) + offset + i; keeping the
 * running counter in device memory lets a captured CUDA graph draw fresh numbers at every replay */
int mmdyn_fill_normal(float* out, long long n, uint64_t seed, uint64_t offset,
                      const uint64_t* ctr_dev, void* stream);
int mmdyn_fill_dropout_mask(float* out, long long n, float p_drop, uint64_t seed, uint64_t offset,
                            const uint64_t* ctr_dev, void* stream);
int mmdyn_rng_advance(uint64_t* ctr_dev, uint64_t inc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMDYN_B200_H_ */
