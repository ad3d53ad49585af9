/*
 * mmvae_b200 -- C ABI of the B200-native latent + objective hot path of multimodal-vae-comparison.
 *
 * Boundary contract:
 *   - extern "C" only; plain pointers and sizes; no torch types, no exceptions, no allocation, no host sync,
 *     no global state.  The caller owns every buffer (inputs, outputs, workspaces) and they are DEVICE pointers.
 *   - every call is stream ordered on `stream` (a cudaStream_t passed as void*); cudaSetDevice is the caller's job.
 *   - return value: 0 ok; <0 argument error (MMVAE_E_*); >0 a cudaError_t from the launch.
 *   - row-major contiguous tensors unless a leading dimension (`ld*`, in ELEMENTS) is given.
 *   - "rows" of a reconstruction are k-major: row = k*B + b uses target row b (reference
 *     objectives.py:103-125 reshape_for_loss: target.repeat(K,1,..)).
 *
 * Each entry point cites the reference code it replaces (paths relative to
 * /root/reference/multimodal_compare).  The reference has no FFI: these are the functions a maintainer binds
 * from Python via ctypes (INTEGRATION.md shows the stub), called by the torch.autograd.Functions of the drop-in
 * model/objective plugins.
 */
#ifndef MMVAE_B200_H
#define MMVAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMVAE_ABI_VERSION 1

/* the library is built with -fvisibility=hidden: only the entry points below are exported */
#if defined(__GNUC__)
#define MMVAE_API __attribute__((visibility("default")))
#else
#define MMVAE_API
#endif

/* error codes (negative) */
#define MMVAE_E_ARG   (-1) /* null pointer / non-positive size                     */
#define MMVAE_E_ENUM  (-2) /* unknown dtype / ltype / dist enum                    */
#define MMVAE_E_LIMIT (-3) /* size beyond a compiled-in limit (M, D, smem)         */

/* element types of big streamed tensors (reconstructions, targets, their gradients) */
enum { MMVAE_F32 = 0, MMVAE_BF16 = 1 };
/* posterior / likelihood families (reference vae.py:142-147 dist_map) */
enum { MMVAE_NORMAL = 0, MMVAE_LAPLACE = 1 };
/* element-wise likelihood terms (reference objectives.py:389-458 ReconLoss.{bce,lprob,mse,l1}) */
enum { MMVAE_LT_BCE = 0, MMVAE_LT_LPROB_NORMAL = 1, MMVAE_LT_LPROB_LAPLACE = 2, MMVAE_LT_MSE = 3, MMVAE_LT_L1 = 4,
       /* BCE on decoder LOGITS with the reference decoder tail fused in (decoders.py:96-97):
          x = clamp(sigmoid(y), 1e-6, 1-1e-6); saves one read+write of the (K*B, P) reconstruction per direction */
       MMVAE_LT_BCE_LOGITS = 5,
       /* lprob with padding masks: the reference overwrites the likelihood's scale with its (cropped) loc
          (objectives.py:43-45), i.e. log_prob of Normal / Laplace(loc = x, scale = x); NaN (x < 0) -> 0.  Reproduced. */
       MMVAE_LT_LPROB_NORMAL_SELF = 6, MMVAE_LT_LPROB_LAPLACE_SELF = 7 };

/* peer-memory all-reduce (mmvae_*_peer entry points): layout constants of the per-rank symmetric buffer */
#define MMVAE_ELBO_MAX_TERMS 48     /* likelihood terms / KL segments of one mmvae_objective_elbo call */
#define MMVAE_ELBO_MAX_CTAS 64      /* grid limit of mmvae_objective_elbo (size of its partials scratch / 2) */
#define MMVAE_PEER_CHANNELS 8        /* independent call sites / streams                       */
#define MMVAE_PEER_MAX_WORLD 32      /* ranks of one NVLink domain                             */
#define MMVAE_PEER_BUFFER_BYTES (72 * 1024) /* bytes to allocate (zeroed) per rank             */

#define MMVAE_MAX_MODS 8    /* modalities per model                       */
#define MMVAE_MAX_COLS 256  /* latent columns per modality (shared+private) */

MMVAE_API int mmvae_version(void);

/* ---------------------------------------------------------------------------------------------------------
 * Likelihood row reductions: replaces  (obj_fn.recon_loss_fn(px_z, x, K) * llik_scaling).sum(-1)
 *   reference: objectives.py:30-52 (recon_loss_fn), :103-125 (reshape_for_loss), :389-458 (ReconLoss.*),
 *   call sites mmvae_models.py:48-49, :54-55, :66-71, :177, :312-313, :448-449, :454.
 *
 *   out_rows[r] = lam * sum_p logp(recon[r,p] | target[r % B, p])           r in [0, rows), rows = K*B
 *     BCE           t*max(log x,-100) + (1-t)*max(log(1-x),-100)      (F.binary_cross_entropy clamps)
 *     LPROB_NORMAL  Normal(x, scale).log_prob(t), NaN -> 0            (objectives.py:422-423)
 *     LPROB_LAPLACE Laplace(x, scale).log_prob(t), NaN -> 0
 *     MSE           -(x-t)^2          L1   -|x-t|
 *     BCE_LOGITS    BCE of clamp(sigmoid(recon), 1e-6, 1-1e-6): recon holds logits, gradient is w.r.t. the logits
 *
 * _fwd   : reads recon + target, writes out_rows.                                   bytes: R + T
 * _bwd   : grad[r,p] = w_rows[r] * lam * dlogp/dx                                  bytes: 2R + T (+rows)
 * _fused : one pass producing out_rows AND grad for weights known a priori (every ELBO): w = w_rows[r] if
 *          w_rows != NULL else w_const.                                            bytes: 2R + T
 *
 * workspace: mmvae_loglik_workspace_bytes(rows, P, dtype) bytes (may be 0), used for a deterministic two stage
 * row sum when a row is split over several CTAs (few rows, long rows).
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int64_t mmvae_loglik_workspace_bytes(int64_t rows, int64_t P, int dtype_recon);

MMVAE_API int mmvae_loglik_rowreduce_fwd(const void* recon, int64_t ld_recon, int dtype_recon,
                               const void* target, int64_t ld_target, int dtype_target,
                               int64_t rows, int64_t B, int64_t P, int ltype, float scale, float lam,
                               float* out_rows, void* workspace, void* stream);

MMVAE_API int mmvae_loglik_rowreduce_bwd(const void* recon, int64_t ld_recon, int dtype_recon,
                               const void* target, int64_t ld_target, int dtype_target,
                               int64_t rows, int64_t B, int64_t P, int ltype, float scale, float lam,
                               const float* w_rows, void* grad_recon, int64_t ld_grad, void* stream);

MMVAE_API int mmvae_loglik_rowreduce_fused(const void* recon, int64_t ld_recon, int dtype_recon,
                                 const void* target, int64_t ld_target, int dtype_target,
                                 int64_t rows, int64_t B, int64_t P, int ltype, float scale, float lam,
                                 const float* w_rows, float w_const,
                                 float* out_rows, void* grad_recon, int64_t ld_grad,
                                 void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * category_ce rows: replaces ReconLoss.category_ce (objectives.py:485-500): nn.CrossEntropyLoss with
 * probability targets -> the class axis is dim 1.  recon (rows, C, d) (d = 1 for 2-D inputs), target (B, C, d):
 *   out_rows[r] = lam * sum_j sum_c t[c,j] * (x[c,j] - logsumexp_c x[:,j])
 *   grad[r,c,j] = w * lam * (t[c,j] - softmax_c(x[:,j])[c] * sum_c' t[c',j])
 * ld_recon / ld_grad: row stride in elements (>= C*d; a mask crop loc[:, :T] keeps the row stride,
 * objectives.py:43-45).  mode: 0 fwd, 1 bwd, 2 fused (same meaning as above).
 * stats: optional (rows, 2, d) floats: the forward stores logsumexp_c and sum_c t per column, the backward then
 * runs as a single streaming pass (no reductions); NULL = recompute in the backward.
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int mmvae_catce_rows(int mode, const void* recon, int64_t ld_recon, int dtype_recon,
                     const void* target, int64_t ld_target, int dtype_target,
                     int64_t rows, int64_t B, int64_t C, int64_t d, float lam,
                     const float* w_rows, float w_const,
                     float* out_rows, void* grad_recon, int64_t ld_grad, float* stats, void* stream);
/* same with the text decoder's tail fused in (reference decoders.py:722, "zero for padded area": output * mask, SURVEY
 * 8f rank 1): recon is the UNMASKED decoder output, mask (B, C) bytes (row r uses mask row r % B, non-zero = keep,
 * ld_mask >= C); the loss is evaluated on x_eff = mask ? x : 0 and grad_recon is the gradient with respect to the
 * unmasked tensor (zero where masked).  mask == NULL: identical to mmvae_catce_rows.  With a mask the backward (mode 1)
 * recomputes the column statistics (stats is written by mode 0 but not read). */
MMVAE_API int mmvae_catce_rows_masked(int mode, const void* recon, int64_t ld_recon, int dtype_recon,
                            const void* target, int64_t ld_target, int dtype_target,
                            int64_t rows, int64_t B, int64_t C, int64_t d, float lam,
                            const float* w_rows, float w_const, float* out_rows, void* grad_recon, int64_t ld_grad,
                            float* stats, const unsigned char* mask, int64_t ld_mask, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * optimal_sigma (sigma-VAE) rows: replaces ReconLoss.optimal_sigma (objectives.py:502-509) + utils.softclip
 * (utils.py:66-69).  Three stream-ordered stages so that a multi-GPU caller can all-reduce the scalar between
 * stage 1 and 2:
 *   _sumsq : sumsq[0] += sum_all (t - x)^2, sumsq[1] += rows*P   (TWO doubles, caller zeroes them; a sharded caller
 *            all-reduces both in one collective).  _fwd / _bwd take the element count from n_total, or from sumsq[1]
 *            when n_total <= 0 (device-resident global count: uneven shards).  row_sumsq (rows floats, may be NULL)
 *            receives sum_p (t - x)^2 of every row.
 *   _fwd   : log_sigma = -6 + softplus(log sqrt(sumsq/n_total) + 6);
 *            out_rows[r] = -lam * sum_p [ ((t-x)/sigma)^2 + log_sigma + 0.5 log 2pi ]; stats = {log_sigma, dlogsigma/du}.
 *            With row_sumsq != NULL (the array stage 1 wrote) the rows follow from it without a pass over recon / target.
 *   _bwd   : only log_sigma carries gradient (the squared term is detached):
 *            grad[i] = -lam * P * (sum_r w_rows[r]) * sigmoid(u+6) * (x_i - t_i) / sumsq
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int mmvae_osigma_sumsq(const void* recon, int64_t ld_recon, int dtype_recon,
                       const void* target, int64_t ld_target, int dtype_target,
                       int64_t rows, int64_t B, int64_t P, double* sumsq, float* row_sumsq, void* stream);
MMVAE_API int mmvae_osigma_fwd(const void* recon, int64_t ld_recon, int dtype_recon,
                     const void* target, int64_t ld_target, int dtype_target,
                     int64_t rows, int64_t B, int64_t P, float lam, const double* sumsq, double n_total,
                     float* out_rows, float* stats2, const float* row_sumsq, void* stream);
MMVAE_API int mmvae_osigma_bwd(const void* recon, int64_t ld_recon, int dtype_recon,
                     const void* target, int64_t ld_target, int dtype_target,
                     int64_t rows, int64_t B, int64_t P, float lam, const double* sumsq, double n_total,
                     const float* w_rows, float* wsum_scratch, void* grad_recon, int64_t ld_grad, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Latent draws: product-of-experts fusion / direct posteriors, reparameterised sampling and KL rows.
 * Replaces TorchMMVAE.product_of_experts (mmvae_base.py:203-222), POE.modality_mixing + prior_expert
 * (mmvae_models.py:210-250), MoPOE.poe_fusion + mixture_component_selection (:385-410), the DMVAE joint
 * (:478-480), every Normal/Laplace rsample on the path (:99, :201, :366, :481-499) and calc_kld ->
 * torch.distributions.kl (objectives.py:148-161, utils.py:399-405).
 *
 * Encoder outputs: mu, s as (M, B, Dtot) fp32.  A draw j is described by mmvae_draw_desc:
 *   fused (default): experts = modalities in `mask` (+ the prior expert (0,0) if MMVAE_DRAW_PRIOR) on columns
 *          [col0, col0+width):  var_e = exp(s_e)+1e-8, T_e = 1/var_e, loc = sum mu_e T_e / sum T_e,
 *          scale = 1/sum T_e   (the PoE *variance* is the Normal scale -- reference quirk, reproduced)
 *   MMVAE_DRAW_DIRECT: loc, scale = mu_m, s_m of the single modality in `mask`
 *   MMVAE_DRAW_ROWMASK: the expert set of row b is row_masks[b] (bit 31 = prior expert) -- MoPoE mixture
 *          component selection as a row -> subset bitmask map
 *   z[k,b,c]  = loc + scale * eps[k,b,c]          (Normal)          k < K  (K may be 0: no samples)
 *             = loc - scale * sign(u) log1p(-|u|) (MMVAE_DRAW_LAPLACE, eps holds u)
 *   kl[b]     = sum_c KL(q || prior): kl_mode 0 none, 1 vs learnable prior N(mu0, s0), 2 vs N(0,1);
 *               q is Normal, or Laplace with MMVAE_DRAW_LAPLACE (torch kl.py _kl_laplace_normal)
 * Offsets are in ELEMENTS into the caller's packed eps / z / par_loc / par_scale / kl buffers; -1 = not wanted.
 * ------------------------------------------------------------------------------------------------------- */
#define MMVAE_DRAW_PRIOR   1
#define MMVAE_DRAW_DIRECT  2
#define MMVAE_DRAW_LAPLACE 4
#define MMVAE_DRAW_ROWMASK 8

typedef struct {
    uint32_t mask;
    int32_t flags;
    int32_t kl_mode;
    int32_t col0, width;
    int32_t K;
    int64_t eps_off; /* (K,B,width) noise                      */
    int64_t z_off;   /* (K,B,width) samples                    */
    int64_t par_off; /* (B,width) loc and scale outputs, or -1 */
    int64_t kl_off;  /* (B) KL rows, or -1                     */
} mmvae_draw_desc;

#define MMVAE_MAX_DRAWS 64

MMVAE_API int mmvae_latent_draws_fwd(const float* mu, const float* s, int M, int64_t B, int Dtot,
                           const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                           const float* mu0, const float* s0, /* (D) learnable prior, may be NULL if unused */
                           const float* eps, float* z, float* par_loc, float* par_scale, float* kl,
                           void* stream);

/* backward: dz (packed like z), dkl (packed like kl: per-row upstream grads), dpar_loc / dpar_scale (packed
 * like par_*; may be NULL) -> dmu, ds (M,B,Dtot) fully written; dprior_ws: (2, grid, Dtot) partials followed by
 * dmu0 (D), ds0 (D) written by the finalisation kernel; size from mmvae_latent_draws_bwd_ws_floats(). */
MMVAE_API int64_t mmvae_latent_draws_bwd_ws_floats(int64_t B, int Dtot);
MMVAE_API int mmvae_latent_draws_bwd(const float* mu, const float* s, int M, int64_t B, int Dtot,
                           const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                           const float* mu0, const float* s0,
                           const float* eps, const float* dz, const float* dkl,
                           const float* dpar_loc, const float* dpar_scale,
                           float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream);
/* Latent draws with the encoder tail fused in (reference encoders.py:49-54): `s` holds the RAW output
 * of the encoder's second Linear head (M,B,Dtot); the kernels evaluate s = softmax(raw, -1) + 1e-6 per row themselves.
 * enc_tail == 0: identical to mmvae_latent_draws_{fwd,bwd}.  s_out (may be NULL): (M,B,Dtot) copy of the scales.  The
 * backward writes into `ds` the gradient with respect to the raw logits: d/draw = p (d/ds - <d/ds, p>), p = s - 1e-6. */
MMVAE_API int mmvae_latent_draws_fwd_tail(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                const float* mu0, const float* s0, const float* eps, float* z, float* par_loc,
                                float* par_scale, float* kl, int enc_tail, float* s_out, void* stream);
MMVAE_API int mmvae_latent_draws_bwd_tail(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                const float* mu0, const float* s0, const float* eps, const float* dz,
                                const float* dkl, const float* dpar_loc, const float* dpar_scale, int enc_tail,
                                float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream);

/* Element-wise KL(q || N(loc0, scale0)): the (n, D) tensor reference calc_kld returns (objectives.py:148-161,
 * utils.py:399-405) and the per-dimension KL tables of the analysis hooks use (utils.py:130-162).  q is Normal or
 * Laplace (dist), the prior row (D) is broadcast over n.  The backward's prior gradient is reduced over n in two
 * stages (shared-memory accumulation inside a CTA, ordered sum across CTAs through ws). */
MMVAE_API int mmvae_kl_elementwise_fwd(const float* loc, const float* scale, const float* loc0, const float* scale0,
                             int dist, int64_t n, int D, float* out, void* stream);
MMVAE_API int64_t mmvae_kl_elementwise_ws_floats(int64_t n, int D);
MMVAE_API int mmvae_kl_elementwise_bwd(const float* loc, const float* scale, const float* loc0, const float* scale0,
                             int dist, int64_t n, int D, const float* upstream, float* dloc, float* dscale,
                             float* ws, float* dloc0, float* dscale0, void* stream);

/* Per-dimension KL tables of the analysis hooks: replaces utils.make_kl_df (utils.py:130-162; called by
 * trainer.analyse_data trainer.py:242-272), which moves every posterior to the CPU and evaluates M + 2*C(M,2) separate
 * torch kl_divergence calls.  loc, scale (M, n, D); prior row (D); out (T, n, D), T = M + M(M-1)/2:
 *   out[i] = KL(q_i || N(loc0, scale0)), then for pairs i < j (itertools.combinations order)
 *   out[M + pair] = 0.5 (KL(q_i || q_j) + KL(q_j || q_i)).  All M posteriors share one family (Normal or Laplace). */
MMVAE_API int mmvae_kl_table(const float* loc, const float* scale, int M, const int32_t* dists_host, const float* loc0,
                   const float* scale0, int64_t n, int D, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * MoE sample + log-densities (fused): replaces MOE.forward's rsample (mmvae_models.py:99), the importance
 * terms of the ELBO branch (:56-62) and the log p(z) / log-mean q_j(z) terms of MultimodalObjective.iwae /
 * _m_dreg_looser (objectives.py:342-373).
 *   z[r,k,b,:]   = rsample of q_r = dist_r(mu_r, s_r) with noise eps[r,k,b,:]
 *   lq[r,j,k,b]  = sum_d log q_j(z[r,k,b,d])                      (M,M,K,B)
 *   lpz[r,k,b]   = sum_d log N(z[r,k,b,d]; mu0_d, s0_d)           (M,K,B)
 * backward: given dz_ext (grad reaching z from the decoders; may be NULL), dlq, dlpz ->
 *   dmu, ds (M,B,D) and per-CTA partials of dmu0, ds0.  `through_z` = 0 reproduces the ELBO branch where z is
 *   detached inside the log-densities (:58) -- dlq then only reaches the parameters of q_j.
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int mmvae_moe_logdens_fwd(const float* mu, const float* s, int M, int64_t B, int D, int K, const int32_t* dists_host,
                          const float* mu0, const float* s0, const float* eps,
                          float* z, float* lq, float* lpz, void* stream);
/* same with the encoder tail fused in (reference encoders.py:49-54): s_raw (M,B,D) holds the RAW output
 * of the encoder's second Linear head, the kernel evaluates s = softmax(s_raw, -1) + 1e-6 itself; s_out (M,B,D, may be
 * NULL) receives the scales (for the torch.distributions objects the plugin API hands out).  Needs the flat kernels:
 * D % 4 == 0, D <= 128, M <= 3, 16-byte aligned buffers (MMVAE_E_LIMIT otherwise). */
MMVAE_API int mmvae_moe_logdens_fwd_tail(const float* mu, const float* s_raw, int M, int64_t B, int D, int K,
                               const int32_t* dists_host, const float* mu0, const float* s0, const float* eps,
                               float* z, float* lq, float* lpz, float* s_out, void* stream);
MMVAE_API int64_t mmvae_moe_logdens_bwd_ws_floats(int64_t B, int D, int K);
MMVAE_API int mmvae_moe_logdens_bwd(const float* mu, const float* s, int M, int64_t B, int D, int K, const int32_t* dists_host,
                          const float* mu0, const float* s0, const float* eps,
                          const float* dz_ext, const float* dlq, const float* dlpz, int through_z,
                          float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream);
/* same, with the per-(r,k) weights of a DReG-style objective folded into the coefficients instead of being
 * materialised by the caller as (M,M,K,B) / (M,K,B) gradient tensors (objectives.py:361-387: the weights are a
 * softmax over K of batch-summed log-weights, constant over b):
 *   rkc[r,k] = rk_mul * (rk_scale_dev ? *rk_scale_dev : 1) * rk_w[r,k]
 *   coefficient of log q_j(z[r,k,b]) = rkc[r,k] * dlq[r,j,k,b]   (dlq holds softmax_j(lq), required)
 *   coefficient of log p(z[r,k,b])   = -rkc[r,k]                 (dlpz must be NULL)
 * rk_w == NULL: identical to mmvae_moe_logdens_bwd.  Needs D % 4 == 0, D <= 128, M <= 3 and 16-byte aligned buffers (MMVAE_E_LIMIT otherwise).
 * dlq_packed != 0 (rk mode, M == 2, dz_ext given, through_z set): dlq is the (K, B, 4) layout [r*2 + j] written by
 * mmvae_objective_dreg_stage1_ptrs(soft_packed = 1). */
MMVAE_API int mmvae_moe_logdens_bwd_rk(const float* mu, const float* s, int M, int64_t B, int D, int K, const int32_t* dists_host,
                             const float* mu0, const float* s0, const float* eps,
                             const float* dz_ext, const float* dlq, const float* dlpz, int through_z,
                             const float* rk_w, const float* rk_scale_dev, float rk_mul, int dlq_packed,
                             float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream);
/* same; enc_tail != 0: `s` holds the raw logits of mmvae_moe_logdens_fwd_tail and `ds` receives the gradient with
 * respect to THEM: d/draw = p (d/ds - <d/ds, p>), p = softmax(raw) (backward through the fused encoder tail). */
MMVAE_API int mmvae_moe_logdens_bwd_tail(const float* mu, const float* s, int M, int64_t B, int D, int K, const int32_t* dists_host,
                               const float* mu0, const float* s0, const float* eps,
                               const float* dz_ext, const float* dlq, const float* dlpz, int through_z,
                               const float* rk_w, const float* rk_scale_dev, float rk_mul, int dlq_packed, int enc_tail,
                               float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Objective combination (forward value + the per-row weights its backward needs, in one launch):
 *   _iwae : MultimodalObjective.iwae objectives.py:342-359 (N1 shim) + utils.log_mean_exp utils.py:395-396
 *           lw[r,k,b] = lpz + sum_l lpx[r,l,k,b] - beta * logmeanexp_j lq[r,j,k,b]
 *           loss_b[b] = -(logsumexp_{r,k} lw - log(M K));   w[r,k,b] = softmax_{r,k}(lw) (= -dloss/dlw)
 *           dlq[r,j,k,b] = beta * w * softmax_j(lq[r,:,k,b])
 *   _dreg : _m_dreg_looser + dreg objectives.py:361-387 (parity mode: batch-summed log-weights)
 *           stage 1: lw_part[r,k] = sum_b (lpz + sum_l lpx - logmeanexp_j lq)   (local batch shard)
 *           [multi-GPU: all-reduce lw_part between the stages, SURVEY 8e (1)]
 *           stage 2: wt = softmax_k(lw[r,:]); loss = -(1/M) sum_{r,k} wt*lw;  out wt (M,K)
 *   _sum  : deterministic single-CTA sum of n floats (ELBO reductions objectives.py:54-67).
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int mmvae_objective_iwae(const float* lpz, const float* lq, const float* lpx, int M, int L, int K, int64_t B,
                         float beta, float* lw, float* loss_b, float* w, float* dlq, void* stream);
/* same, the M*L likelihood row vectors (K*B each) addressed through a HOST array of device pointers
 * (index r*L + l) instead of one stacked (M,L,K,B) tensor: no concatenation copy.  lpx may be NULL then. */
MMVAE_API int mmvae_objective_iwae_ptrs(const float* lpz, const float* lq, const float* lpx,
                              const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B, float beta,
                              float* lw, float* loss_b, float* w, float* dlq, void* stream);
/* same, plus what a backward with a UNIT upstream gradient needs and the batch sum of the loss, so that the serial
 * section between the forward and the backward kernels of a step is one launch instead of four:
 *   dlpz_unit (M,K,B) = -w (gradient of lpz and of every likelihood row vector of modality r for g == 1; dlq already
 *   holds the g == 1 gradient of lq); loss_sum = sum_b loss_b, summed in a fixed order by the CTA that finishes last
 *   (ticket: one zero-initialised unsigned int owned by the caller; it is zero again when the kernel ends).
 * dlpz_unit, loss_sum and ticket may be NULL (loss_sum and ticket only together). */
MMVAE_API int mmvae_objective_iwae_fused(const float* lpz, const float* lq, const float* lpx,
                               const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B, float beta,
                               float* lw, float* loss_b, float* w, float* dlq, float* dlpz_unit, float* loss_sum,
                               unsigned int* ticket, void* stream);
/* IWAE backward in one launch: dlpz_out = -g*w (n_w floats; also the gradient of each likelihood row vector of the
 * same modality), dlq_inout *= g (n_dlq floats); g_dev: device scalar (upstream gradient of the loss). */
MMVAE_API int mmvae_objective_iwae_bwd(const float* g_dev, const float* w, float* dlq_inout, float* dlpz_out,
                             int64_t n_w, int64_t n_dlq, void* stream);
/* learnable prior scale s0 = softmax(logits)*D (reference mmvae_models.py:28-30 pz_params) and its backward
 * dlogits = D p (ds0 - <ds0, p>); replaces four tiny eager kernels per step. */
MMVAE_API int mmvae_prior_scale_fwd(const float* logits, int D, float* s0, void* stream);
MMVAE_API int mmvae_prior_scale_bwd(const float* s0, const float* ds0, int D, float* dlogits, void* stream);
/* lw_part: (MMVAE_DREG_MAX_SPLIT + 1) * M * K DOUBLES; the first M*K receive the local batch sums, the rest is
 * scratch for the deterministic two-stage batch reduction.  The sums are carried in fp64: they feed a softmax over
 * K whose conditioning is set by their absolute error, and the reference holds them in fp64 for lprob likelihoods
 * (objectives.py:422).  lq_soft (M,M,K,B) = softmax_j(lq), may be NULL. */
#define MMVAE_DREG_MAX_SPLIT 64
MMVAE_API int mmvae_objective_dreg_stage1(const float* lpz, const float* lq, const float* lpx, int M, int L, int K, int64_t B,
                                double* lw_part, float* lq_soft, void* stream);
/* same, likelihood rows through a HOST array of M*L device pointers (index r*L + l, (K*B) each): no stack copy.
 * soft_packed != 0 (M == 2 only): lq_soft is written as (K, B, 4) vectors [r*2 + j] -- the layout
 * mmvae_moe_logdens_bwd_rk reads with one 16-byte copy per (k, b) when its dlq_packed flag is set. */
MMVAE_API int mmvae_objective_dreg_stage1_ptrs(const float* lpz, const float* lq, const float* lpx,
                                     const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B,
                                     double* lw_part, float* lq_soft, int soft_packed, void* stream);
MMVAE_API int mmvae_objective_dreg_stage2(const double* lw, int M, int K, float* wt, float* loss, void* stream);
/* DReG backward for the (M,L,K,B) likelihood rows (and log p(z)): d_rows[r,l,k,b] = -(g/M) * wt[r,k], g = *g_dev
 * (upstream gradient of the loss; NULL = 1).  The log q gradients are folded into mmvae_moe_logdens_bwd_rk. */
MMVAE_API int mmvae_objective_dreg_rowgrads(const float* g_dev, const float* wt, int M, int L, int K, int64_t B,
                                  float* d_rows, void* stream);
/* ELBO combination -- replaces reference objectives.py:54-67 (BaseObjective.elbo), :316-340 (calculate_loss) and the
 * model-level sums of mmvae_models.py:181-187 (POE.objective), :314-320 (MoPOE.objective), :455-465 (DMVAE.objective):
 *   loss = sum_{i<n_terms} coef[i] * sum_{r<n[i]} term_i[r]  +  sum_{j<n_kl} kl_coef[j] * sum_{b<B} kl[j*B + b]
 *   kld  = sum_j kl_log_coef[j] * sum_b kl[j*B + b]              (the logged "kld"; kld and kl_log_coef may be NULL)
 *   dkl_unit[j*B + b] = kl_coef[j]   (gradient of the packed KL rows for a unit upstream gradient; may be NULL)
 * term_ptrs_host / term_n_host / term_coef_host / kl_coef_host / kl_log_coef_host are HOST arrays (copied into the
 * kernel parameters; at most MMVAE_ELBO_MAX_TERMS entries each); term_i: device fp32 vectors (likelihood row vectors,
 * or already reduced scalars with n = 1); kl: the packed (n_kl, B) KL rows of mmvae_latent_draws_fwd.  One launch of
 * up to MMVAE_ELBO_MAX_CTAS CTAs; ws (2 * MMVAE_ELBO_MAX_CTAS floats) and ticket (one zero-initialised unsigned int,
 * zero again when the kernel ends) let the last CTA add the per-CTA partials in CTA order (deterministic); with
 * ws == NULL or ticket == NULL a single CTA does everything. */
MMVAE_API int mmvae_objective_elbo(const float* const* term_ptrs_host, const int64_t* term_n_host,
                         const float* term_coef_host, int n_terms, const float* kl, int64_t B,
                         const float* kl_coef_host, const float* kl_log_coef_host, int n_kl, float* loss, float* kld,
                         float* dkl_unit, float* ws, unsigned int* ticket, void* stream);
MMVAE_API int mmvae_reduce_sum(const float* x, int64_t n, float scale, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused compute + collective over NVLink peer memory (multi-GPU): the three reductions of this path whose
 * result every rank needs run as ONE kernel each -- local math, exchange through peer-mapped symmetric buffers
 * (P2P stores / loads over NVSwitch, flags with system-scope release / acquire), the math that consumes the sum --
 * instead of kernel -> NCCL all-reduce -> kernel.  Sums are taken in rank order: bit-identical on every rank.
 *   peer_bufs_dev: DEVICE array of `world` pointers, entry q = rank q's symmetric buffer of MMVAE_PEER_BUFFER_BYTES
 *   bytes (zero-initialised once, before the first call, followed by a barrier) mapped into this rank's address
 *   space; `channel` < MMVAE_PEER_CHANNELS separates call sites that may be in flight at the same time.  Every rank
 *   must issue the same calls in the same order per channel.  CUDA-graph capturable (sequence numbers live in the
 *   buffer).  A wait that exceeds ~2 s sets the error word (mmvae_peer_error_offset()) instead of hanging.
 *   _prior_scale_bwd_peer : dlogits = all-reduce( D p (ds0 - <ds0, p>) )   (mmvae_prior_scale_bwd + gradient sync of
 *                           the replicated prior logits _pz_params[1], reference mmvae_models.py:28-30)
 *   _dreg_stage2_peer     : lw <- all-reduce(lw) (global batch sums, objectives.py:361-387), then stage 2 as above;
 *                           lw is updated in place with the global sums
 *   _peer_allreduce_f64   : in-place sum of n <= 512 doubles (optimal_sigma sum of squares + element count)
 * ------------------------------------------------------------------------------------------------------- */
MMVAE_API int64_t mmvae_peer_error_offset(void);
MMVAE_API int mmvae_prior_scale_bwd_peer(const float* s0, const float* ds0, int D, float* dlogits,
                               void* const* peer_bufs_dev, int rank, int world, int channel, void* stream);
MMVAE_API int mmvae_objective_dreg_stage2_peer(double* lw_inout, int M, int K, float* wt, float* loss,
                                     void* const* peer_bufs_dev, int rank, int world, int channel, void* stream);
MMVAE_API int mmvae_peer_allreduce_f64(double* data_inout, int n, void* const* peer_bufs_dev, int rank, int world,
                             int channel, void* stream);

/* in-place scale of a gradient buffer by a DEVICE scalar, skipped entirely (early exit) when the scalar == 1:
 * lets the fused ELBO path keep autograd semantics for loss.backward(gradient=g) at zero cost when g == 1. */
MMVAE_API int mmvae_scale_inplace(void* buf, int dtype, int64_t n, const float* scalar_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMVAE_B200_H */
