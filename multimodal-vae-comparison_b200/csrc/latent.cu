// Latent draws: precision-weighted product-of-experts fusion (or a direct posterior), reparameterised K-sample
// draws and KL rows for an arbitrary list of "draw descriptors" in ONE launch -- all 2^M-1 PoE subsets of MVAE,
// the MoPoE joint (row -> subset bitmask map), the DMVAE joint / shared / private / cross draws.
//
// Replaces reference mmvae_base.py:203-222 (product_of_experts), mmvae_models.py:210-250 (POE.modality_mixing,
// prior_expert), :385-410 (MoPOE.poe_fusion, mixture_component_selection), :478-499 (DMVAE draws), the
// Normal/Laplace rsample calls and objectives.py:148-161 -> torch.distributions.kl closed forms.
//
// Mapping: one warp per batch row (grid-stride), lanes stride over the latent columns so every global access is a
// contiguous <=128 B segment; KL row sums are warp-shuffle reductions.  The backward stages the per-row
// (M, Dtot) gradient accumulators in shared memory (lane-private columns) and reduces the learnable-prior gradient
// over the batch in two deterministic stages (per-CTA partials -> finalisation kernel), no atomics on HBM.
#include "common.cuh"

namespace mmvae {

constexpr int kWarps = 8;

struct DrawsParams {
    const float *mu, *s;
    const uint32_t* row_masks;
    const float *mu0, *s0, *eps;
    float *z, *ploc, *pscale, *kl;            // forward outputs
    const float *dz, *dkl, *dploc, *dpscale;  // backward inputs
    float *dmu, *ds, *ws;                     // backward outputs
    int64_t B;
    int M, Dtot, n;
    // encoder tail fused in (reference encoders.py:49-54): `s` holds the raw logits of the encoder's second head, a warp
    // evaluates s = softmax(raw, -1) + 1e-6 for its row into shared memory before it runs the draws; s_out: optional
    // (M,B,Dtot) copy of the scales; the backward returns d/draw = p (d/ds - <d/ds, p>), p = s - 1e-6
    int enc_tail;
    float* s_out;
    mmvae_draw_desc d[MMVAE_MAX_DRAWS];
};
static_assert(sizeof(DrawsParams) <= 4050, "kernel parameter block too large");

__device__ __forceinline__ float noise_eff(float e, bool laplace) {
    // Normal: z = loc + scale*eps.  Laplace (torch laplace.py:73-84): z = loc - scale*sign(u)*log1p(-|u|)
    if (!laplace) return e;
    const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
    return -sg * log1pf(-fabsf(e));
}

__device__ __forceinline__ void desc_mask(const DrawsParams& p, const mmvae_draw_desc& d, int64_t b, uint32_t& mask,
                                          bool& prior) {
    if (d.flags & MMVAE_DRAW_ROWMASK) {
        const uint32_t rm = __ldg(p.row_masks + b);
        mask = rm & 0x7fffffffu;
        prior = (rm >> 31) != 0;
    } else {
        mask = d.mask;
        prior = (d.flags & MMVAE_DRAW_PRIOR) != 0;
    }
}

constexpr float kEncEta = 1e-6f;  // reference utils.Constants.eta

// Fused encoder tail: s_row[m*Dtot + c] = softmax(raw[m, b, :])[c] + eta for the warp's row b (lanes stride over the
// columns, two warp reductions per modality).
__device__ __forceinline__ void stage_scales(const DrawsParams& p, int64_t b, float* s_row, int lane) {
    for (int m = 0; m < p.M; ++m) {
        const float* raw = p.s + ((int64_t)m * p.B + b) * p.Dtot;
        float mx = -INFINITY;
        for (int c = lane; c < p.Dtot; c += 32) mx = fmaxf(mx, __ldg(raw + c));
        mx = warp_max(mx);
        float se = 0.f;
        for (int c = lane; c < p.Dtot; c += 32) {
            const float e = expf(__ldg(raw + c) - mx);
            s_row[m * p.Dtot + c] = e;
            se += e;
        }
        se = warp_sum(se);
        const float inv = 1.0f / se;
        for (int c = lane; c < p.Dtot; c += 32) {
            const float v = s_row[m * p.Dtot + c] * inv + kEncEta;
            s_row[m * p.Dtot + c] = v;
            if (p.s_out) p.s_out[((int64_t)m * p.B + b) * p.Dtot + c] = v;
        }
    }
    __syncwarp();
}

// loc / scale of one column; Tsum returned for the backward.  s_row: the row's scales in shared memory (fused
// encoder tail) or NULL (scales read from p.s)
__device__ __forceinline__ void fuse_column(const DrawsParams& p, const mmvae_draw_desc& d, uint32_t mask, bool prior,
                                            int64_t b, int c, float& loc, float& scale, float& Tsum,
                                            const float* s_row = nullptr) {
    if (d.flags & MMVAE_DRAW_DIRECT) {
        const int m = __ffs(mask) - 1;
        const int64_t o = ((int64_t)m * p.B + b) * p.Dtot + c;
        loc = __ldg(p.mu + o);
        scale = s_row ? s_row[m * p.Dtot + c] : __ldg(p.s + o);
        Tsum = 0.f;
        return;
    }
    // prior expert: mu = 0, logvar = 0 -> var = exp(0) + 1e-8, T = 1/var
    float ts = prior ? 1.0f / (1.0f + 1e-8f) : 0.f;
    float num = 0.f;
    for (int m = 0; m < p.M; ++m) {
        if (!((mask >> m) & 1u)) continue;
        const int64_t o = ((int64_t)m * p.B + b) * p.Dtot + c;
        const float T = 1.0f / (expf(s_row ? s_row[m * p.Dtot + c] : __ldg(p.s + o)) + 1e-8f);
        ts += T;
        num += __ldg(p.mu + o) * T;
    }
    loc = num / ts;
    scale = 1.0f / ts;  // the PoE variance, used as the Normal scale (reference quirk)
    Tsum = ts;
}

__device__ __forceinline__ float kl_value(bool laplace, float l, float sg, float m0, float s0) {
    if (!laplace) {  // torch kl.py _kl_normal_normal
        const float r = sg / s0;
        const float vr = r * r;
        const float t1 = (l - m0) / s0;
        return 0.5f * (vr + t1 * t1 - 1.0f - logf(vr));
    }
    // torch kl.py _kl_laplace_normal
    const float var0 = s0 * s0;
    const float ratio = sg * sg / var0;
    return -0.5f * logf(2.0f * ratio / 3.14159265358979323846f) + ratio + 0.5f * l * l / var0 - l * m0 / var0 +
           0.5f * m0 * m0 / var0 - 1.0f;
}

__device__ __forceinline__ void kl_grads(bool laplace, float l, float sg, float m0, float s0, float& dl, float& dsg,
                                         float& dm0, float& ds0) {
    const float inv_var0 = 1.0f / (s0 * s0);
    const float df = l - m0;
    dl = df * inv_var0;
    dm0 = -dl;
    if (!laplace) {
        dsg = sg * inv_var0 - 1.0f / sg;
        ds0 = -(sg * sg + df * df) * inv_var0 / s0 + 1.0f / s0;
    } else {
        dsg = -1.0f / sg + 2.0f * sg * inv_var0;
        ds0 = 1.0f / s0 - (2.0f * sg * sg + df * df) * inv_var0 / s0;
    }
}

__global__ void __launch_bounds__(kWarps * 32) draws_fwd_kernel(const DrawsParams p) {
    extern __shared__ float sm[];  // fused encoder tail only: (M, Dtot) scales per warp
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    float* s_row = p.enc_tail ? sm + (size_t)(threadIdx.x >> 5) * p.M * p.Dtot : nullptr;
    for (int64_t b = warp0; b < p.B; b += nwarps) {
        if (s_row) {
            __syncwarp();  // the previous row's readers are done
            stage_scales(p, b, s_row, lane);
        }
        for (int j = 0; j < p.n; ++j) {
            const mmvae_draw_desc& d = p.d[j];
            uint32_t mask;
            bool prior;
            desc_mask(p, d, b, mask, prior);
            const bool lap = (d.flags & MMVAE_DRAW_LAPLACE) != 0;
            float klacc = 0.f;
            for (int cc = lane; cc < d.width; cc += 32) {
                float loc, scale, ts;
                fuse_column(p, d, mask, prior, b, d.col0 + cc, loc, scale, ts, s_row);
                if (d.par_off >= 0) {
                    p.ploc[d.par_off + b * d.width + cc] = loc;
                    p.pscale[d.par_off + b * d.width + cc] = scale;
                }
                if (d.kl_mode == 1)
                    klacc += kl_value(lap, loc, scale, __ldg(p.mu0 + cc), __ldg(p.s0 + cc));
                else if (d.kl_mode == 2)
                    klacc += kl_value(lap, loc, scale, 0.f, 1.f);
                for (int k = 0; k < d.K; ++k) {
                    const int64_t idx = ((int64_t)k * p.B + b) * d.width + cc;
                    p.z[d.z_off + idx] = loc + scale * noise_eff(__ldg(p.eps + d.eps_off + idx), lap);
                }
            }
            if (d.kl_mode != 0 && d.kl_off >= 0) {
                klacc = warp_sum(klacc);
                if (lane == 0) p.kl[d.kl_off + b] = klacc;
            }
        }
    }
}

__global__ void __launch_bounds__(kWarps * 32) draws_bwd_kernel(const DrawsParams p) {
    extern __shared__ float sm[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int MD = p.M * p.Dtot;
    const int per_warp = 2 * MD + 2 * p.Dtot + (p.enc_tail ? MD : 0);
    float* acc_mu = sm + (size_t)wid * per_warp;  // (M, Dtot)
    float* acc_s = acc_mu + MD;                   // (M, Dtot)
    float* pr_mu = acc_s + MD;                    // (Dtot) learnable prior grads, summed over rows
    float* pr_s = pr_mu + p.Dtot;
    float* s_row = p.enc_tail ? pr_s + p.Dtot : nullptr;  // (M, Dtot) scales of the row (fused encoder tail)
    for (int i = lane; i < 2 * p.Dtot; i += 32) pr_mu[i] = 0.f;
    const int64_t warp0 = (int64_t)blockIdx.x * kWarps + wid;
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    for (int64_t b = warp0; b < p.B; b += nwarps) {
        for (int i = lane; i < 2 * MD; i += 32) acc_mu[i] = 0.f;
        __syncwarp();
        if (s_row) stage_scales(p, b, s_row, lane);
        for (int j = 0; j < p.n; ++j) {
            const mmvae_draw_desc& d = p.d[j];
            uint32_t mask;
            bool prior;
            desc_mask(p, d, b, mask, prior);
            const bool lap = (d.flags & MMVAE_DRAW_LAPLACE) != 0;
            const float up = (d.kl_mode != 0 && d.kl_off >= 0 && p.dkl) ? __ldg(p.dkl + d.kl_off + b) : 0.f;
            for (int cc = lane; cc < d.width; cc += 32) {
                const int c = d.col0 + cc;
                float loc, scale, ts;
                fuse_column(p, d, mask, prior, b, c, loc, scale, ts, s_row);
                float g_loc = 0.f, g_scale = 0.f;
                if (p.dz) {
                    for (int k = 0; k < d.K; ++k) {
                        const int64_t idx = ((int64_t)k * p.B + b) * d.width + cc;
                        const float dzv = __ldg(p.dz + d.z_off + idx);
                        g_loc += dzv;
                        g_scale += dzv * noise_eff(__ldg(p.eps + d.eps_off + idx), lap);
                    }
                }
                if (d.par_off >= 0 && p.dploc) {
                    g_loc += __ldg(p.dploc + d.par_off + b * d.width + cc);
                    g_scale += __ldg(p.dpscale + d.par_off + b * d.width + cc);
                }
                if (up != 0.f) {
                    float dl, dsg, dm0, ds0;
                    if (d.kl_mode == 1) {
                        kl_grads(lap, loc, scale, __ldg(p.mu0 + cc), __ldg(p.s0 + cc), dl, dsg, dm0, ds0);
                        pr_mu[cc] += up * dm0;
                        pr_s[cc] += up * ds0;
                    } else {
                        kl_grads(lap, loc, scale, 0.f, 1.f, dl, dsg, dm0, ds0);
                    }
                    g_loc += up * dl;
                    g_scale += up * dsg;
                }
                if (d.flags & MMVAE_DRAW_DIRECT) {
                    const int m = __ffs(mask) - 1;
                    acc_mu[m * p.Dtot + c] += g_loc;
                    acc_s[m * p.Dtot + c] += g_scale;
                } else {
                    const float inv_ts = 1.0f / ts;
                    for (int m = 0; m < p.M; ++m) {
                        if (!((mask >> m) & 1u)) continue;
                        const int64_t o = ((int64_t)m * p.B + b) * p.Dtot + c;
                        const float ex = expf(s_row ? s_row[m * p.Dtot + c] : __ldg(p.s + o));
                        const float T = 1.0f / (ex + 1e-8f);
                        const float dT = g_loc * (__ldg(p.mu + o) - loc) * inv_ts - g_scale * inv_ts * inv_ts;
                        acc_mu[m * p.Dtot + c] += g_loc * T * inv_ts;
                        acc_s[m * p.Dtot + c] += -dT * T * T * ex;
                    }
                }
            }
            __syncwarp();  // column ownership may move between lanes when col0 changes
        }
        if (s_row) {  // back through s = softmax(raw) + eta: d/draw = p (d/ds - <d/ds, p>), p = s - eta
            for (int m = 0; m < p.M; ++m) {
                float dot = 0.f;
                for (int c = lane; c < p.Dtot; c += 32) dot += acc_s[m * p.Dtot + c] * (s_row[m * p.Dtot + c] - kEncEta);
                dot = warp_sum(dot);
                for (int c = lane; c < p.Dtot; c += 32)
                    acc_s[m * p.Dtot + c] = (s_row[m * p.Dtot + c] - kEncEta) * (acc_s[m * p.Dtot + c] - dot);
            }
            __syncwarp();
        }
        for (int i = lane; i < MD; i += 32) {
            const int m = i / p.Dtot, c = i - m * p.Dtot;
            const int64_t o = ((int64_t)m * p.B + b) * p.Dtot + c;
            p.dmu[o] = acc_mu[i];
            p.ds[o] = acc_s[i];
        }
        __syncwarp();
    }
    // CTA partial of the prior gradient: fixed warp order -> deterministic
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.Dtot; i += blockDim.x) {
        float tot = 0.f;
        for (int w = 0; w < kWarps; ++w) tot += sm[(size_t)w * per_warp + 2 * MD + i];
        p.ws[(size_t)blockIdx.x * 2 * p.Dtot + i] = tot;
    }
}

// Element-wise KL(q || N(loc0, scale0)) for a (n, D) posterior against a (D) prior row -- the (B, D) tensor that
// calc_kld returns in the reference (objectives.py:148-161) and that the per-dimension analysis tables need
// (utils.py:130-162).  mode 0: forward; mode 1: backward (dl, ds and per-CTA partials of the prior gradient).
__global__ void __launch_bounds__(256) kl_elem_kernel(const float* __restrict__ loc, const float* __restrict__ scale,
                                                      const float* __restrict__ loc0, const float* __restrict__ scale0,
                                                      int laplace, int64_t n, int D, float* __restrict__ out,
                                                      const float* __restrict__ up, float* __restrict__ dloc,
                                                      float* __restrict__ dscale, float* __restrict__ ws) {
    extern __shared__ float pri[];  // 2*D prior-gradient accumulators of this CTA (backward only)
    const bool bwd = up != nullptr;
    if (bwd) {
        for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) pri[i] = 0.f;
        __syncthreads();
    }
    const int64_t total = n * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const float l = loc[i], sg = scale[i], m0 = __ldg(loc0 + c), s0 = __ldg(scale0 + c);
        if (!bwd) {
            out[i] = kl_value(laplace != 0, l, sg, m0, s0);
        } else {
            float dl, dsg, dm0, ds0;
            kl_grads(laplace != 0, l, sg, m0, s0, dl, dsg, dm0, ds0);
            const float u = up[i];
            dloc[i] = u * dl;
            dscale[i] = u * dsg;
            atomicAdd(&pri[c], u * dm0);  // shared-memory accumulation, the cross-CTA stage is ordered
            atomicAdd(&pri[D + c], u * ds0);
        }
    }
    if (bwd) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) ws[(size_t)blockIdx.x * 2 * D + i] = pri[i];
    }
}

// KL between two posteriors of the same family (torch kl.py _kl_normal_normal / _kl_laplace_laplace)
__device__ __forceinline__ float kl_pair(bool laplace, float lp, float sp, float lq, float sq) {
    if (!laplace) return kl_value(false, lp, sp, lq, sq);
    const float ratio = sp / sq, ad = fabsf(lp - lq);
    return -logf(ratio) + ad / sq + ratio * expf(-ad / sp) - 1.0f;
}

// Per-dimension KL tables of the analysis hooks (reference utils.py:130-162 make_kl_df, called from
// trainer.py:242-272 analyse_data): for M posteriors (M, n, D) and the prior row (D), ONE pass writes
//   out[i]           = KL(q_i || p)                                  i < M
//   out[M + pair]    = 0.5 (KL(q_i || q_j) + KL(q_j || q_i))         pairs (i < j) in itertools.combinations order
// as (T, n, D), T = M + M(M-1)/2.  The reference moves every distribution to the CPU and evaluates M + 2*C(M,2)
// separate torch.distributions.kl_divergence calls there.  Every thread reads the 2M parameters of its (row, column)
// once and produces all T values from registers.
__global__ void __launch_bounds__(256) kl_table_kernel(const float* __restrict__ loc, const float* __restrict__ scale,
                                                       const float* __restrict__ loc0, const float* __restrict__ scale0,
                                                       int M, int lap_mask, int64_t n, int D, float* __restrict__ out) {
    const int64_t total = n * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % D);
        const float m0 = __ldg(loc0 + c), s0 = __ldg(scale0 + c);
        float l[MMVAE_MAX_MODS], sg[MMVAE_MAX_MODS];
#pragma unroll
        for (int m = 0; m < MMVAE_MAX_MODS; ++m)
            if (m < M) {
                l[m] = loc[(int64_t)m * total + i];
                sg[m] = scale[(int64_t)m * total + i];
                out[(int64_t)m * total + i] = kl_value((lap_mask >> m) & 1, l[m], sg[m], m0, s0);
            }
        int t = M;
#pragma unroll
        for (int a = 0; a < MMVAE_MAX_MODS; ++a)
#pragma unroll
            for (int b = a + 1; b < MMVAE_MAX_MODS; ++b)
                if (b < M) {
                    const bool lap = (lap_mask >> a) & 1;  // host guarantees equal families inside a pair
                    out[(int64_t)t * total + i] =
                        0.5f * (kl_pair(lap, l[a], sg[a], l[b], sg[b]) + kl_pair(lap, l[b], sg[b], l[a], sg[a]));
                    ++t;
                }
    }
}

static unsigned draws_grid(int64_t B) {
    int64_t g = (B + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)kNumSMs * 4;
    if (g > cap) g = cap;
    return (unsigned)(g < 1 ? 1 : g);
}

static int fill(DrawsParams& p, const float* mu, const float* s, int M, int64_t B, int Dtot,
                const mmvae_draw_desc* descs, int n, const uint32_t* row_masks, const float* mu0, const float* s0,
                const float* eps) {
    if (!mu || !s || !descs || M <= 0 || B <= 0 || Dtot <= 0 || n <= 0) return MMVAE_E_ARG;
    if (M > MMVAE_MAX_MODS || Dtot > MMVAE_MAX_COLS || n > MMVAE_MAX_DRAWS) return MMVAE_E_LIMIT;
    p.mu = mu; p.s = s; p.M = M; p.B = B; p.Dtot = Dtot; p.n = n; p.row_masks = row_masks; p.mu0 = mu0; p.s0 = s0;
    p.eps = eps;
    for (int j = 0; j < n; ++j) {
        const mmvae_draw_desc& d = descs[j];
        if (d.col0 < 0 || d.width <= 0 || d.col0 + d.width > Dtot || d.K < 0) return MMVAE_E_ARG;
        if ((d.flags & MMVAE_DRAW_ROWMASK) && !row_masks) return MMVAE_E_ARG;
        if (!(d.flags & MMVAE_DRAW_ROWMASK) && (d.mask == 0 || (d.mask >> M) != 0)) return MMVAE_E_ARG;
        if ((d.flags & MMVAE_DRAW_DIRECT) && (d.mask & (d.mask - 1))) return MMVAE_E_ARG;  // exactly one modality
        if ((d.flags & MMVAE_DRAW_LAPLACE) && !(d.flags & MMVAE_DRAW_DIRECT)) return MMVAE_E_ARG;
        if (d.kl_mode < 0 || d.kl_mode > 2) return MMVAE_E_ENUM;
        if (d.kl_mode == 1 && (!mu0 || !s0)) return MMVAE_E_ARG;
        if (d.K > 0 && !eps) return MMVAE_E_ARG;
        p.d[j] = d;
    }
    return 0;
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int mmvae_latent_draws_fwd(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                      const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                      const float* mu0, const float* s0, const float* eps, float* z, float* par_loc,
                                      float* par_scale, float* kl, void* stream) {
    return mmvae_latent_draws_fwd_tail(mu, s, M, B, Dtot, descs_host, n_draws, row_masks, mu0, s0, eps, z, par_loc,
                                       par_scale, kl, 0, nullptr, stream);
}

extern "C" int mmvae_latent_draws_fwd_tail(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                           const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                           const float* mu0, const float* s0, const float* eps, float* z, float* par_loc,
                                           float* par_scale, float* kl, int enc_tail, float* s_out, void* stream) {
    DrawsParams p{};
    int rc = fill(p, mu, s, M, B, Dtot, descs_host, n_draws, row_masks, mu0, s0, eps);
    if (rc) return rc;
    for (int j = 0; j < n_draws; ++j) {
        if (p.d[j].K > 0 && !z) return MMVAE_E_ARG;
        if (p.d[j].par_off >= 0 && (!par_loc || !par_scale)) return MMVAE_E_ARG;
        if (p.d[j].kl_mode != 0 && p.d[j].kl_off >= 0 && !kl) return MMVAE_E_ARG;
    }
    p.z = z; p.ploc = par_loc; p.pscale = par_scale; p.kl = kl;
    p.enc_tail = enc_tail; p.s_out = s_out;
    const size_t smem_f = enc_tail ? (size_t)kWarps * M * Dtot * sizeof(float) : 0;  // <= 8 * 8 * 256 * 4 = 64 KB
    if (smem_f > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(draws_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
        if (e != cudaSuccess) return (int)e;
    }
    draws_fwd_kernel<<<draws_grid(B), kWarps * 32, smem_f, (cudaStream_t)stream>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t mmvae_latent_draws_bwd_ws_floats(int64_t B, int Dtot) {
    return (int64_t)draws_grid(B) * 2 * Dtot;
}

extern "C" int mmvae_latent_draws_bwd(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                      const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                      const float* mu0, const float* s0, const float* eps, const float* dz,
                                      const float* dkl, const float* dpar_loc, const float* dpar_scale, float* dmu,
                                      float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream) {
    return mmvae_latent_draws_bwd_tail(mu, s, M, B, Dtot, descs_host, n_draws, row_masks, mu0, s0, eps, dz, dkl,
                                       dpar_loc, dpar_scale, 0, dmu, ds, dprior_ws, dmu0, ds0, stream);
}

extern "C" int mmvae_latent_draws_bwd_tail(const float* mu, const float* s, int M, int64_t B, int Dtot,
                                           const mmvae_draw_desc* descs_host, int n_draws, const uint32_t* row_masks,
                                           const float* mu0, const float* s0, const float* eps, const float* dz,
                                           const float* dkl, const float* dpar_loc, const float* dpar_scale,
                                           int enc_tail, float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0,
                                           void* stream) {
    DrawsParams p{};
    int rc = fill(p, mu, s, M, B, Dtot, descs_host, n_draws, row_masks, mu0, s0, eps);
    if (rc) return rc;
    if (!dmu || !ds || !dprior_ws) return MMVAE_E_ARG;
    if ((dpar_loc == nullptr) != (dpar_scale == nullptr)) return MMVAE_E_ARG;
    p.dz = dz; p.dkl = dkl; p.dploc = dpar_loc; p.dpscale = dpar_scale; p.dmu = dmu; p.ds = ds; p.ws = dprior_ws;
    p.enc_tail = enc_tail;
    const unsigned grid = draws_grid(B);
    const size_t smem = (size_t)kWarps * (2 * M * Dtot + 2 * Dtot + (enc_tail ? M * Dtot : 0)) * sizeof(float);
    if (smem > 200 * 1024) return MMVAE_E_LIMIT;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(draws_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaStream_t st = (cudaStream_t)stream;
    draws_bwd_kernel<<<grid, kWarps * 32, smem, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    if (dmu0 && ds0) {
        partial_sum_kernel<<<2 * Dtot, 128, 0, st>>>(dprior_ws, (int)grid, 2 * Dtot, Dtot, dmu0, ds0);
        MMVAE_LAUNCH_CHECK();
    }
    return 0;
}

static unsigned kl_grid(int64_t total) {
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 4;
    return (unsigned)(g > cap ? cap : (g < 1 ? 1 : g));
}

extern "C" int mmvae_kl_table(const float* loc, const float* scale, int M, const int32_t* dists_host, const float* loc0,
                              const float* scale0, int64_t n, int D, float* out, void* stream) {
    if (!loc || !scale || !dists_host || !loc0 || !scale0 || !out || M <= 0 || n <= 0 || D <= 0) return MMVAE_E_ARG;
    if (M > MMVAE_MAX_MODS) return MMVAE_E_LIMIT;
    int mask = 0;
    for (int m = 0; m < M; ++m) {
        if (dists_host[m] != MMVAE_NORMAL && dists_host[m] != MMVAE_LAPLACE) return MMVAE_E_ENUM;
        if (dists_host[m] != dists_host[0]) return MMVAE_E_ENUM;  // symmetric J needs both directions in closed form
        mask |= (dists_host[m] == MMVAE_LAPLACE) << m;
    }
    kl_table_kernel<<<kl_grid(n * D), 256, 0, (cudaStream_t)stream>>>(loc, scale, loc0, scale0, M, mask, n, D, out);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t mmvae_kl_elementwise_ws_floats(int64_t n, int D) { return (int64_t)kl_grid(n * D) * 2 * D; }

extern "C" int mmvae_kl_elementwise_fwd(const float* loc, const float* scale, const float* loc0, const float* scale0,
                                        int dist, int64_t n, int D, float* out, void* stream) {
    if (!loc || !scale || !loc0 || !scale0 || !out || n <= 0 || D <= 0) return MMVAE_E_ARG;
    if (dist != MMVAE_NORMAL && dist != MMVAE_LAPLACE) return MMVAE_E_ENUM;
    kl_elem_kernel<<<kl_grid(n * D), 256, 0, (cudaStream_t)stream>>>(loc, scale, loc0, scale0, dist, n, D, out, nullptr,
                                                                     nullptr, nullptr, nullptr);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_kl_elementwise_bwd(const float* loc, const float* scale, const float* loc0, const float* scale0,
                                        int dist, int64_t n, int D, const float* upstream, float* dloc, float* dscale,
                                        float* ws, float* dloc0, float* dscale0, void* stream) {
    if (!loc || !scale || !loc0 || !scale0 || !upstream || !dloc || !dscale || !ws || n <= 0 || D <= 0)
        return MMVAE_E_ARG;
    if (dist != MMVAE_NORMAL && dist != MMVAE_LAPLACE) return MMVAE_E_ENUM;
    if (D > 4096) return MMVAE_E_LIMIT;
    const unsigned grid = kl_grid(n * D);
    cudaStream_t st = (cudaStream_t)stream;
    kl_elem_kernel<<<grid, 256, 2 * D * sizeof(float), st>>>(loc, scale, loc0, scale0, dist, n, D, nullptr, upstream,
                                                             dloc, dscale, ws);
    MMVAE_LAUNCH_CHECK();
    if (dloc0 && dscale0) {
        partial_sum_kernel<<<2 * D, 128, 0, st>>>(ws, (int)grid, 2 * D, D, dloc0, dscale0);
        MMVAE_LAUNCH_CHECK();
    }
    return 0;
}
