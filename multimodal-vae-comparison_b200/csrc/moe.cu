// MoE (MMVAE) fused reparameterised sampling + log-densities, forward and backward.
//
// Replaces reference mmvae_models.py:99 (q_r.rsample([K])), :56-62 (ELBO importance terms log q_r(z_s) - log q_s(z_s))
// and objectives.py:342-373 (log p(z_r), log-mean_j q_j(z_r) of iwae / _m_dreg_looser) -- there a python loop of
// torch.distributions calls that materialises M*M (K,B,D) log-prob tensors; here one pass per direction:
//   fwd: eps -> z (written once, needed by the decoders) + lq (M,M,K,B) + lpz (M,K,B)
//   bwd: dz_ext (from the decoders) + eps (z is recomputed, never re-read) -> dmu, ds, dprior
// i.e. the 4*K*D*e bytes per drawn sample tensor that SURVEY.md 8d counts as the algorithmic minimum.
//
// Mapping: one CTA (4 warps) per batch row, grid-stride; the row's (M, D) posterior parameters are staged in
// shared memory once ("inv sigma" and the log normaliser precomputed), warps split the K samples, lanes stride
// over the latent columns (contiguous 128 B segments), the D-sums are warp-shuffle reductions.  Backward keeps
// per-warp, lane-private (M, D) accumulators in shared memory and combines them in fixed order: deterministic,
// no atomics.
#include "common.cuh"

namespace mmvae {

constexpr int kMoeMaxWarps = 16;  // warps per CTA = min(K, 16): the K samples of a row are split over the warps
constexpr float kLogSqrt2Pi = 0.91893853320467274178f;

struct MoeParams {
    const float *mu, *s, *mu0, *s0, *eps;
    float *z, *lq, *lpz;                 // forward outputs
    const float *dz_ext, *dlq, *dlpz;    // backward inputs
    float *dmu, *ds, *ws;                // backward outputs
    int64_t B;
    int M, D, K, through_z;
    int dist[MMVAE_MAX_MODS];
};

// -log1p(-a) for a in [0, 1): torch's Laplace.rsample evaluates log1p(-|u|) (laplace.py:84).  The libdevice log1pf
// costs ~35 instructions and made the Laplace kernels instruction bound (r1: 1.7 TB/s on C4 latent-only).  Here:
// a < 1/8: 8-term series (truncation < 7e-9 relative); otherwise -log(w + d) with w = fl(1-a), d the exactly
// recovered rounding residual, = -(log w + d/w) through MUFU.LG2/RCP: absolute error 1.7e-7 on |value| >= 0.13,
// i.e. <= 1.3e-6 relative -- an order of magnitude inside the 1e-5 parity tolerance (tests/test_ops_gpu.py).
__device__ __forceinline__ float neg_log1p_neg(float a) {
    const float w = 1.0f - a;
    const float d = (1.0f - w) - a;  // exact: (1 - a) = w + d
    const float big = -(__log2f(w) * 0.69314718055994530942f + __fdividef(d, w));
    float sm = 0.125f;  // 1/8
    sm = fmaf(sm, a, 1.0f / 7.0f);
    sm = fmaf(sm, a, 1.0f / 6.0f);
    sm = fmaf(sm, a, 0.2f);
    sm = fmaf(sm, a, 0.25f);
    sm = fmaf(sm, a, 1.0f / 3.0f);
    sm = fmaf(sm, a, 0.5f);
    sm = fmaf(sm, a, 1.0f);
    return a < 0.125f ? sm * a : big;
}

__device__ __forceinline__ float eff_noise(float e, bool laplace) {
    if (!laplace) return e;
    const float v = neg_log1p_neg(fabsf(e));  // -log1p(-|u|) >= 0
    return e > 0.f ? v : (e < 0.f ? -v : 0.f);
}

// stage mu / 1/sigma / log-normaliser of the row's M posteriors (+ the prior once per CTA)
__device__ __forceinline__ void stage_row(const MoeParams& p, int64_t b, float* smu, float* ssig, float* sinv,
                                          float* scst) {
    const int MD = p.M * p.D;
    for (int i = threadIdx.x; i < MD; i += blockDim.x) {
        const int m = i / p.D, c = i - m * p.D;
        const int64_t o = ((int64_t)m * p.B + b) * p.D + c;
        const float sg = __ldg(p.s + o);
        smu[i] = __ldg(p.mu + o);
        ssig[i] = sg;
        sinv[i] = 1.0f / sg;
        scst[i] = p.dist[m] == MMVAE_LAPLACE ? -logf(2.0f * sg) : -logf(sg) - kLogSqrt2Pi;
    }
}

// MT: compile-time modality count (0 = generic, loops predicated up to MMVAE_MAX_MODS)
template <int MT>
__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_fwd_kernel(const MoeParams p) {
    constexpr int MM = MT ? MT : MMVAE_MAX_MODS;
    const int kMoeWarps = blockDim.x >> 5;
    extern __shared__ float sm[];
    const int MD = p.M * p.D;
    float* smu = sm;
    float* ssig = smu + MD;
    float* sinv = ssig + MD;
    float* scst = sinv + MD;
    float* pmu = scst + MD;
    float* pinv = pmu + p.D;
    float* pcst = pinv + p.D;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = threadIdx.x; c < p.D; c += blockDim.x) {
        const float sg = __ldg(p.s0 + c);
        pmu[c] = __ldg(p.mu0 + c);
        pinv[c] = 1.0f / sg;
        pcst[c] = -logf(sg) - kLogSqrt2Pi;
    }
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        stage_row(p, b, smu, ssig, sinv, scst);
        __syncthreads();
        for (int k = wid; k < p.K; k += kMoeWarps) {
            for (int r = 0; r < p.M; ++r) {
                const bool lap = p.dist[r] == MMVAE_LAPLACE;
                float lq[MMVAE_MAX_MODS];
#pragma unroll
                for (int j = 0; j < MM; ++j) lq[j] = 0.f;
                float lp = 0.f;
                const int64_t base = (((int64_t)r * p.K + k) * p.B + b) * p.D;
                for (int c = lane; c < p.D; c += 32) {
                    const float zz = smu[r * p.D + c] + eff_noise(__ldg(p.eps + base + c), lap) * ssig[r * p.D + c];
                    p.z[base + c] = zz;
#pragma unroll
                    for (int j = 0; j < MM; ++j) {
                        if (MT || j < p.M) {
                            const float u = (zz - smu[j * p.D + c]) * sinv[j * p.D + c];
                            lq[j] += (p.dist[j] == MMVAE_LAPLACE ? -fabsf(u) : -0.5f * u * u) + scst[j * p.D + c];
                        }
                    }
                    const float u0 = (zz - pmu[c]) * pinv[c];
                    lp += -0.5f * u0 * u0 + pcst[c];
                }
#pragma unroll
                for (int j = 0; j < MM; ++j) {
                    if (MT || j < p.M) {
                        const float v = warp_sum(lq[j]);
                        if (lane == 0) p.lq[(((int64_t)r * p.M + j) * p.K + k) * p.B + b] = v;
                    }
                }
                lp = warp_sum(lp);
                if (lane == 0) p.lpz[((int64_t)r * p.K + k) * p.B + b] = lp;
            }
        }
    }
}

template <int MT>
__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_bwd_kernel(const MoeParams p) {
    constexpr int MM = MT ? MT : MMVAE_MAX_MODS;
    const int kMoeWarps = blockDim.x >> 5;
    extern __shared__ float sm[];
    const int MD = p.M * p.D;
    float* smu = sm;
    float* ssig = smu + MD;
    float* sinv = ssig + MD;
    float* scst = sinv + MD;  // unused in bwd but keeps stage_row shared
    float* pmu = scst + MD;
    float* pinv = pmu + p.D;
    float* acc = pinv + p.D;  // per warp: acc_mu (MD), acc_s (MD), pr_mu (D), pr_s (D)
    const int per_warp = 2 * MD + 2 * p.D;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* a_mu = acc + (size_t)wid * per_warp;
    float* a_s = a_mu + MD;
    float* q_mu = a_s + MD;
    float* q_s = q_mu + p.D;
    for (int c = threadIdx.x; c < p.D; c += blockDim.x) {
        pmu[c] = __ldg(p.mu0 + c);
        pinv[c] = 1.0f / __ldg(p.s0 + c);
    }
    for (int i = lane; i < 2 * p.D; i += 32) q_mu[i] = 0.f;
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        stage_row(p, b, smu, ssig, sinv, scst);
        for (int i = lane; i < 2 * MD; i += 32) a_mu[i] = 0.f;
        __syncthreads();
        for (int k = wid; k < p.K; k += kMoeWarps) {
            for (int r = 0; r < p.M; ++r) {
                const bool lap = p.dist[r] == MMVAE_LAPLACE;
                float cj[MMVAE_MAX_MODS];
#pragma unroll
                for (int j = 0; j < MM; ++j)
                    cj[j] = ((MT || j < p.M) && p.dlq) ? __ldg(p.dlq + (((int64_t)r * p.M + j) * p.K + k) * p.B + b) : 0.f;
                const float cp = p.dlpz ? __ldg(p.dlpz + ((int64_t)r * p.K + k) * p.B + b) : 0.f;
                const int64_t base = (((int64_t)r * p.K + k) * p.B + b) * p.D;
                for (int c = lane; c < p.D; c += 32) {
                    const float ef = eff_noise(__ldg(p.eps + base + c), lap);
                    const float zz = smu[r * p.D + c] + ef * ssig[r * p.D + c];
                    float dzt = p.dz_ext ? __ldg(p.dz_ext + base + c) : 0.f;
                    float dz_ld = 0.f;  // d(sum_j c_j log q_j + c_p log p)/dz
#pragma unroll
                    for (int j = 0; j < MM; ++j) {
                        if (MT || j < p.M) {
                            const float inv = sinv[j * p.D + c];
                            const float df = zz - smu[j * p.D + c];
                            float dmu_j, ds_j;
                            if (p.dist[j] == MMVAE_LAPLACE) {
                                const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
                                dmu_j = sg * inv;
                                ds_j = fabsf(df) * inv * inv - inv;
                            } else {
                                dmu_j = df * inv * inv;
                                ds_j = df * df * inv * inv * inv - inv;
                            }
                            a_mu[j * p.D + c] += cj[j] * dmu_j;
                            a_s[j * p.D + c] += cj[j] * ds_j;
                            dz_ld -= cj[j] * dmu_j;
                        }
                    }
                    {
                        const float inv = pinv[c];
                        const float df = zz - pmu[c];
                        const float dm0 = df * inv * inv;
                        q_mu[c] += cp * dm0;
                        q_s[c] += cp * (df * df * inv * inv * inv - inv);
                        dz_ld -= cp * dm0;
                    }
                    if (p.through_z) dzt += dz_ld;
                    a_mu[r * p.D + c] += dzt;
                    a_s[r * p.D + c] += dzt * ef;
                }
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < MD; i += blockDim.x) {
            float tm = 0.f, ts = 0.f;
            for (int w = 0; w < kMoeWarps; ++w) {
                tm += acc[(size_t)w * per_warp + i];
                ts += acc[(size_t)w * per_warp + MD + i];
            }
            const int m = i / p.D, c = i - m * p.D;
            const int64_t o = ((int64_t)m * p.B + b) * p.D + c;
            p.dmu[o] = tm;
            p.ds[o] = ts;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * p.D; i += blockDim.x) {
        float tot = 0.f;
        for (int w = 0; w < kMoeWarps; ++w) tot += acc[(size_t)w * per_warp + 2 * MD + i];
        p.ws[(size_t)blockIdx.x * 2 * p.D + i] = tot;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Register-resident variants (MT modalities, NC columns per lane, MT*NC <= 8): every per-column constant of the row
// (mu, sigma, 1/sigma, log normaliser of each posterior, the prior) and -- in the backward -- every gradient
// accumulator lives in registers for the whole (k, r) loop; shared memory is touched once per row, for the fixed-order
// cross-warp combine.  r1 ncu on C4 (D=64, K=50): the smem-staged kernels executed 35 M / 44 M warp instructions
// (~175 per column, constants re-read and accumulators read-modify-written in smem inside the loop).
// ---------------------------------------------------------------------------------------------------------
// LM: 0 = every posterior Normal, 1 = every posterior Laplace, 2 = mixed (decided at run time per modality);
// FULL: D == 32*NC, every lane owns NC valid columns (no column guards, no divergence around the shuffles).
template <int MT, int NC, int LM, bool FULL>
__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_fwd_reg_kernel(const MoeParams p) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    bool lap[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) lap[j] = LM == 2 ? (p.dist[j] == MMVAE_LAPLACE) : (LM == 1);
    float pmu[NC], pinv[NC], pcst[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        const float sg = (FULL || c < p.D) ? __ldg(p.s0 + c) : 1.f;
        pmu[i] = (FULL || c < p.D) ? __ldg(p.mu0 + c) : 0.f;
        pinv[i] = 1.0f / sg;
        pcst[i] = -logf(sg) - kLogSqrt2Pi;
    }
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        float rmu[MT][NC], rsig[MT][NC], rinv[MT][NC], rcst[MT][NC];
#pragma unroll
        for (int j = 0; j < MT; ++j)
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int c = lane + 32 * i;
                const int64_t o = ((int64_t)j * p.B + b) * p.D + c;
                const float sg = (FULL || c < p.D) ? __ldg(p.s + o) : 1.f;
                rmu[j][i] = (FULL || c < p.D) ? __ldg(p.mu + o) : 0.f;
                rsig[j][i] = sg;
                rinv[j][i] = 1.0f / sg;
                rcst[j][i] = lap[j] ? -logf(2.0f * sg) : -logf(sg) - kLogSqrt2Pi;
            }
        for (int k = wid; k < p.K; k += nw) {
            float e[MT][NC];
#pragma unroll
            for (int r = 0; r < MT; ++r)
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int c = lane + 32 * i;
                    e[r][i] = (FULL || c < p.D) ? __ldg(p.eps + (((int64_t)r * p.K + k) * p.B + b) * p.D + c) : 0.f;
                }
#pragma unroll
            for (int r = 0; r < MT; ++r) {
                const int64_t base = (((int64_t)r * p.K + k) * p.B + b) * p.D;
                float lq[MT], lp = 0.f;
#pragma unroll
                for (int j = 0; j < MT; ++j) lq[j] = 0.f;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int c = lane + 32 * i;
                    if (FULL || c < p.D) {
                        const float zz = rmu[r][i] + eff_noise(e[r][i], lap[r]) * rsig[r][i];
                        p.z[base + c] = zz;
#pragma unroll
                        for (int j = 0; j < MT; ++j) {
                            const float u = (zz - rmu[j][i]) * rinv[j][i];
                            lq[j] += (lap[j] ? -fabsf(u) : -0.5f * u * u) + rcst[j][i];
                        }
                        const float u0 = (zz - pmu[i]) * pinv[i];
                        lp += -0.5f * u0 * u0 + pcst[i];
                    }
                }
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    const float v = warp_sum(lq[j]);
                    if (lane == 0) p.lq[(((int64_t)r * MT + j) * p.K + k) * p.B + b] = v;
                }
                lp = warp_sum(lp);
                if (lane == 0) p.lpz[((int64_t)r * p.K + k) * p.B + b] = lp;
            }
        }
    }
}

template <int MT, int NC, int LM, bool FULL>
__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_bwd_reg_kernel(const MoeParams p) {
    extern __shared__ float sm[];  // nw x (2*MT*NC + 2*NC) x 32 : per-warp register dumps for the combine
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    constexpr int kAcc = 2 * MT * NC, kPri = 2 * NC, kPer = (kAcc + kPri) * 32;
    bool lap[MT];
#pragma unroll
    for (int j = 0; j < MT; ++j) lap[j] = LM == 2 ? (p.dist[j] == MMVAE_LAPLACE) : (LM == 1);
    float pmu[NC], pinv[NC], q_mu[NC], q_s[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        const int c = lane + 32 * i;
        pmu[i] = (FULL || c < p.D) ? __ldg(p.mu0 + c) : 0.f;
        pinv[i] = 1.0f / ((FULL || c < p.D) ? __ldg(p.s0 + c) : 1.f);
        q_mu[i] = 0.f;
        q_s[i] = 0.f;
    }
    for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
        float rmu[MT][NC], rsig[MT][NC], rinv[MT][NC], a_mu[MT][NC], a_s[MT][NC];
#pragma unroll
        for (int j = 0; j < MT; ++j)
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const int c = lane + 32 * i;
                const int64_t o = ((int64_t)j * p.B + b) * p.D + c;
                const float sg = (FULL || c < p.D) ? __ldg(p.s + o) : 1.f;
                rmu[j][i] = (FULL || c < p.D) ? __ldg(p.mu + o) : 0.f;
                rsig[j][i] = sg;
                rinv[j][i] = 1.0f / sg;
                a_mu[j][i] = 0.f;
                a_s[j][i] = 0.f;
            }
        for (int k = wid; k < p.K; k += nw) {
            float e[MT][NC], dzx[MT][NC], cj[MT][MT], cp[MT];
#pragma unroll
            for (int r = 0; r < MT; ++r) {
#pragma unroll
                for (int j = 0; j < MT; ++j)
                    cj[r][j] = p.dlq ? __ldg(p.dlq + (((int64_t)r * MT + j) * p.K + k) * p.B + b) : 0.f;
                cp[r] = p.dlpz ? __ldg(p.dlpz + ((int64_t)r * p.K + k) * p.B + b) : 0.f;
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const int c = lane + 32 * i;
                    const int64_t o = (((int64_t)r * p.K + k) * p.B + b) * p.D + c;
                    e[r][i] = (FULL || c < p.D) ? __ldg(p.eps + o) : 0.f;
                    dzx[r][i] = ((FULL || c < p.D) && p.dz_ext) ? __ldg(p.dz_ext + o) : 0.f;
                }
            }
#pragma unroll
            for (int r = 0; r < MT; ++r)
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    if (FULL || lane + 32 * i < p.D) {
                        const float ef = eff_noise(e[r][i], lap[r]);
                        const float zz = rmu[r][i] + ef * rsig[r][i];
                        float dzt = dzx[r][i], dz_ld = 0.f;
#pragma unroll
                        for (int j = 0; j < MT; ++j) {
                            const float inv = rinv[j][i], df = zz - rmu[j][i];
                            float dmu_j, ds_j;
                            if (lap[j]) {
                                const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
                                dmu_j = sg * inv;
                                ds_j = fabsf(df) * inv * inv - inv;
                            } else {
                                dmu_j = df * inv * inv;
                                ds_j = df * df * inv * inv * inv - inv;
                            }
                            a_mu[j][i] += cj[r][j] * dmu_j;
                            a_s[j][i] += cj[r][j] * ds_j;
                            dz_ld -= cj[r][j] * dmu_j;
                        }
                        {
                            const float inv = pinv[i], df = zz - pmu[i], dm0 = df * inv * inv;
                            q_mu[i] += cp[r] * dm0;
                            q_s[i] += cp[r] * (df * df * inv * inv * inv - inv);
                            dz_ld -= cp[r] * dm0;
                        }
                        if (p.through_z) dzt += dz_ld;
                        a_mu[r][i] += dzt;
                        a_s[r][i] += dzt * ef;
                    }
                }
        }
        // fixed-order cross-warp combine through shared memory (deterministic, no atomics)
        __syncthreads();
        float* mine = sm + (size_t)wid * kPer;
#pragma unroll
        for (int j = 0; j < MT; ++j)
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                mine[((j * NC + i) * 2 + 0) * 32 + lane] = a_mu[j][i];
                mine[((j * NC + i) * 2 + 1) * 32 + lane] = a_s[j][i];
            }
        __syncthreads();
        for (int t = threadIdx.x; t < kAcc * 32; t += blockDim.x) {
            float tot = 0.f;
            for (int w = 0; w < nw; ++w) tot += sm[(size_t)w * kPer + t];
            const int ln = t & 31, q = t >> 5, which = q & 1, ji = q >> 1, j = ji / NC, i = ji - j * NC;
            const int c = ln + 32 * i;
            if (FULL || c < p.D) {
                const int64_t o = ((int64_t)j * p.B + b) * p.D + c;
                if (which == 0) p.dmu[o] = tot;
                else p.ds[o] = tot;
            }
        }
    }
    __syncthreads();
    float* mine = sm + (size_t)wid * kPer + kAcc * 32;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        mine[(i * 2 + 0) * 32 + lane] = q_mu[i];
        mine[(i * 2 + 1) * 32 + lane] = q_s[i];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kPri * 32; t += blockDim.x) {
        float tot = 0.f;
        for (int w = 0; w < nw; ++w) tot += sm[(size_t)w * kPer + kAcc * 32 + t];
        const int ln = t & 31, q = t >> 5, which = q & 1, i = q >> 1;
        const int c = ln + 32 * i;
        if (c < p.D) p.ws[(size_t)blockIdx.x * 2 * p.D + which * p.D + c] = tot;
    }
}

typedef void (*moe_kernel_t)(const MoeParams);
template <bool FWD, int MT_, int NC_>
static moe_kernel_t pick_variant(int lm, bool full) {
#define MOE_V(LM_, FULL_) \
    (FWD ? (moe_kernel_t)moe_fwd_reg_kernel<MT_, NC_, LM_, FULL_> : (moe_kernel_t)moe_bwd_reg_kernel<MT_, NC_, LM_, FULL_>)
    if (MT_ * NC_ <= 4) {  // common shapes get the family / full-lane specialisations
        if (lm == 0) return full ? MOE_V(0, true) : MOE_V(0, false);
        if (lm == 1) return full ? MOE_V(1, true) : MOE_V(1, false);
    }
    return MOE_V(2, false);
#undef MOE_V
}

template <bool FWD>
static moe_kernel_t pick_reg_kernel(int M, int D, const int* dist, int* nc_out) {
    const int nc = D <= 32 ? 1 : (D <= 64 ? 2 : (D <= 128 ? 4 : 8));
    *nc_out = nc;
    int nlap = 0;
    for (int j = 0; j < M; ++j) nlap += dist[j] == MMVAE_LAPLACE;
    const int lm = nlap == 0 ? 0 : (nlap == M ? 1 : 2);
    const bool full = D == 32 * nc;
#define MOE_PICK(MT_, NC_) \
    if (M == MT_ && nc == NC_) return pick_variant<FWD, MT_, NC_>(lm, full);
    MOE_PICK(1, 1) MOE_PICK(1, 2) MOE_PICK(1, 4) MOE_PICK(1, 8)
    MOE_PICK(2, 1) MOE_PICK(2, 2) MOE_PICK(2, 4)
    MOE_PICK(3, 1) MOE_PICK(3, 2)
    MOE_PICK(4, 1) MOE_PICK(4, 2)
#undef MOE_PICK
    return nullptr;
}

static unsigned moe_grid(int64_t B) {
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(B < cap ? B : cap);
}
static int moe_warps(int K) { return K < kMoeMaxWarps ? (K < 1 ? 1 : K) : kMoeMaxWarps; }

static int moe_fill(MoeParams& p, const float* mu, const float* s, int M, int64_t B, int D, int K,
                    const int32_t* dists, const float* mu0, const float* s0, const float* eps) {
    if (!mu || !s || !dists || !mu0 || !s0 || !eps || M <= 0 || B <= 0 || D <= 0 || K <= 0) return MMVAE_E_ARG;
    if (M > MMVAE_MAX_MODS || D > MMVAE_MAX_COLS) return MMVAE_E_LIMIT;
    p.mu = mu; p.s = s; p.mu0 = mu0; p.s0 = s0; p.eps = eps; p.M = M; p.B = B; p.D = D; p.K = K;
    for (int m = 0; m < M; ++m) {
        if (dists[m] != MMVAE_NORMAL && dists[m] != MMVAE_LAPLACE) return MMVAE_E_ENUM;
        p.dist[m] = dists[m];
    }
    return 0;
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int mmvae_moe_logdens_fwd(const float* mu, const float* s, int M, int64_t B, int D, int K,
                                     const int32_t* dists_host, const float* mu0, const float* s0, const float* eps,
                                     float* z, float* lq, float* lpz, void* stream) {
    MoeParams p{};
    int rc = moe_fill(p, mu, s, M, B, D, K, dists_host, mu0, s0, eps);
    if (rc) return rc;
    if (!z || !lq || !lpz) return MMVAE_E_ARG;
    p.z = z; p.lq = lq; p.lpz = lpz;
    const size_t smem = (size_t)(4 * M * D + 3 * D) * sizeof(float);
    int nc = 0;
    if (moe_kernel_t kr = pick_reg_kernel<true>(M, D, p.dist, &nc)) {
        kr<<<moe_grid(B), moe_warps(K) * 32, 0, (cudaStream_t)stream>>>(p);
        MMVAE_LAUNCH_CHECK();
        return 0;
    }
    auto kf = M == 1 ? moe_fwd_kernel<1> : M == 2 ? moe_fwd_kernel<2> : M == 3 ? moe_fwd_kernel<3>
                                                                              : M == 4 ? moe_fwd_kernel<4> : moe_fwd_kernel<0>;
    kf<<<moe_grid(B), moe_warps(K) * 32, smem, (cudaStream_t)stream>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int64_t mmvae_moe_logdens_bwd_ws_floats(int64_t B, int D, int K) {
    (void)K;
    return (int64_t)moe_grid(B) * 2 * D;
}

extern "C" int mmvae_moe_logdens_bwd(const float* mu, const float* s, int M, int64_t B, int D, int K,
                                     const int32_t* dists_host, const float* mu0, const float* s0, const float* eps,
                                     const float* dz_ext, const float* dlq, const float* dlpz, int through_z,
                                     float* dmu, float* ds, float* dprior_ws, float* dmu0, float* ds0, void* stream) {
    MoeParams p{};
    int rc = moe_fill(p, mu, s, M, B, D, K, dists_host, mu0, s0, eps);
    if (rc) return rc;
    if (!dmu || !ds || !dprior_ws) return MMVAE_E_ARG;
    p.dz_ext = dz_ext; p.dlq = dlq; p.dlpz = dlpz; p.through_z = through_z; p.dmu = dmu; p.ds = ds; p.ws = dprior_ws;
    const int nw = moe_warps(K);
    const size_t smem = (size_t)(4 * M * D + 2 * D + nw * (2 * M * D + 2 * D)) * sizeof(float);
    if (smem > 200 * 1024) return MMVAE_E_LIMIT;
    auto kb = M == 1 ? moe_bwd_kernel<1> : M == 2 ? moe_bwd_kernel<2> : M == 3 ? moe_bwd_kernel<3>
                                                                              : M == 4 ? moe_bwd_kernel<4> : moe_bwd_kernel<0>;
    int nc = 0;
    if (moe_kernel_t kr = pick_reg_kernel<false>(M, D, p.dist, &nc)) {
        const size_t smem_r = (size_t)nw * (2 * M * nc + 2 * nc) * 32 * sizeof(float);
        if (smem_r > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute((const void*)kr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r);
            if (e != cudaSuccess) return (int)e;
        }
        const unsigned grid_r = moe_grid(B);
        kr<<<grid_r, nw * 32, smem_r, (cudaStream_t)stream>>>(p);
        MMVAE_LAUNCH_CHECK();
        if (dmu0 && ds0) {
            partial_sum_kernel<<<2 * D, 128, 0, (cudaStream_t)stream>>>(dprior_ws, (int)grid_r, 2 * D, D, dmu0, ds0);
            MMVAE_LAUNCH_CHECK();
        }
        return 0;
    }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = moe_grid(B);
    kb<<<grid, nw * 32, smem, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    if (dmu0 && ds0) {
        partial_sum_kernel<<<2 * D, 128, 0, st>>>(dprior_ws, (int)grid, 2 * D, D, dmu0, ds0);
        MMVAE_LAUNCH_CHECK();
    }
    return 0;
}
