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

__device__ __forceinline__ float eff_noise(float e, bool laplace) {
    if (!laplace) return e;
    const float sg = e > 0.f ? 1.f : (e < 0.f ? -1.f : 0.f);
    return -sg * log1pf(-fabsf(e));
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

__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_fwd_kernel(const MoeParams p) {
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
                for (int j = 0; j < MMVAE_MAX_MODS; ++j) lq[j] = 0.f;
                float lp = 0.f;
                const int64_t base = (((int64_t)r * p.K + k) * p.B + b) * p.D;
                for (int c = lane; c < p.D; c += 32) {
                    const float zz = smu[r * p.D + c] + eff_noise(__ldg(p.eps + base + c), lap) * ssig[r * p.D + c];
                    p.z[base + c] = zz;
#pragma unroll
                    for (int j = 0; j < MMVAE_MAX_MODS; ++j) {
                        if (j < p.M) {
                            const float u = (zz - smu[j * p.D + c]) * sinv[j * p.D + c];
                            lq[j] += (p.dist[j] == MMVAE_LAPLACE ? -fabsf(u) : -0.5f * u * u) + scst[j * p.D + c];
                        }
                    }
                    const float u0 = (zz - pmu[c]) * pinv[c];
                    lp += -0.5f * u0 * u0 + pcst[c];
                }
#pragma unroll
                for (int j = 0; j < MMVAE_MAX_MODS; ++j) {
                    if (j < p.M) {
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

__global__ void __launch_bounds__(kMoeMaxWarps * 32) moe_bwd_kernel(const MoeParams p) {
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
                for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                    cj[j] = (j < p.M && p.dlq) ? __ldg(p.dlq + (((int64_t)r * p.M + j) * p.K + k) * p.B + b) : 0.f;
                const float cp = p.dlpz ? __ldg(p.dlpz + ((int64_t)r * p.K + k) * p.B + b) : 0.f;
                const int64_t base = (((int64_t)r * p.K + k) * p.B + b) * p.D;
                for (int c = lane; c < p.D; c += 32) {
                    const float ef = eff_noise(__ldg(p.eps + base + c), lap);
                    const float zz = smu[r * p.D + c] + ef * ssig[r * p.D + c];
                    float dzt = p.dz_ext ? __ldg(p.dz_ext + base + c) : 0.f;
                    float dz_ld = 0.f;  // d(sum_j c_j log q_j + c_p log p)/dz
#pragma unroll
                    for (int j = 0; j < MMVAE_MAX_MODS; ++j) {
                        if (j < p.M) {
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
    moe_fwd_kernel<<<moe_grid(B), moe_warps(K) * 32, smem, (cudaStream_t)stream>>>(p);
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
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(moe_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = moe_grid(B);
    moe_bwd_kernel<<<grid, nw * 32, smem, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    if (dmu0 && ds0) {
        partial_sum_kernel<<<2 * D, 128, 0, st>>>(dprior_ws, (int)grid, 2 * D, D, dmu0, ds0);
        MMVAE_LAUNCH_CHECK();
    }
    return 0;
}
