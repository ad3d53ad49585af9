// Objective combination kernels: IWAE logsumexp over (modality, K), DReG softmax over K of batch-summed
// log-weights, deterministic sums for the ELBO, and the conditional gradient rescale.
// Replaces reference objectives.py:54-67 (elbo), :342-359 (iwae, with the N1 shim), :361-387 (_m_dreg_looser, dreg)
// and utils.py:395-396 (log_mean_exp).  All row tensors keep the reference's k-major layout (row = k*B + b), so
// lanes run over b (coalesced) and the (r,k) logsumexp is an online-softmax merge across warps through shared
// memory; batch sums are warp-shuffle + smem block reductions.
#include "common.cuh"

namespace mmvae {


struct CombParams {
    const float *lpz, *lq, *lpx;
    float *lw, *loss_b, *w, *dlq;
    int64_t B;
    int M, L, K;
    float beta;
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
    for (int l = 0; l < p.L; ++l) v += __ldg(p.lpx + (((int64_t)r * p.L + l) * p.K + k) * p.B + b);
    return v - beta * lme_j(p, r, k, b, vals, mx, se);
}

// CTA: 32 consecutive batch rows (lanes, coalesced along b) x up to 32 warps that split the M*K (r,k) pairs.  The
// log-weights are parked in shared memory between the logsumexp pass and the weight pass; the (r,k) logsumexp is
// an online-softmax merge across warps.
__global__ void __launch_bounds__(1024) iwae_kernel(const CombParams p) {
    extern __shared__ float sm[];
    const int n = p.M * p.K;
    const int nw = blockDim.x >> 5;
    float* s_lw = sm;               // n x 32
    float* s_m = s_lw + n * 32;     // nw x 32
    float* s_s = s_m + nw * 32;     // nw x 32
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * 32 + lane;
    const bool ok = b < p.B;
    float vals[MMVAE_MAX_MODS];
    float run_m = -INFINITY, run_s = 0.f;
    if (ok) {
        for (int q = wid; q < n; q += nw) {
            const int r = q / p.K, k = q - r * p.K;
            float mx, se;
            const float lw = lw_value(p, r, k, b, p.beta, vals, mx, se);
            p.lw[((int64_t)r * p.K + k) * p.B + b] = lw;
            s_lw[q * 32 + lane] = lw;
            const float nm = fmaxf(run_m, lw);
            run_s = run_s * __expf(run_m - nm) + __expf(lw - nm);
            run_m = nm;
        }
    }
    s_m[wid * 32 + lane] = run_m;
    s_s[wid * 32 + lane] = run_s;
    __syncthreads();
    float tm = -INFINITY;
    for (int w = 0; w < nw; ++w) tm = fmaxf(tm, s_m[w * 32 + lane]);
    float ts = 0.f;
    for (int w = 0; w < nw; ++w)
        if (s_s[w * 32 + lane] > 0.f) ts += s_s[w * 32 + lane] * __expf(s_m[w * 32 + lane] - tm);
    const float lse = tm + logf(ts);
    if (!ok) return;
    if (wid == 0) p.loss_b[b] = -(lse - logf((float)n));
    for (int q = wid; q < n; q += nw) {
        const int r = q / p.K, k = q - r * p.K;
        const float wv = expf(s_lw[q * 32 + lane] - lse);
        p.w[((int64_t)r * p.K + k) * p.B + b] = wv;
        if (p.dlq) {
            float mx, se;
            lme_j(p, r, k, b, vals, mx, se);
            const float c = p.beta * wv / se;
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) p.dlq[(((int64_t)r * p.M + j) * p.K + k) * p.B + b] = c * expf(vals[j] - mx);
        }
    }
}

// DReG stage 1: one CTA per ((r,k), batch split); writes partial batch sums and softmax_j(lq) for the backward.
__global__ void __launch_bounds__(256) dreg_stage1_kernel(const CombParams p, double* __restrict__ part, int nsplit,
                                                          float* __restrict__ lq_soft) {
    __shared__ double red[32];
    const int q = blockIdx.x, sp = blockIdx.y;
    const int r = q / p.K, k = q - r * p.K;
    const int64_t per = (p.B + nsplit - 1) / nsplit;
    const int64_t b0 = sp * per, b1 = min(p.B, b0 + per);
    float vals[MMVAE_MAX_MODS];
    // batch sums feed a softmax over K: |lw| grows with B*P while the softmax needs its ABSOLUTE error small, and
    // the reference carries these sums in fp64 whenever the likelihood is lprob (objectives.py:422) -> double here
    double acc = 0.0;
    for (int64_t b = b0 + threadIdx.x; b < b1; b += blockDim.x) {
        float mx, se;
        acc += (double)lw_value(p, r, k, b, 1.0f, vals, mx, se);  // no beta in _m_dreg_looser (objectives.py:371)
        if (lq_soft) {
#pragma unroll
            for (int j = 0; j < MMVAE_MAX_MODS; ++j)
                if (j < p.M) lq_soft[(((int64_t)r * p.M + j) * p.K + k) * p.B + b] = expf(vals[j] - mx) / se;
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) part[(size_t)sp * p.M * p.K + q] = tot;
}

__global__ void dreg_partial_sum_kernel(const double* __restrict__ ws, int parts, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double tot = 0.0;
    for (int q = 0; q < parts; ++q) tot += ws[(size_t)q * n + i];
    out[i] = tot;
}

// DReG stage 2 (single CTA, one warp per modality row): wt = softmax_k(lw[r,:]); loss = -(1/M) sum wt*lw
__global__ void __launch_bounds__(256) dreg_stage2_kernel(const double* __restrict__ lw, int M, int K,
                                                          float* __restrict__ wt, float* __restrict__ loss) {
    __shared__ double s_part[8];
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

__global__ void __launch_bounds__(1024) reduce_sum_kernel(const float* __restrict__ x, int64_t n, float scale,
                                                          float* __restrict__ out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += x[i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) *out = scale * tot;
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

static int comb_fill(CombParams& p, const float* lpz, const float* lq, const float* lpx, int M, int L, int K,
                     int64_t B) {
    if (!lpz || !lq || (L > 0 && !lpx) || M <= 0 || L < 0 || K <= 0 || B <= 0) return MMVAE_E_ARG;
    if (M > MMVAE_MAX_MODS) return MMVAE_E_LIMIT;
    p.lpz = lpz; p.lq = lq; p.lpx = lpx; p.M = M; p.L = L; p.K = K; p.B = B;
    return 0;
}

extern "C" int mmvae_objective_iwae(const float* lpz, const float* lq, const float* lpx, int M, int L, int K,
                                    int64_t B, float beta, float* lw, float* loss_b, float* w, float* dlq,
                                    void* stream) {
    CombParams p{};
    int rc = comb_fill(p, lpz, lq, lpx, M, L, K, B);
    if (rc) return rc;
    if (!lw || !loss_b || !w) return MMVAE_E_ARG;
    p.beta = beta; p.lw = lw; p.loss_b = loss_b; p.w = w; p.dlq = dlq;
    const int n = M * K;
    int nw = n < 32 ? n : 32;
    const size_t smem = (size_t)(n * 32 + 2 * nw * 32) * sizeof(float);
    if (smem > 200 * 1024) return MMVAE_E_LIMIT;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(iwae_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    iwae_kernel<<<(unsigned)((B + 31) / 32), nw * 32, smem, (cudaStream_t)stream>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

#define DREG_MAX_SPLIT MMVAE_DREG_MAX_SPLIT
extern "C" int mmvae_objective_dreg_stage1(const float* lpz, const float* lq, const float* lpx, int M, int L, int K,
                                           int64_t B, double* lw_part, float* lq_soft, void* stream) {
    // lw_part: (DREG_MAX_SPLIT + 1, M*K) doubles: [0] receives the local batch sums, [1..] is scratch
    CombParams p{};
    int rc = comb_fill(p, lpz, lq, lpx, M, L, K, B);
    if (rc) return rc;
    if (!lw_part) return MMVAE_E_ARG;
    int nsplit = (int)((B + 2047) / 2048);
    if (nsplit < 1) nsplit = 1;
    if (nsplit > DREG_MAX_SPLIT) nsplit = DREG_MAX_SPLIT;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(M * K, nsplit);
    dreg_stage1_kernel<<<grid, 256, 0, st>>>(p, lw_part + (size_t)M * K, nsplit, lq_soft);
    MMVAE_LAUNCH_CHECK();
    dreg_partial_sum_kernel<<<(M * K + 127) / 128, 128, 0, st>>>(lw_part + (size_t)M * K, nsplit, M * K, lw_part);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_objective_dreg_stage2(const double* lw, int M, int K, float* wt, float* loss, void* stream) {
    if (!lw || !wt || !loss || M <= 0 || K <= 0) return MMVAE_E_ARG;
    dreg_stage2_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(lw, M, K, wt, loss);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_reduce_sum(const float* x, int64_t n, float scale, float* out, void* stream) {
    if (!x || !out || n <= 0) return MMVAE_E_ARG;
    reduce_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out);
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
