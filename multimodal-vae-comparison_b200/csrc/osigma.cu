// optimal_sigma (sigma-VAE) likelihood rows -- reference objectives.py:502-509 + utils.softclip utils.py:66-69.
// Three stream-ordered stages (global sum of squares -> rows -> gradient) so that a batch-sharded caller can
// all-reduce the single scalar in between.  Only log_sigma carries gradient: the squared term
// is detached in the reference.
#include "common.cuh"

namespace mmvae {

// VEC: fp32 / fp32, P % 4 == 0, 16-byte aligned rows: 128-bit loads (the scalar form reached 68 % of the HBM rate on the
// VILANRO shapes).  row_ss (may be NULL): the per-row sums of squares -- with them the row VALUES of stage 2 follow
// without a second pass over the reconstruction (r2: three passes over x -> two).
template <typename TX, typename TT, bool VEC>
__global__ void __launch_bounds__(256) osigma_sumsq_kernel(const TX* __restrict__ x, int64_t ldx,
                                                           const TT* __restrict__ t, int64_t ldt, int64_t rows,
                                                           int64_t B, int64_t P, double* sumsq, float* __restrict__ row_ss) {
    __shared__ double red[32];
    __shared__ float redf[32];
    double acc = 0.0;
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const TX* xr = x + row * ldx;
        const TT* tr = t + (row % B) * ldt;
        float part = 0.f;
        if (VEC) {
            for (int64_t i = (int64_t)threadIdx.x * 4; i < P; i += (int64_t)blockDim.x * 4) {
                const uint4 a = ldg_stream(xr + i), b = ldg_keep(tr + i);
                const float d0 = __uint_as_float(b.x) - __uint_as_float(a.x), d1 = __uint_as_float(b.y) - __uint_as_float(a.y);
                const float d2 = __uint_as_float(b.z) - __uint_as_float(a.z), d3 = __uint_as_float(b.w) - __uint_as_float(a.w);
                part += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
            }
        } else {
            for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
                const float d = Elem<TT>::load1(tr + i) - Elem<TX>::load1(xr + i);
                part += d * d;
            }
        }
        if (row_ss) {  // uniform branch
            const float rs = block_sum(part, redf);
            if (threadIdx.x == 0) {
                row_ss[row] = rs;
                acc += (double)rs;
            }
        } else {
            acc += (double)part;
        }
    }
    const double tot = block_sum(acc, red);
    if (threadIdx.x == 0) {
        atomicAdd(sumsq, tot);
        // element count of this call next to the sum, so that a batch-sharded caller all-reduces both in ONE collective
        // and uneven shards still get the global mean (n_total was rows*P*world, wrong for B % world != 0)
        if (blockIdx.x == 0) atomicAdd(sumsq + 1, (double)rows * (double)P);
    }
}

// stage 2 from the per-row sums of squares: row = -lam * (ss / sigma^2 + P * (log sigma + 0.5 log 2 pi)); no pass over x
__global__ void __launch_bounds__(256) osigma_rows_from_ss_kernel(const float* __restrict__ row_ss, int64_t rows, int64_t P,
                                                                  float lam, const double* sumsq, double n_total,
                                                                  float* __restrict__ out_rows, float* stats2);

__device__ __forceinline__ void osigma_stats(double sumsq, double n_total, float& log_sigma, float& dsoft) {
    // log_sigma = softclip(log sqrt(mean sq), -6) = -6 + softplus(u + 6)
    const float u = logf(sqrtf((float)(sumsq / n_total)));
    const float a = u + 6.0f;
    const float sp = a > 20.0f ? a : log1pf(expf(a));  // F.softplus threshold = 20
    log_sigma = -6.0f + sp;
    dsoft = a > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-a));
}

template <typename TX, typename TT>
__global__ void __launch_bounds__(256) osigma_rows_kernel(const TX* __restrict__ x, int64_t ldx,
                                                          const TT* __restrict__ t, int64_t ldt, int64_t rows,
                                                          int64_t B, int64_t P, float lam, const double* sumsq,
                                                          double n_total, float* out_rows, float* stats2) {
    __shared__ float red[32];
    float log_sigma, dsoft;
    osigma_stats(*sumsq, n_total > 0 ? n_total : sumsq[1], log_sigma, dsoft);
    const float inv_sigma = expf(-log_sigma);
    const float cst = log_sigma + 0.91893853320467274178f;  // + 0.5 log(2 pi)
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats2) {
        stats2[0] = log_sigma;
        stats2[1] = dsoft;
    }
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const TX* xr = x + row * ldx;
        const TT* tr = t + (row % B) * ldt;
        float acc = 0.f;
        for (int64_t i = threadIdx.x; i < P; i += blockDim.x) {
            const float d = (Elem<TT>::load1(tr + i) - Elem<TX>::load1(xr + i)) * inv_sigma;
            acc += d * d + cst;
        }
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) out_rows[row] = -lam * tot;
    }
}

__global__ void __launch_bounds__(256) osigma_rows_from_ss_kernel(const float* __restrict__ row_ss, int64_t rows, int64_t P,
                                                                  float lam, const double* sumsq, double n_total,
                                                                  float* __restrict__ out_rows, float* stats2) {
    float log_sigma, dsoft;
    osigma_stats(*sumsq, n_total > 0 ? n_total : sumsq[1], log_sigma, dsoft);
    const float inv_sigma = expf(-log_sigma);
    const float cst = log_sigma + 0.91893853320467274178f;
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats2) {
        stats2[0] = log_sigma;
        stats2[1] = dsoft;
    }
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
        out_rows[r] = -lam * fmaf(row_ss[r], inv_sigma * inv_sigma, (float)P * cst);
}

__global__ void __launch_bounds__(1024) sum_rows_kernel(const float* __restrict__ w, int64_t n, float* out) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += w[i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) *out = tot;
}

template <typename TX, typename TT, bool VEC>
__global__ void __launch_bounds__(256) osigma_bwd_kernel(const TX* __restrict__ x, int64_t ldx,
                                                         const TT* __restrict__ t, int64_t ldt, int64_t rows,
                                                         int64_t B, int64_t P, float lam, const double* sumsq,
                                                         double n_total, const float* wsum, TX* __restrict__ g,
                                                         int64_t ldg) {
    float log_sigma, dsoft;
    const double ss = *sumsq;
    osigma_stats(ss, n_total > 0 ? n_total : sumsq[1], log_sigma, dsoft);
    // d(sum_r w_r * row_r)/dx_i = -lam * P * wsum * dlog_sigma/dx_i,  dlog_sigma/dx_i = dsoft * (x_i - t_i) / sumsq
    const float coef = (float)(-(double)lam * (double)P * (double)(*wsum) * (double)dsoft / ss);
    for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const TX* xr = x + row * ldx;
        const TT* tr = t + (row % B) * ldt;
        TX* gr = g + row * ldg;
        if (VEC) {
            for (int64_t i = (int64_t)threadIdx.x * 4; i < P; i += (int64_t)blockDim.x * 4) {
                const uint4 a = ldg_stream(xr + i), b = ldg_keep(tr + i);
                stg_stream(gr + i, make_uint4(__float_as_uint(coef * (__uint_as_float(a.x) - __uint_as_float(b.x))),
                                              __float_as_uint(coef * (__uint_as_float(a.y) - __uint_as_float(b.y))),
                                              __float_as_uint(coef * (__uint_as_float(a.z) - __uint_as_float(b.z))),
                                              __float_as_uint(coef * (__uint_as_float(a.w) - __uint_as_float(b.w)))));
            }
        } else {
            for (int64_t i = threadIdx.x; i < P; i += blockDim.x)
                Elem<TX>::store1(gr + i, coef * (Elem<TX>::load1(xr + i) - Elem<TT>::load1(tr + i)));
        }
    }
}

static unsigned rows_grid(int64_t rows) {
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(rows < cap ? rows : cap);
}

#define OSIGMA_DISPATCH(CALL)                                                                                   \
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_F32) { CALL(float, float); }                          \
    else if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_F32) { CALL(__nv_bfloat16, float); }            \
    else if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_BF16) { CALL(__nv_bfloat16, __nv_bfloat16); }   \
    else if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_BF16) { CALL(float, __nv_bfloat16); }            \
    else return MMVAE_E_ENUM;

}  // namespace mmvae

using namespace mmvae;

static bool osigma_vec(const void* a, int64_t lda, const void* b, int64_t ldb, const void* c, int64_t ldc, int64_t P,
                       int dta, int dtb) {
    return dta == MMVAE_F32 && dtb == MMVAE_F32 && P % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 && aligned16(a) &&
           aligned16(b) && (!c || (ldc % 4 == 0 && aligned16(c)));
}

extern "C" int mmvae_osigma_sumsq(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                  int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P,
                                  double* sumsq, float* row_sumsq, void* stream) {
    if (!recon || !target || !sumsq || rows <= 0 || B <= 0 || P <= 0) return MMVAE_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (osigma_vec(recon, ld_recon, target, ld_target, nullptr, 0, P, dtype_recon, dtype_target)) {
        osigma_sumsq_kernel<float, float, true><<<rows_grid(rows), 256, 0, st>>>(
            (const float*)recon, ld_recon, (const float*)target, ld_target, rows, B, P, sumsq, row_sumsq);
        MMVAE_LAUNCH_CHECK();
        return 0;
    }
#define CALL(TX, TT) \
    osigma_sumsq_kernel<TX, TT, false><<<rows_grid(rows), 256, 0, st>>>((const TX*)recon, ld_recon, (const TT*)target, \
                                                                        ld_target, rows, B, P, sumsq, row_sumsq)
    OSIGMA_DISPATCH(CALL)
#undef CALL
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_osigma_fwd(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P, float lam,
                                const double* sumsq, double n_total, float* out_rows, float* stats2,
                                const float* row_sumsq, void* stream) {
    if (!recon || !target || !sumsq || !out_rows || rows <= 0 || B <= 0 || P <= 0) return MMVAE_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (row_sumsq) {  // per-row sums of squares from stage 1: no pass over the reconstruction
        int64_t grid = (rows + 255) / 256;
        if (grid > kNumSMs * 8) grid = kNumSMs * 8;
        osigma_rows_from_ss_kernel<<<(unsigned)grid, 256, 0, st>>>(row_sumsq, rows, P, lam, sumsq, n_total, out_rows, stats2);
        MMVAE_LAUNCH_CHECK();
        return 0;
    }
#define CALL(TX, TT)                                                                                              \
    osigma_rows_kernel<TX, TT><<<rows_grid(rows), 256, 0, st>>>((const TX*)recon, ld_recon, (const TT*)target,    \
                                                                ld_target, rows, B, P, lam, sumsq, n_total,       \
                                                                out_rows, stats2)
    OSIGMA_DISPATCH(CALL)
#undef CALL
    MMVAE_LAUNCH_CHECK();
    return 0;
}

extern "C" int mmvae_osigma_bwd(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P, float lam,
                                const double* sumsq, double n_total, const float* w_rows, float* wsum_scratch,
                                void* grad_recon, int64_t ld_grad, void* stream) {
    if (!recon || !target || !sumsq || !w_rows || !wsum_scratch || !grad_recon || rows <= 0 || B <= 0 || P <= 0)
        return MMVAE_E_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    sum_rows_kernel<<<1, 1024, 0, st>>>(w_rows, rows, wsum_scratch);
    MMVAE_LAUNCH_CHECK();
    if (osigma_vec(recon, ld_recon, target, ld_target, grad_recon, ld_grad, P, dtype_recon, dtype_target)) {
        osigma_bwd_kernel<float, float, true><<<rows_grid(rows), 256, 0, st>>>(
            (const float*)recon, ld_recon, (const float*)target, ld_target, rows, B, P, lam, sumsq, n_total, wsum_scratch,
            (float*)grad_recon, ld_grad);
        MMVAE_LAUNCH_CHECK();
        return 0;
    }
#define CALL(TX, TT)                                                                                                   \
    osigma_bwd_kernel<TX, TT, false><<<rows_grid(rows), 256, 0, st>>>((const TX*)recon, ld_recon, (const TT*)target,   \
                                                                      ld_target, rows, B, P, lam, sumsq, n_total,      \
                                                                      wsum_scratch, (TX*)grad_recon, ld_grad)
    OSIGMA_DISPATCH(CALL)
#undef CALL
    MMVAE_LAUNCH_CHECK();
    return 0;
}
