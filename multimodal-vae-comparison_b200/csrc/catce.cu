// category_ce rows: cross entropy with probability targets whose class axis is dim 1 (reference
// objectives.py:485-500; for (rows, T, 27) text the softmax runs over the SEQUENCE axis -- reproduced as is).
// One CTA per reconstruction row: the (C, d) slab of x and t is staged in shared memory with coalesced loads,
// each thread then owns a column j and walks the class axis; the row value is a warp-shuffle + smem block sum.
#include "common.cuh"

namespace mmvae {

struct CatceParams {
    const void* x;
    const void* t;
    void* g;
    const float* w_rows;
    float* out_rows;
    int64_t ldx, ldt, ldg, rows, B;
    int C, d;
    float lam, w_const;
};

template <typename TX, typename TT, int MODE>  // 0 fwd, 1 bwd, 2 fused
__global__ void __launch_bounds__(128) catce_kernel(const CatceParams p) {
    extern __shared__ float sm[];
    const int n = p.C * p.d;
    float* sx = sm;           // n
    float* st = sm + n;       // n
    float* s_lse = st + n;    // d
    float* s_tsum = s_lse + p.d;  // d
    __shared__ float red[32];

    const int64_t row = blockIdx.x;
    const TX* __restrict__ x = reinterpret_cast<const TX*>(p.x) + row * p.ldx;
    const TT* __restrict__ t = reinterpret_cast<const TT*>(p.t) + (row % p.B) * p.ldt;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        sx[e] = Elem<TX>::load1(x + e);
        st[e] = Elem<TT>::load1(t + e);
    }
    __syncthreads();
    float acc = 0.f;
    for (int j = threadIdx.x; j < p.d; j += blockDim.x) {
        float m = -INFINITY;
        for (int c = 0; c < p.C; ++c) m = fmaxf(m, sx[c * p.d + j]);
        float se = 0.f, ts = 0.f, txs = 0.f;
        for (int c = 0; c < p.C; ++c) {
            const float xv = sx[c * p.d + j], tv = st[c * p.d + j];
            se += expf(xv - m);
            ts += tv;
            txs += tv * xv;
        }
        const float lse = m + logf(se);
        s_lse[j] = lse;
        s_tsum[j] = ts;
        acc += txs - lse * ts;
    }
    if (MODE != 1) {
        const float tot = block_sum(acc, red);  // contains a __syncthreads -> s_lse / s_tsum visible below
        if (threadIdx.x == 0) p.out_rows[row] = p.lam * tot;
    } else {
        __syncthreads();
    }
    if (MODE != 0) {
        TX* __restrict__ g = reinterpret_cast<TX*>(p.g) + row * p.ldg;
        const float wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int j = e % p.d;
            const float sftm = expf(sx[e] - s_lse[j]);
            Elem<TX>::store1(g + e, wl * (st[e] - sftm * s_tsum[j]));
        }
    }
}

template <typename TX, typename TT>
static int launch_catce(int mode, const CatceParams& p, cudaStream_t st) {
    const size_t smem = (size_t)(2 * p.C * p.d + 2 * p.d) * sizeof(float);
    if (smem > 200 * 1024) return MMVAE_E_LIMIT;
    if (p.rows > 0x7fffffffLL) return MMVAE_E_LIMIT;
    auto k = mode == 0 ? catce_kernel<TX, TT, 0> : (mode == 1 ? catce_kernel<TX, TT, 1> : catce_kernel<TX, TT, 2>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    k<<<(unsigned)p.rows, 128, smem, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int mmvae_catce_rows(int mode, const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t C, int64_t d,
                                float lam, const float* w_rows, float w_const, float* out_rows, void* grad_recon,
                                int64_t ld_grad, void* stream) {
    if (!recon || !target || rows <= 0 || B <= 0 || C <= 0 || d <= 0) return MMVAE_E_ARG;
    if (mode < 0 || mode > 2) return MMVAE_E_ENUM;
    if (mode != 1 && !out_rows) return MMVAE_E_ARG;
    if (mode != 0 && !grad_recon) return MMVAE_E_ARG;
    if (mode == 1 && !w_rows) return MMVAE_E_ARG;
    if (ld_recon < C * d || ld_target < C * d || (mode != 0 && ld_grad < C * d)) return MMVAE_E_ARG;
    if (C * d > (1 << 20)) return MMVAE_E_LIMIT;
    CatceParams p{};
    p.x = recon; p.t = target; p.g = grad_recon; p.w_rows = w_rows; p.out_rows = out_rows;
    p.ldx = ld_recon; p.ldt = ld_target; p.ldg = ld_grad; p.rows = rows; p.B = B; p.C = (int)C; p.d = (int)d;
    p.lam = lam; p.w_const = w_const;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_F32) return launch_catce<float, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_F32) return launch_catce<__nv_bfloat16, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_BF16)
        return launch_catce<__nv_bfloat16, __nv_bfloat16>(mode, p, st);
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_BF16) return launch_catce<float, __nv_bfloat16>(mode, p, st);
    return MMVAE_E_ENUM;
}
