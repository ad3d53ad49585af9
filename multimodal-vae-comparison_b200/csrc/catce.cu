// category_ce rows: cross entropy with probability targets whose class axis is dim 1 (reference
// objectives.py:485-500; for (rows, T, 27) text the softmax runs over the SEQUENCE axis -- reproduced as is).
//
// Rows are tiny (C*d = 9 ... 6642 elements), so the kernel is organised around staging, not arithmetic:
//   * a CTA owns R consecutive rows; their (C, d) slabs of x and t are ONE contiguous run in HBM, brought into
//     shared memory by a single TMA bulk copy each (cp.async.bulk + mbarrier) whenever the run is 16-byte aligned
//     (R is chosen so that R*C*d*elemsize is a multiple of 16); otherwise by coalesced element loads;
//   * W warps per row split the class axis (W = 1 for short class axes: then there is no block barrier on the compute
//     path at all); lanes own columns j, so the smem walk along the class axis is conflict free;
//   * (measured r1: staging only x and reading the L2-resident targets from global halves the smem footprint but
//     slowed the streaming backward from 21.7 to 35.6 us and left the forward unchanged -- both slabs stay staged)
//   * two passes over the class axis (max, then exp-sum / target sums: one MUFU.EX2 per element), warp-shuffle row
//     sum; the gradient is written in place over the staged x and leaves through one TMA bulk store.
#include "common.cuh"

namespace mmvae {

struct CatceParams {
    const void* x;
    const void* t;
    void* g;
    const float* w_rows;
    float* out_rows;
    float* stats;  // (rows, 2, d): logsumexp and target sum per column; written by fwd, read by bwd (may be NULL)
    int64_t ldx, ldt, ldg, rows, B;
    int C, d, R, W, tma;
    float lam, w_const;
};

__host__ __device__ __forceinline__ size_t up16(size_t v) { return (v + 15) & ~(size_t)15; }

// n elements global -> shared with 4-byte cp.async when both addresses are 4-byte aligned (always for fp32; for bf16
// when the row starts on an even element), element-wise otherwise.
template <typename T>
__device__ __forceinline__ void stage_row_async(T* sdst, const T* gsrc, int n, int tid, int nthreads) {
    constexpr int per = 4 / (int)sizeof(T);  // elements per 4-byte word
    const bool ok = ((reinterpret_cast<uintptr_t>(gsrc) | reinterpret_cast<uintptr_t>(sdst)) & 3u) == 0;
    if (ok) {
        const int words = n / per;
        const uint32_t s0 = smem_u32(sdst);
        for (int w = tid; w < words; w += nthreads)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s0 + 4u * w),
                         "l"(reinterpret_cast<const char*>(gsrc) + 4 * (size_t)w)
                         : "memory");
        for (int e = words * per + tid; e < n; e += nthreads) sdst[e] = gsrc[e];
    } else {
        for (int e = tid; e < n; e += nthreads) sdst[e] = gsrc[e];
    }
}

template <typename TX, typename TT, int MODE>  // 0 fwd, 1 bwd, 2 fused
__global__ void __launch_bounds__(512) catce_kernel(const CatceParams p) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int n = p.C * p.d, R = p.R, W = p.W;
    TX* sx = reinterpret_cast<TX*>(smraw);
    size_t off = up16((size_t)R * n * sizeof(TX));
    TT* st = reinterpret_cast<TT*>(smraw + off);
    off += up16((size_t)R * n * sizeof(TT));
    float* s_lse = reinterpret_cast<float*>(smraw + off);  // R*d
    float* s_ts = s_lse + R * p.d;                          // R*d
    float* part = s_ts + R * p.d;                           // R*W*d*4 (only touched when W > 1)
    uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + up16(off + (size_t)(2 * R * p.d + 4 * R * W * p.d) * 4));

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = warp / W, w = warp - rl * W;  // local row, class slice
    const int64_t row0 = (int64_t)blockIdx.x * R;
    const int nrows = (int)min((int64_t)R, p.rows - row0);
    const TX* __restrict__ xg = reinterpret_cast<const TX*>(p.x);
    const TT* __restrict__ tg = reinterpret_cast<const TT*>(p.t);

    // ---- stage R rows of x and t ----------------------------------------------------------------------------
    if (p.tma) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bx = (uint32_t)((size_t)R * n * sizeof(TX)), bt = (uint32_t)((size_t)R * n * sizeof(TT));
            mbar_expect_tx(bar, bx + bt);
            bulk_g2s(sx, xg + row0 * p.ldx, bx, bar);
            bulk_g2s(st, tg + (row0 % p.B) * p.ldt, bt, bar);
        }
        // ONE warp polls the mbarrier; the others park on the CTA barrier (a spinning try_wait loop in every warp
        // burned 3x more issue slots than the arithmetic of the whole kernel -- ncu r1: 18.7 M warp instructions)
        if (warp == 0) mbar_wait(bar, 0);
        __syncthreads();
    } else {
        // rows that miss TMA's 16-byte alignment: 4-byte cp.async (LDGSTS) -- asynchronous and register free, so a
        // CTA keeps its whole slab in flight (scalar staging loads left the SM with ~16 KB in flight: ncu r1,
        // long-scoreboard bound at 1.5 TB/s on the CUB caption rows); plain loads only for odd 2-byte rows
        for (int r = 0; r < nrows; ++r) {
            const TX* xr = xg + (row0 + r) * p.ldx;
            const TT* tr = tg + ((row0 + r) % p.B) * p.ldt;
            stage_row_async(sx + (size_t)r * n, xr, n, threadIdx.x, blockDim.x);
            stage_row_async(st + (size_t)r * n, tr, n, threadIdx.x, blockDim.x);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    const bool live = rl < nrows;
    const TX* rx = sx + (size_t)rl * n;
    const TT* rt = st + (size_t)rl * n;
    float acc = 0.f;

    if (MODE == 1 && p.stats) {
        // backward with the column statistics cached by the forward: one streaming pass, no reductions.
        // (class-row, column) walk: warps stride over the R*C class rows, lanes over the columns.
        for (int i = threadIdx.x; i < nrows * 2 * p.d; i += blockDim.x) s_lse[i] = __ldg(p.stats + row0 * 2 * p.d + i);
        __syncthreads();
        const int nwarps = blockDim.x >> 5;
        for (int r = 0; r < nrows; ++r) {
            const float wl = __ldg(p.w_rows + row0 + r) * p.lam;
            const float* sl = s_lse + r * 2 * p.d;
            TX* gr = reinterpret_cast<TX*>(p.g) + (row0 + r) * p.ldg;
            for (int j = lane; j < p.d; j += 32) {
                const float nl = -sl[j] * kLog2e, wts = -wl * sl[p.d + j];
                for (int c = warp; c < p.C; c += nwarps) {
                    const int e = r * n + c * p.d + j;
                    const float gv = fmaf(exp_shifted(Elem<TX>::get(sx + e), nl), wts, wl * Elem<TT>::get(st + e));
                    if (p.tma)
                        Elem<TX>::store1(sx + e, gv);
                    else
                        Elem<TX>::store1(gr + c * p.d + j, gv);
                }
            }
        }
        if (p.tma) {
            fence_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(reinterpret_cast<TX*>(p.g) + row0 * p.ldg, sx, (uint32_t)((size_t)R * n * sizeof(TX)));
                bulk_wait_read();
            }
        }
        return;
    }

    if (W == 1) {  // one warp owns the whole row: no block barrier on the compute path
        if (live) {
            for (int j = lane; j < p.d; j += 32) {
                float m = -INFINITY;
                const TX* px = rx + j;
                const TT* pt = rt + j;
#pragma unroll 5
                for (int c = 0; c < p.C; ++c) m = fmaxf(m, Elem<TX>::get(px + c * p.d));
                float se = 0.f, ts = 0.f, txs = 0.f;
                const float nm = -m * kLog2e;
#pragma unroll 5
                for (int c = 0; c < p.C; ++c) {
                    const float xv = Elem<TX>::get(px + c * p.d), tv = Elem<TT>::get(pt + c * p.d);
                    se += exp_shifted(xv, nm);
                    ts += tv;
                    txs = fmaf(tv, xv, txs);
                }
                const float lse = m + logf(se);
                s_lse[rl * p.d + j] = lse;
                s_ts[rl * p.d + j] = ts;
                acc += txs - lse * ts;
            }
            __syncwarp();
        }
    } else {  // W warps split the class axis of a row; merged through shared memory
        const int step = W * p.d;  // running pointers: one IADD per access instead of an IMAD (ncu r1: 94 thread
                                   // instructions per element on the CUB caption rows, issue bound)
        if (live)
            for (int j = lane; j < p.d; j += 32) {
                float m = -INFINITY;
                const TX* px = rx + w * p.d + j;
#pragma unroll 4
                for (int c = w; c < p.C; c += W, px += step) m = fmaxf(m, Elem<TX>::get(px));
                part[((size_t)(rl * W + w) * p.d + j) * 4] = m;
            }
        __syncthreads();
        if (live)
            for (int j = lane; j < p.d; j += 32) {
                float m = -INFINITY;
                for (int ww = 0; ww < W; ++ww) m = fmaxf(m, part[((size_t)(rl * W + ww) * p.d + j) * 4]);
                float se = 0.f, ts = 0.f, txs = 0.f;
                const float nm = -m * kLog2e;
                const TX* px = rx + w * p.d + j;
                const TT* pt = rt + w * p.d + j;
#pragma unroll 4
                for (int c = w; c < p.C; c += W, px += step, pt += step) {
                    const float xv = Elem<TX>::get(px), tv = Elem<TT>::get(pt);
                    se += exp_shifted(xv, nm);
                    ts += tv;
                    txs = fmaf(tv, xv, txs);
                }
                float* q = part + ((size_t)(rl * W + w) * p.d + j) * 4;
                q[1] = se;
                q[2] = ts;
                q[3] = txs;
            }
        __syncthreads();
        if (live && w == 0)
            for (int j = lane; j < p.d; j += 32) {
                float mm = -INFINITY;  // every slice summed exp(x - mm) against the same global max
                for (int ww = 0; ww < W; ++ww) mm = fmaxf(mm, part[((size_t)(rl * W + ww) * p.d + j) * 4]);
                float se = 0.f, ts = 0.f, txs = 0.f;
                for (int ww = 0; ww < W; ++ww) {
                    const float* q = part + ((size_t)(rl * W + ww) * p.d + j) * 4;
                    se += q[1];
                    ts += q[2];
                    txs += q[3];
                }
                const float lse = mm + logf(se);
                s_lse[rl * p.d + j] = lse;
                s_ts[rl * p.d + j] = ts;
                acc += txs - lse * ts;
            }
        __syncthreads();
    }
    if (MODE != 1 && live && w == 0) {
        acc = warp_sum(acc);
        if (lane == 0) p.out_rows[row0 + rl] = p.lam * acc;
        if (p.stats)
            for (int j = lane; j < p.d; j += 32) {
                p.stats[(row0 + rl) * 2 * p.d + j] = s_lse[rl * p.d + j];
                p.stats[(row0 + rl) * 2 * p.d + p.d + j] = s_ts[rl * p.d + j];
            }
    }
    if (MODE != 0) {
        TX* gr = reinterpret_cast<TX*>(p.g) + (row0 + rl) * p.ldg;
        TX* gs = sx + (size_t)rl * n;  // in-place staging of the gradient when it leaves through TMA
        if (live) {
            const float wl = (p.w_rows ? __ldg(p.w_rows + row0 + rl) : p.w_const) * p.lam;
            const int step = W * p.d;
            TX* gdst = p.tma ? gs : gr;  // in-place smem staging (leaves through TMA) or straight to global
            for (int j = lane; j < p.d; j += 32) {
                const float nl = -s_lse[rl * p.d + j] * kLog2e, wts = -wl * s_ts[rl * p.d + j];
                const TX* px = rx + w * p.d + j;
                const TT* pt = rt + w * p.d + j;
                TX* pg = gdst + w * p.d + j;
#pragma unroll 4
                for (int c = w; c < p.C; c += W, px += step, pt += step, pg += step)  // wl*t - wl*ts*softmax
                    Elem<TX>::store1(pg, fmaf(exp_shifted(Elem<TX>::get(px), nl), wts, wl * Elem<TT>::get(pt)));
            }
        }
        if (p.tma) {
            fence_async_smem();  // generic-proxy smem writes -> visible to the async (TMA) proxy
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(reinterpret_cast<TX*>(p.g) + row0 * p.ldg, sx, (uint32_t)((size_t)R * n * sizeof(TX)));
                bulk_wait_read();  // smem must stay alive until the bulk store has read it
            }
        }
    }
}

static size_t catce_smem(int R, int W, int n, int d, int sx, int st) {
    size_t off = up16((size_t)R * n * sx) + up16((size_t)R * n * st);
    off = up16(off + (size_t)(2 * R * d + 4 * R * W * d) * 4);
    return off + 16;
}

template <typename TX, typename TT>
static int launch_catce(int mode, CatceParams p, cudaStream_t st) {
    const int n = p.C * p.d, sx = (int)sizeof(TX), stt = (int)sizeof(TT);
    const bool fast_bwd = (mode == 1 && p.stats != nullptr);
    // warps per row split the class axis (short dependency chains); the cached-statistics backward has no per-row
    // reduction and simply strides all warps over the staged class rows
    // (measured r1, C = 45: W = 1 -> 17 us, W = 4 -> 27 us: the merge costs more than the shorter chains save)
    int W = fast_bwd ? 1 : (p.C >= 192 ? 8 : (p.C >= 96 ? 4 : 1));
    int R = fast_bwd ? 4 : (W == 1 ? 8 : 16 / W);
    // a CTA's life is TMA latency + a short compute phase: favour many resident CTAs (<= 48 KB each) ...
    while (R > 1 && catce_smem(R, W, n, p.d, sx, stt) > 48 * 1024) R >>= 1;
    // ... but keep the staged run a multiple of 16 bytes (TMA-able) when that still leaves 2 CTAs per SM
    auto tma_ok = [&](int r) { return ((size_t)r * n * sx) % 16 == 0 && ((size_t)r * n * stt) % 16 == 0; };
    for (int r2 = R; r2 <= 8 && !tma_ok(R); r2 <<= 1)
        if (tma_ok(r2) && catce_smem(r2, W, n, p.d, sx, stt) <= 110 * 1024 && (fast_bwd || r2 * W <= 16)) R = r2;
    if (catce_smem(R, W, n, p.d, sx, stt) > 200 * 1024) return MMVAE_E_LIMIT;
    bool dense = p.ldx == n && p.ldt == n && (mode == 0 || p.ldg == n);
    bool tma = dense && aligned16(p.x) && aligned16(p.t) && (mode == 0 || aligned16(p.g)) && p.rows % R == 0 &&
               p.B % R == 0 && ((size_t)R * n * sx) % 16 == 0 && ((size_t)R * n * stt) % 16 == 0;
    p.R = R;
    p.W = W;
    p.tma = tma ? 1 : 0;
    const size_t smem = catce_smem(R, W, n, p.d, sx, stt);
    const int64_t grid = (p.rows + R - 1) / R;
    if (grid > 0x7fffffffLL) return MMVAE_E_LIMIT;
    auto k = mode == 0 ? catce_kernel<TX, TT, 0> : (mode == 1 ? catce_kernel<TX, TT, 1> : catce_kernel<TX, TT, 2>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    const int threads = fast_bwd ? 256 : R * W * 32;
    k<<<(unsigned)grid, threads, smem, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int mmvae_catce_rows(int mode, const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t C, int64_t d,
                                float lam, const float* w_rows, float w_const, float* out_rows, void* grad_recon,
                                int64_t ld_grad, float* stats, void* stream) {
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
    p.lam = lam; p.w_const = w_const; p.stats = stats;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_F32) return launch_catce<float, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_F32) return launch_catce<__nv_bfloat16, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_BF16)
        return launch_catce<__nv_bfloat16, __nv_bfloat16>(mode, p, st);
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_BF16) return launch_catce<float, __nv_bfloat16>(mode, p, st);
    return MMVAE_E_ENUM;
}
