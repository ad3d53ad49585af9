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
#include <type_traits>

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
    int G;               // rows a warp works on at once (32 / d for d < 32, else 1): lane -> (row in group, column)
    int S, TS;           // ring kernel: x stages, target slots
    int npw, lg_ns, same_rows;  // resident kernel: periods per warp, log2 of the merge segment, rows == B
    const unsigned char* mask;  // (B, C) bytes or NULL: padding mask of the text decoder, fused multiply (masked kernel)
    int64_t ldm;
    float inv_n, inv_d;  // v2 flat backward: 1/(C*d), 1/d for the exact float-reciprocal index split
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


// ---- register-chunked walk along the class axis ---------------------------------------------------------------
// A lane owns one column j and walks the class axis in chunks of CH values held in registers: CH independent loads
// per tensor issued back to back (smem: ~30 cycles each, global: one DRAM round trip for the whole chunk), then an
// online softmax update (running max, rescaled sum) with TWO accumulators per sum.  ncu r1 on the element-at-a-time
// loops this replaces: every class paid LDS -> FFMA -> MUFU -> FADD in sequence (~57 cycles, 45 times per row), the
// staged kernels were bound by that dependency chain at 2-5 warps per scheduler, not by memory.
#ifndef MMVAE_CATCE_CH
#define MMVAE_CATCE_CH 16  // global-memory column kernel
#endif
#ifndef MMVAE_CATCE_SCH
#define MMVAE_CATCE_SCH 15  // smem kernels (C = 45 -> three full chunks)
#endif

template <typename TX, typename TT, int CH, bool FULL, bool GLOBAL>
__device__ __forceinline__ void cols_load(const TX* px, const TT* pt, int c0, int C, int d, float* xv, float* tv) {
#pragma unroll
    for (int u = 0; u < CH; ++u)
        xv[u] = (FULL || c0 + u < C) ? (GLOBAL ? Elem<TX>::load1(px + (c0 + u) * d) : Elem<TX>::get(px + (c0 + u) * d))
                                     : -INFINITY;
#pragma unroll
    for (int u = 0; u < CH; ++u)
        tv[u] = (FULL || c0 + u < C) ? (GLOBAL ? Elem<TT>::load1(pt + (c0 + u) * d) : Elem<TT>::get(pt + (c0 + u) * d))
                                     : 0.f;
}

struct ColAcc {  // running column statistics: max, sum exp(x - max), sum t, sum t*x (two partial sums each)
    float m = -INFINITY, se0 = 0.f, se1 = 0.f, ts0 = 0.f, ts1 = 0.f, tx0 = 0.f, tx1 = 0.f;
    __device__ __forceinline__ float se() const { return se0 + se1; }
    __device__ __forceinline__ float ts() const { return ts0 + ts1; }
    __device__ __forceinline__ float txs() const { return tx0 + tx1; }
};

template <int CH, bool FULL>
__device__ __forceinline__ void cols_accum(const float* xv, const float* tv, int c0, int C, ColAcc& a) {
    float cm = xv[0];
#pragma unroll
    for (int u = 1; u < CH; ++u) cm = fmaxf(cm, xv[u]);
    const float mn = fmaxf(a.m, cm);
    const float nm = (mn == -INFINITY) ? 0.f : -mn * kLog2e;  // a column of -inf so far keeps the sums at 0
    const float rs = ex2_ftz(fmaf(a.m, kLog2e, nm));           // rescale the running sum to the new max
    a.se0 *= rs;
    a.se1 *= rs;
#pragma unroll
    for (int u = 0; u < CH; ++u) {
        const float e = ex2_ftz(fmaf(xv[u], kLog2e, nm));
        if (u & 1) {
            a.se1 += e;
            a.ts1 += tv[u];
            if (FULL || c0 + u < C) a.tx1 = fmaf(tv[u], xv[u], a.tx1);
        } else {
            a.se0 += e;
            a.ts0 += tv[u];
            if (FULL || c0 + u < C) a.tx0 = fmaf(tv[u], xv[u], a.tx0);
        }
    }
    a.m = mn;
}

// running statistics of C classes of one column; px / pt point at the first class, consecutive classes d apart
template <typename TX, typename TT, int CH, bool GLOBAL>
__device__ __forceinline__ ColAcc col_partial(const TX* px, const TT* pt, int C, int d) {
    ColAcc a;
    int c0 = 0;
    for (; c0 + CH <= C; c0 += CH) {
        float xv[CH], tv[CH];
        cols_load<TX, TT, CH, true, GLOBAL>(px, pt, c0, C, d, xv, tv);
        cols_accum<CH, true>(xv, tv, c0, C, a);
    }
    if (c0 < C) {
        float xv[CH], tv[CH];
        cols_load<TX, TT, CH, false, GLOBAL>(px, pt, c0, C, d, xv, tv);
        cols_accum<CH, false>(xv, tv, c0, C, a);
    }
    return a;
}

// (logsumexp, sum t, sum t*x) of one column; px / pt point at class 0 of the column, consecutive classes d apart
template <typename TX, typename TT, int CH, bool GLOBAL>
__device__ __forceinline__ void col_stats(const TX* px, const TT* pt, int C, int d, float& lse, float& ts, float& txs) {
    ColAcc a;
    int c0 = 0;
    for (; c0 + CH <= C; c0 += CH) {
        float xv[CH], tv[CH];
        cols_load<TX, TT, CH, true, GLOBAL>(px, pt, c0, C, d, xv, tv);
        cols_accum<CH, true>(xv, tv, c0, C, a);
    }
    if (c0 < C) {
        float xv[CH], tv[CH];
        cols_load<TX, TT, CH, false, GLOBAL>(px, pt, c0, C, d, xv, tv);
        cols_accum<CH, false>(xv, tv, c0, C, a);
    }
    lse = a.m + logf(a.se());
    ts = a.ts();
    txs = a.txs();
}

// gradient of one column, chunked the same way: g[c] = wl*t[c] - wl*ts*exp(x[c] - lse); pg may alias px (in place)
template <typename TX, typename TT, int CH, bool GLOBAL>
__device__ __forceinline__ void col_grad(const TX* px, const TT* pt, TX* pg, int C, int d, float lse, float ts, float wl) {
    const float nl = -lse * kLog2e, wts = -wl * ts;
    for (int c0 = 0; c0 < C; c0 += CH) {
        float xv[CH], tv[CH];
        if (c0 + CH <= C) cols_load<TX, TT, CH, true, GLOBAL>(px, pt, c0, C, d, xv, tv);
        else cols_load<TX, TT, CH, false, GLOBAL>(px, pt, c0, C, d, xv, tv);
#pragma unroll
        for (int u = 0; u < CH; ++u)
            if (c0 + u < C) Elem<TX>::store1(pg + (c0 + u) * d, fmaf(ex2_ftz(fmaf(xv[u], kLog2e, nl)), wts, wl * tv[u]));
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
                float lse, ts, txs;
                col_stats<TX, TT, MMVAE_CATCE_SCH, false>(rx + j, rt + j, p.C, p.d, lse, ts, txs);
                s_lse[rl * p.d + j] = lse;
                s_ts[rl * p.d + j] = ts;
                acc += txs - lse * ts;
            }
            __syncwarp();
        }
    } else {
        // W warps split the class axis of a row into contiguous slices: partial (max, sum exp, sum t, sum t*x) per
        // slice and column -> smem -> ONE barrier -> every warp merges the W partials of its columns (4 W loads) and
        // goes straight on to the gradient of its own slice.  (r1: the strided element-at-a-time version with three
        // barriers ran the CUB caption rows, C = 246, at 2.9 TB/s through the 4-byte cp.async path.)
        const int Cw = (p.C + W - 1) / W, c_lo = w * Cw;
        const int c_n = max(0, min(p.C, c_lo + Cw) - c_lo);
        if (live)
            for (int j = lane; j < p.d; j += 32) {
                const ColAcc a = col_partial<TX, TT, MMVAE_CATCE_SCH, false>(rx + c_lo * p.d + j, rt + c_lo * p.d + j, c_n, p.d);
                float* q = part + ((size_t)(rl * W + w) * p.d + j) * 4;
                q[0] = a.m;
                q[1] = a.se();
                q[2] = a.ts();
                q[3] = a.txs();
            }
        __syncthreads();
        if (live) {
            float wl = 0.f;
            if (MODE != 0) wl = (p.w_rows ? __ldg(p.w_rows + row0 + rl) : p.w_const) * p.lam;
            TX* gdst = p.tma ? sx + (size_t)rl * n : reinterpret_cast<TX*>(p.g) + (row0 + rl) * p.ldg;
            for (int j = lane; j < p.d; j += 32) {
                float mm = -INFINITY;
                for (int ww = 0; ww < W; ++ww) mm = fmaxf(mm, part[((size_t)(rl * W + ww) * p.d + j) * 4]);
                const float nm = (mm == -INFINITY) ? 0.f : -mm * kLog2e;
                float se = 0.f, ts = 0.f, txs = 0.f;
                for (int ww = 0; ww < W; ++ww) {
                    const float* q = part + ((size_t)(rl * W + ww) * p.d + j) * 4;
                    se = fmaf(q[1], ex2_ftz(fmaf(q[0], kLog2e, nm)), se);  // rescale each slice to the global max
                    ts += q[2];
                    txs += q[3];
                }
                const float lse = mm + logf(se);
                if (w == 0) {
                    s_lse[rl * p.d + j] = lse;
                    s_ts[rl * p.d + j] = ts;
                    acc += txs - lse * ts;
                }
                if (MODE != 0)
                    col_grad<TX, TT, MMVAE_CATCE_SCH, false>(rx + c_lo * p.d + j, rt + c_lo * p.d + j, gdst + c_lo * p.d + j,
                                                             c_n, p.d, lse, ts, wl);
            }
        }
        __syncwarp();
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
            TX* gdst = p.tma ? gs : gr;  // in-place smem staging (leaves through TMA) or straight to global
            if (W == 1)  // (W > 1: the gradient of every slice was written right after the merge above)
                for (int j = lane; j < p.d; j += 32)
                    col_grad<TX, TT, MMVAE_CATCE_SCH, false>(rx + j, rt + j, gdst + j, p.C, p.d, s_lse[rl * p.d + j],
                                                             s_ts[rl * p.d + j], wl);
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


// ---------------------------------------------------------------------------------------------------------------
// bf16 pair kernel (r2): long class axes in bf16 (CUB captions, C = 246, d = 27 -- BASELINE configs[4]).
//
// ncu r1 on the staged kernel above at that shape: instruction-latency bound (issue-active 45 %), 2.2 TB/s.  A lane
// there owns one column and pays, per element, a 2-byte LDS per tensor, a conversion per tensor and scalar math.  Here a
// lane owns one 32-bit WORD position of the class-row period, i.e. TWO element streams, and everything is done on pairs:
// one LDS.32 per tensor and pair, the two bf16 -> fp32 conversions are a shift and a mask into an aligned register pair,
// exp arguments / running sums / gradient are packed f32x2 operations (common.cuh), the gradient leaves as one packed
// cvt.rn.bf16x2 + STS.32.
//   d even: a period is one class row (d/2 words); a lane's streams are the columns (2p, 2p+1), every class row.
//   d odd : a period is TWO class rows (d words; needs C even); the streams are (parity, column) pairs -- columns 2p, 2p+1
//           of the even rows for p < (d-1)/2, the straddling word (even row, d-1 | odd row, 0), then columns of the odd
//           rows.  The two parities of a column are merged like the W class-axis slices of the warps.
// W warps per row split the periods; partial (max, sum exp, sum t, sum t*x) per (warp, parity, column) go through
// shared memory, ONE barrier, every warp merges the statistics of all columns itself and proceeds to the gradient of
// its own slice (written in place over the staged x, one TMA bulk store per CTA).  Targets are staged next to x or read
// straight from global memory (STAGE_T = 0: half the shared memory per CTA, more resident warps).
// ---------------------------------------------------------------------------------------------------------------
#ifndef MMVAE_CATCE_PAIRS
#define MMVAE_CATCE_PAIRS 1
#endif
#ifndef MMVAE_CATCE_PAIRS_W
#define MMVAE_CATCE_PAIRS_W 4
#endif
#ifndef MMVAE_CATCE_PAIRS_R
#define MMVAE_CATCE_PAIRS_R 4
#endif
#ifndef MMVAE_CATCE_PAIRS_STAGE_T
#define MMVAE_CATCE_PAIRS_STAGE_T 0
#endif
// STAGE_X = 0 (r2, measured and rejected; kept as a build option): x not staged either -- the words of a lane's streams
// come straight from global memory in register chunks (pass 1 allocates in L1, pass 2 re-reads the slice from L1 / L2),
// the gradient goes straight back, ~15 KB of shared memory per CTA.  C5 captions (4096 x 246 x 27 bf16), fused:
// 63.9 us (R4 W4) / 60.2 us (R4 W2) against 59.8 us for the TMA-staged form -- neither the TMA round trip nor the
// occupancy is the limit.  The SASS is: 16 (pass 1) / 22 (pass 2) instructions per element, 57 % of them 64-bit
// address arithmetic and bounds predicates of the strided word accesses (run-time period length).  Hence DT below.
#ifndef MMVAE_CATCE_PAIRS_STAGE_X
#define MMVAE_CATCE_PAIRS_STAGE_X 1
#endif
#ifndef MMVAE_CATCE_RESIDENT
#define MMVAE_CATCE_RESIDENT 1
#endif
#ifndef MMVAE_CATCE_RESIDENT_OCC
#define MMVAE_CATCE_RESIDENT_OCC 3  // CTAs of 256 threads per SM the register allocation aims for
#endif
#ifndef MMVAE_CATCE_RESIDENT_MINW
#define MMVAE_CATCE_RESIDENT_MINW 1
#endif
#ifndef MMVAE_CATCE_PAIRS_CH
#define MMVAE_CATCE_PAIRS_CH 8
#endif
constexpr int kPairCh = MMVAE_CATCE_PAIRS_CH;  // periods per register chunk: words per tensor in flight per lane

__device__ __forceinline__ f32x2 bf2_to_f2(uint32_t w) { return f2_pack(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }
__device__ __forceinline__ uint32_t f2_to_bf2(f32x2 v) {
    uint32_t r;
    float lo, hi;
    f2_unpack(v, lo, hi);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// DT: compile-time d (27 = the character vocabulary of every text modality of the reference, datasets.py / config_cub;
// 0 = run-time d).  With it the period length is an immediate: the word accesses of a register chunk become
// [pointer + constant] operands of one running pointer per tensor instead of a 64-bit multiply-add each.
template <int MODE, bool STAGE_T, bool STAGE_X, int DT>  // MODE 0: forward (+ statistics), 2: fused value + gradient
__global__ void __launch_bounds__(512) catce_pairs_kernel(const CatceParams p) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int d = DT ? DT : p.d;
    const int n = p.C * d, R = p.R, W = p.W;
    const bool odd = (d & 1) != 0;
    const int Pw = odd ? d : d / 2;        // words per period
    const int NP = odd ? p.C / 2 : p.C;    // periods per row
    const int NPAR = odd ? 2 : 1;
    uint32_t* sx = reinterpret_cast<uint32_t*>(smraw);
    size_t off = STAGE_X ? up16((size_t)R * n * 2) : 0;
    uint32_t* stg = reinterpret_cast<uint32_t*>(smraw + off);
    if (STAGE_T) off += up16((size_t)R * n * 2);
    float* s_lse = reinterpret_cast<float*>(smraw + off);  // R*d merged logsumexp
    float* s_ts = s_lse + R * d;                            // R*d merged sum t
    float4* part = reinterpret_cast<float4*>(smraw + up16(off + (size_t)2 * R * d * 4));  // R*W*NPAR*d
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(part) + (size_t)R * W * NPAR * d * 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rl = warp / W, w = warp - rl * W;
    const int64_t row0 = (int64_t)blockIdx.x * R;
    const __nv_bfloat16* xg = reinterpret_cast<const __nv_bfloat16*>(p.x);
    const __nv_bfloat16* tg = reinterpret_cast<const __nv_bfloat16*>(p.t);
    if (STAGE_X || STAGE_T) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bx = (uint32_t)((size_t)R * n * 2);
            mbar_expect_tx(bar, (STAGE_T ? bx : 0) + (STAGE_X ? bx : 0));
            if (STAGE_X) bulk_g2s(sx, xg + row0 * p.ldx, bx, bar);
            if (STAGE_T) bulk_g2s(stg, tg + (row0 % p.B) * p.ldt, bx, bar);
        }
    }
    // this lane's two streams: (parity, column) of the low and the high half of its word
    const bool lane_ok = lane < Pw;
    int parL = 0, colL = 2 * lane, parH = 0, colH = 2 * lane + 1;
    if (odd) {
        const int h = (d - 1) / 2;
        if (lane == h) { colL = d - 1; parH = 1; colH = 0; }
        else if (lane > h) { parL = parH = 1; colL = 2 * lane - d; colH = colL + 1; }
    }
    const int NPw = (NP + W - 1) / W, q_lo = w * NPw, q_hi = min(NP, q_lo + NPw);
    const uint32_t* rx = STAGE_X ? sx + (size_t)rl * (n / 2) + lane
                                 : reinterpret_cast<const uint32_t*>(xg + (row0 + rl) * p.ldx) + lane;
    const uint32_t* rt = STAGE_T ? stg + (size_t)rl * (n / 2) + lane
                                 : reinterpret_cast<const uint32_t*>(tg + ((row0 + rl) % p.B) * p.ldt) + lane;
    if (STAGE_X || STAGE_T) {
        if (warp == 0) mbar_wait(bar, 0);
        __syncthreads();
    }

    // ---- pass 1: online softmax statistics of the two streams over this warp's periods ------------------------------
    float mL = -INFINITY, mH = -INFINITY;
    f32x2 SE = 0ull, TS = 0ull, TX = 0ull;
    const f32x2 L2E = f2_bcast(kLog2e);
    if (lane_ok) {
        // (a chunk that lies entirely inside the slice -- all but the last -- runs without per-word bounds tests: the
        // body is instantiated twice, `full` is a compile-time constant in each copy)
        auto chunk1 = [&](const int q0, auto full_tag) {
            constexpr bool full = decltype(full_tag)::value;
            uint32_t xw[kPairCh], tw[kPairCh];
            const uint32_t* cx = rx + q0 * Pw;  // one pointer per tensor and chunk; u * Pw is an immediate when DT != 0
            const uint32_t* ct = rt + q0 * Pw;
#pragma unroll
            for (int u = 0; u < kPairCh; ++u)  // (-inf pairs past the slice; un-staged: L1-allocating read-only loads)
                xw[u] = (full || q0 + u < q_hi) ? (STAGE_X ? cx[u * Pw] : __ldg(cx + u * Pw)) : 0xff80ff80u;
#pragma unroll
            for (int u = 0; u < kPairCh; ++u) tw[u] = (full || q0 + u < q_hi) ? (STAGE_T ? ct[u * Pw] : __ldg(ct + u * Pw)) : 0u;
            float cL = -INFINITY, cH = -INFINITY;
#pragma unroll
            for (int u = 0; u < kPairCh; ++u) {
                cL = fmaxf(cL, __uint_as_float(xw[u] << 16));
                cH = fmaxf(cH, __uint_as_float(xw[u] & 0xffff0000u));
            }
            const float nL = fmaxf(mL, cL), nH = fmaxf(mH, cH);
            const float kL = nL == -INFINITY ? 0.f : -nL * kLog2e, kH = nH == -INFINITY ? 0.f : -nH * kLog2e;
            SE = f2_mul(SE, f2_pack(ex2_ftz(fmaf(mL, kLog2e, kL)), ex2_ftz(fmaf(mH, kLog2e, kH))));  // rescale to the new max
            const f32x2 NM = f2_pack(kL, kH);
#pragma unroll
            for (int u = 0; u < kPairCh; ++u) {
                const f32x2 X = bf2_to_f2(xw[u]), T = bf2_to_f2(tw[u]);
                float a0, a1;
                f2_unpack(f2_fma(X, L2E, NM), a0, a1);
                SE = f2_add(SE, f2_pack(ex2_ftz(a0), ex2_ftz(a1)));
                TS = f2_add(TS, T);
                if (full || q0 + u < q_hi) TX = f2_fma(T, X, TX);  // (-inf padding: t = 0 but 0 * -inf is NaN)
            }
            mL = nL;
            mH = nH;
        };
        int q0 = q_lo;
        for (; q0 + kPairCh <= q_hi; q0 += kPairCh) chunk1(q0, std::true_type{});
        if (q0 < q_hi) chunk1(q0, std::false_type{});
        part[((size_t)(rl * W + w) * NPAR + parL) * d + colL] = make_float4(mL, f2_lo(SE), f2_lo(TS), f2_lo(TX));
        if (colH < d) part[((size_t)(rl * W + w) * NPAR + parH) * d + colH] = make_float4(mH, f2_hi(SE), f2_hi(TS), f2_hi(TX));
    }
    __syncthreads();
    // ---- merge: every warp merges the W x NPAR partials of all columns of its row (lane j <-> column j) -------------
    float lse_j = 0.f, ts_j = 0.f, acc = 0.f;
    for (int j = lane; j < d; j += 32) {
        float mm = -INFINITY;
        for (int q = 0; q < W * NPAR; ++q) mm = fmaxf(mm, part[((size_t)rl * W * NPAR + q) * d + j].x);
        const float nm = mm == -INFINITY ? 0.f : -mm * kLog2e;
        float se = 0.f, ts = 0.f, txs = 0.f;
        for (int q = 0; q < W * NPAR; ++q) {
            const float4 v = part[((size_t)rl * W * NPAR + q) * d + j];
            se = fmaf(v.y, ex2_ftz(fmaf(v.x, kLog2e, nm)), se);
            ts += v.z;
            txs += v.w;
        }
        const float lse = mm + logf(se);
        acc += txs - lse * ts;
        if (d <= 32) { lse_j = lse; ts_j = ts; }
        else { s_lse[rl * d + j] = lse; s_ts[rl * d + j] = ts; }
        if (w == 0 && p.stats) {
            p.stats[(row0 + rl) * 2 * d + j] = lse;
            p.stats[(row0 + rl) * 2 * d + d + j] = ts;
        }
    }
    if (w == 0) {
        acc = warp_sum(acc);
        if (lane == 0) p.out_rows[row0 + rl] = p.lam * acc;
    }
    if (MODE == 2) {
        // statistics of this lane's two streams' columns: held by lanes colL / colH (d <= 32) or in shared memory
        float lseL, lseH, tsL, tsH;
        if (d <= 32) {
            lseL = __shfl_sync(0xffffffffu, lse_j, colL & 31); tsL = __shfl_sync(0xffffffffu, ts_j, colL & 31);
            lseH = __shfl_sync(0xffffffffu, lse_j, colH & 31); tsH = __shfl_sync(0xffffffffu, ts_j, colH & 31);
        } else {
            __syncwarp();
            lseL = s_lse[rl * d + colL]; tsL = s_ts[rl * d + colL];
            lseH = s_lse[rl * d + min(colH, d - 1)]; tsH = s_ts[rl * d + min(colH, d - 1)];
        }
        if (lane_ok) {
            const float wl = (p.w_rows ? __ldg(p.w_rows + row0 + rl) : p.w_const) * p.lam;
            const f32x2 NL = f2_pack(-lseL * kLog2e, -lseH * kLog2e), WTS = f2_pack(-wl * tsL, -wl * tsH), WL = f2_bcast(wl);
            // gradient: in place over the staged x (one bulk store per CTA), or straight to global memory
            uint32_t* gx = STAGE_X ? sx + (size_t)rl * (n / 2) + lane
                                   : reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.g) + (row0 + rl) * p.ldg) + lane;
            auto chunk2 = [&](const int q0, auto full_tag) {
                constexpr bool full = decltype(full_tag)::value;
                uint32_t xw[kPairCh], tw[kPairCh];
                const uint32_t* cx = rx + q0 * Pw;
                const uint32_t* ct = rt + q0 * Pw;
                uint32_t* cg = gx + q0 * Pw;
#pragma unroll
                for (int u = 0; u < kPairCh; ++u)
                    xw[u] = (full || q0 + u < q_hi) ? (STAGE_X ? cx[u * Pw] : __ldg(cx + u * Pw)) : 0u;
#pragma unroll
                for (int u = 0; u < kPairCh; ++u) tw[u] = (full || q0 + u < q_hi) ? (STAGE_T ? ct[u * Pw] : __ldg(ct + u * Pw)) : 0u;
#pragma unroll
                for (int u = 0; u < kPairCh; ++u) {
                    float a0, a1;
                    f2_unpack(f2_fma(bf2_to_f2(xw[u]), L2E, NL), a0, a1);
                    const f32x2 G = f2_fma(f2_pack(ex2_ftz(a0), ex2_ftz(a1)), WTS, f2_mul(WL, bf2_to_f2(tw[u])));
                    if (full || q0 + u < q_hi) cg[u * Pw] = f2_to_bf2(G);
                }
            };
            int q0 = q_lo;
            for (; q0 + kPairCh <= q_hi; q0 += kPairCh) chunk2(q0, std::true_type{});
            if (q0 < q_hi) chunk2(q0, std::false_type{});
        }
        if (STAGE_X) {
            fence_async_smem();  // generic-proxy smem writes -> visible to the async (TMA) proxy
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_s2g(reinterpret_cast<__nv_bfloat16*>(p.g) + row0 * p.ldg, sx, (uint32_t)((size_t)R * n * 2));
                bulk_wait_read();  // smem must stay alive until the bulk store has read it
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-resident rows (r2): ONE row per CTA, W = ceil(periods / 16) warps; a lane loads ITS words of the warp's
// <= 16 periods of x and of t up front -- 32 independent 4-byte loads in flight per lane, one global round trip -- and
// keeps them in registers through the statistics pass, the single barrier of the column merge and the gradient pass;
// the gradient goes straight back to global memory.  No staging, no TMA round trip, no second read of anything.
// ncu on the TMA-staged form above at the C5 captions (4096 x 246 x 27 bf16, fused, 49.7 us): issue-active 49 % at 8
// warps per scheduler, stalls split between the barrier (3.2 per issue: 16 warps of a CTA wait for its slowest slice),
// the scoreboard of the chunked target loads (3.3) and fixed-latency waits; DRAM 32 % of peak.  Halving its
// instruction count (compile-time d) moved it by 2 %: latency bound, not issue bound.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kResCh = 16;

template <int MODE, int DT, int TB>  // TB: threads per CTA bound (256: three CTAs per SM, W <= 8; else 512)
__global__ void __launch_bounds__(TB, TB == 256 ? MMVAE_CATCE_RESIDENT_OCC : 1) catce_resident_kernel(const CatceParams p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int d = DT ? DT : p.d;
    const int W = p.W;
    const bool odd = (d & 1) != 0;
    const int Pw = odd ? d : d / 2;      // words per period
    const int NP = odd ? p.C / 2 : p.C;  // periods per row
    const int NPAR = odd ? 2 : 1;
    float4* part = reinterpret_cast<float4*>(smraw);                      // W*NPAR*d partial statistics
    float* s_lse = reinterpret_cast<float*>(part + (size_t)W * NPAR * d);  // d merged logsumexp
    float* s_ts = s_lse + d;                                                // d merged sum t
    float* s_acc = s_ts + d;                                                // W per-warp parts of the row value
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = blockIdx.x;
    const bool lane_ok = lane < Pw;
    int parL = 0, colL = 2 * lane, parH = 0, colH = 2 * lane + 1;
    if (odd) {
        const int h = (d - 1) / 2;
        if (lane == h) { colL = d - 1; parH = 1; colH = 0; }
        else if (lane > h) { parL = parH = 1; colL = 2 * lane - d; colH = colL + 1; }
    }
    const int NPw = p.npw, q_lo = w * NPw;        // (host: ceil(NP / W); no integer division in the kernel)
    const int cnt = min(NP, q_lo + NPw) - q_lo;  // <= kResCh (host); <= 0 for warps past the end
    const int64_t trow = p.same_rows ? row : row % p.B;
    const uint32_t* rx = reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p.x) + row * p.ldx) + lane + q_lo * Pw;
    const uint32_t* rt = reinterpret_cast<const uint32_t*>(reinterpret_cast<const __nv_bfloat16*>(p.t) + trow * p.ldt) + lane + q_lo * Pw;
    uint32_t xw[kResCh], tw[kResCh];
    if (cnt == kResCh && lane_ok) {  // full slice: no per-word bounds tests
#pragma unroll
        for (int u = 0; u < kResCh; ++u) xw[u] = ldg_stream_u32(rx + u * Pw);
#pragma unroll
        for (int u = 0; u < kResCh; ++u) tw[u] = __ldg(rt + u * Pw);
    } else {
#pragma unroll
        for (int u = 0; u < kResCh; ++u) xw[u] = (lane_ok && u < cnt) ? ldg_stream_u32(rx + u * Pw) : 0xff80ff80u;  // -inf pairs
#pragma unroll
        for (int u = 0; u < kResCh; ++u) tw[u] = (lane_ok && u < cnt) ? __ldg(rt + u * Pw) : 0u;
    }
    // ---- statistics of the lane's two streams over the slice ---------------------------------------------------
    float mL = -INFINITY, mH = -INFINITY;
#pragma unroll
    for (int u = 0; u < kResCh; ++u) {
        mL = fmaxf(mL, __uint_as_float(xw[u] << 16));
        mH = fmaxf(mH, __uint_as_float(xw[u] & 0xffff0000u));
    }
    const f32x2 L2E = f2_bcast(kLog2e);
    f32x2 SE = 0ull, TS = 0ull, TX = 0ull;
    {
        const f32x2 NM = f2_pack(mL == -INFINITY ? 0.f : -mL * kLog2e, mH == -INFINITY ? 0.f : -mH * kLog2e);
#pragma unroll
        for (int u = 0; u < kResCh; ++u) {
            const f32x2 X = bf2_to_f2(xw[u]), T = bf2_to_f2(tw[u]);
            float a0, a1;
            f2_unpack(f2_fma(X, L2E, NM), a0, a1);
            SE = f2_add(SE, f2_pack(ex2_ftz(a0), ex2_ftz(a1)));
            TS = f2_add(TS, T);
            if (u < cnt) TX = f2_fma(T, X, TX);  // (-inf padding: t = 0 but 0 * -inf is NaN)
        }
    }
    if (lane_ok) {
        part[((size_t)w * NPAR + parL) * d + colL] = make_float4(mL, f2_lo(SE), f2_lo(TS), f2_lo(TX));
        if (colH < d) part[((size_t)w * NPAR + parH) * d + colH] = make_float4(mH, f2_hi(SE), f2_hi(TS), f2_hi(TX));
    }
    __syncthreads();
    // ---- merge, spread over the CTA: warp w takes the column groups w, w + W, ...; a group is GC = 32 / NS columns, the
    // NS = pow2 >= W*NPAR lanes of a column hold ONE partial each and combine them with log2(NS) shuffle steps.  (The
    // first version had every warp merge all d columns over all W*NPAR partials in two serial loops: ~250 dependent
    // instructions per warp, as many as the two passes over its data.)
    const int np = W * NPAR;
    const int NS = 1 << p.lg_ns;  // (host: pow2 >= np)
    const int GC = 32 >> p.lg_ns, sub = lane & (NS - 1), cg_i = lane >> p.lg_ns;
    float accw = 0.f;
    for (int j0 = w * GC; j0 < d; j0 += W * GC) {
        const int j = j0 + cg_i;
        const bool have = j < d && sub < np;
        const float4 v = have ? part[(size_t)sub * d + j] : make_float4(-INFINITY, 0.f, 0.f, 0.f);
        // segmented butterflies over the NS lanes of a column: the five steps are unrolled, a step that would cross a
        // segment is masked by a uniform predicate (a run-time trip count cost a branch + a register shuffle per step)
        float mm = v.x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_xor_sync(0xffffffffu, mm, o);
            mm = o < NS ? fmaxf(mm, t) : mm;
        }
        const float nm = mm == -INFINITY ? 0.f : -mm * kLog2e;
        float se = have ? v.y * ex2_ftz(fmaf(v.x, kLog2e, nm)) : 0.f, ts = v.z, txs = v.w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float a = __shfl_xor_sync(0xffffffffu, se, o), b = __shfl_xor_sync(0xffffffffu, ts, o),
                        c = __shfl_xor_sync(0xffffffffu, txs, o);
            if (o < NS) {
                se += a;
                ts += b;
                txs += c;
            }
        }
        if (j < d && sub == 0) {
            const float lse = mm + logf(se);
            s_lse[j] = lse;
            s_ts[j] = ts;
            accw += txs - lse * ts;
            if (p.stats) {
                p.stats[row * 2 * d + j] = lse;
                p.stats[row * 2 * d + d + j] = ts;
            }
        }
    }
    accw = warp_sum(accw);
    if (lane == 0) s_acc[w] = accw;
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f;
        for (int q = 0; q < W; ++q) a += s_acc[q];  // fixed order
        p.out_rows[row] = p.lam * a;
    }
    if (MODE == 2) {
        const float lseL = s_lse[min(colL, d - 1)], tsL = s_ts[min(colL, d - 1)];
        const float lseH = s_lse[min(colH, d - 1)], tsH = s_ts[min(colH, d - 1)];
        if (lane_ok && cnt > 0) {
            const float wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;
            const f32x2 NL = f2_pack(-lseL * kLog2e, -lseH * kLog2e), WTS = f2_pack(-wl * tsL, -wl * tsH), WL = f2_bcast(wl);
            uint32_t* cg = reinterpret_cast<uint32_t*>(reinterpret_cast<__nv_bfloat16*>(p.g) + row * p.ldg) + lane + q_lo * Pw;
#pragma unroll
            for (int u = 0; u < kResCh; ++u) {
                float a0, a1;
                f2_unpack(f2_fma(bf2_to_f2(xw[u]), L2E, NL), a0, a1);
                const f32x2 G = f2_fma(f2_pack(ex2_ftz(a0), ex2_ftz(a1)), WTS, f2_mul(WL, bf2_to_f2(tw[u])));
                if (u < cnt) stg_stream_u32(cg + u * Pw, f2_to_bf2(G));
            }
        }
    }
}

// bf16 / bf16 rows whose periods fit 16 per warp with <= 16 warps: register-resident kernel; false otherwise
static bool launch_catce_resident(int mode, CatceParams p, cudaStream_t st, int* rc) {
#if MMVAE_CATCE_RESIDENT
    const int d = p.d;
    if (mode == 1 || p.C < 64) return false;
    if ((d & 1) ? (d > 31 || (p.C & 1)) : d > 64) return false;
    const int NP = (d & 1) ? p.C / 2 : p.C, npar = (d & 1) ? 2 : 1;
    int W = (NP + kResCh - 1) / kResCh;
    if (W > 16) return false;
    const int Pw = (d & 1) ? d : d / 2;
    if (Pw < 24) return false;  // a lane per word of a period: shorter periods leave most of the warp idle
    if (W < MMVAE_CATCE_RESIDENT_MINW) W = MMVAE_CATCE_RESIDENT_MINW;
    // 4-byte words: even row strides and 4-byte aligned bases
    const bool ok = (p.ldx % 2 == 0) && (p.ldt % 2 == 0) && (mode == 0 || p.ldg % 2 == 0) && ((p.C * d) % 2 == 0) &&
                    (reinterpret_cast<uintptr_t>(p.x) % 4 == 0) && (reinterpret_cast<uintptr_t>(p.t) % 4 == 0) &&
                    (mode == 0 || reinterpret_cast<uintptr_t>(p.g) % 4 == 0);
    if (!ok) return false;
    p.R = 1;
    p.W = W;
    p.npw = (NP + W - 1) / W;
    p.lg_ns = 0;
    while ((1 << p.lg_ns) < W * npar) ++p.lg_ns;
    p.same_rows = p.rows == p.B;
    const size_t smem = (size_t)W * npar * d * 16 + 2 * d * 4 + 16 * 4;
    void (*k)(const CatceParams);
    if (W <= 8)
        k = d == 27 ? (mode == 0 ? catce_resident_kernel<0, 27, 256> : catce_resident_kernel<2, 27, 256>)
                    : (mode == 0 ? catce_resident_kernel<0, 0, 256> : catce_resident_kernel<2, 0, 256>);
    else
        k = d == 27 ? (mode == 0 ? catce_resident_kernel<0, 27, 512> : catce_resident_kernel<2, 27, 512>)
                    : (mode == 0 ? catce_resident_kernel<0, 0, 512> : catce_resident_kernel<2, 0, 512>);
    k<<<(unsigned)p.rows, W * 32, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? 0 : (int)e;
    return true;
#else
    (void)mode; (void)p; (void)st; (void)rc;
    return false;
#endif
}

static size_t pairs_smem(int R, int W, int n, int d, bool stage_t) {
    const int npar = (d & 1) ? 2 : 1;
    size_t off = up16((size_t)R * n * 2) * ((stage_t ? 1 : 0) + (MMVAE_CATCE_PAIRS_STAGE_X ? 1 : 0));
    off = up16(off + (size_t)2 * R * d * 4);
    return off + (size_t)R * W * npar * d * 16 + 16;
}

// bf16 / bf16, forward or fused, dense TMA-able slabs, d odd (<= 31, C even) or d even (<= 64); false otherwise
static bool launch_catce_pairs(int mode, CatceParams p, cudaStream_t st, int* rc) {
#if MMVAE_CATCE_PAIRS
    const int n = p.C * p.d, d = p.d;
    if (mode == 1 || p.C < 64) return false;  // long class axes only: short rows stay on the tuned r1 kernels
    if ((d & 1) ? (d > 31 || (p.C & 1)) : d > 64) return false;
    int R = MMVAE_CATCE_PAIRS_R, W = MMVAE_CATCE_PAIRS_W;
    while (R > 1 && (pairs_smem(R, W, n, d, MMVAE_CATCE_PAIRS_STAGE_T) > 100 * 1024 || R * W > 16)) R >>= 1;
    const bool dense = p.ldx == n && p.ldt == n && (mode == 0 || p.ldg == n);
    const bool ok = dense && aligned16(p.x) && aligned16(p.t) && (mode == 0 || aligned16(p.g)) && p.rows % R == 0 &&
                    p.B % R == 0 && ((size_t)R * n * 2) % 16 == 0 && (n % 2) == 0;
    if (!ok) return false;
    p.R = R;
    p.W = W;
    const size_t smem = pairs_smem(R, W, n, d, MMVAE_CATCE_PAIRS_STAGE_T);
    constexpr bool kST = MMVAE_CATCE_PAIRS_STAGE_T != 0, kSX = MMVAE_CATCE_PAIRS_STAGE_X != 0;
    auto k = d == 27 ? (mode == 0 ? catce_pairs_kernel<0, kST, kSX, 27> : catce_pairs_kernel<2, kST, kSX, 27>)
                     : (mode == 0 ? catce_pairs_kernel<0, kST, kSX, 0> : catce_pairs_kernel<2, kST, kSX, 0>);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { *rc = (int)e; return true; }
    }
    k<<<(unsigned)(p.rows / R), R * W * 32, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    *rc = e == cudaSuccess ? 0 : (int)e;
    return true;
#else
    (void)mode; (void)p; (void)st; (void)rc;
    return false;
#endif
}

// ---------------------------------------------------------------------------------------------------------------
// v2 (default): no shared-memory staging.
//
// ncu r1 on the TMA-staged kernel above (C2 text term, 7680 x 45 x 27 fp32): 18.7 us forward / 21.7 us backward for
// 37 / 76 MB -- one-shot CTAs alternate between a TMA round trip and a short compute phase, and the L2-resident target
// slab goes through the same smem pipe as the streamed reconstruction.  v2 splits the work by access pattern:
//
//   catce_cols_kernel      a warp owns 32/d rows (d < 32) or one row; a lane owns one column j and walks the class
//                          axis in register chunks of CH values with an online (running max / rescaled sum)
//                          softmax: CH independent loads per tensor in flight per lane, one pass over x and t, no
//                          barrier, no smem, any alignment or row stride.  Writes the row value and the per-column
//                          (logsumexp, sum t) statistics; MODE 1/2 add a second chunked pass that re-reads the row
//                          (L1/L2 hits) and writes the gradient.
//   catce_flat_bwd_kernel  with cached statistics the gradient is ELEMENT-WISE: g[e] = wl*t[e] - wl*ts[j]*exp(x[e] -
//                          lse[j]); a CTA takes a dense slab of R rows as a flat run of 128-bit vectors (streaming
//                          loads / stores as in loglik.cu) and looks its column statistics up in a few hundred bytes
//                          of smem.  2R+T bytes, ~12 instructions per element.
// ---------------------------------------------------------------------------------------------------------------
#ifndef MMVAE_CATCE_IMPL
#define MMVAE_CATCE_IMPL 1  // 0: TMA-staged kernel for everything, 1: + flat streaming backward, 2: v2 kernels only
#endif
#ifndef MMVAE_CATCE_COLS_MINBLOCKS
#define MMVAE_CATCE_COLS_MINBLOCKS 3
#endif

template <typename TX, typename TT, int MODE>  // 0 fwd (+ stats), 1 bwd without cached stats, 2 fused
__global__ void __launch_bounds__(256, MMVAE_CATCE_COLS_MINBLOCKS) catce_cols_kernel(const CatceParams p) {
    constexpr int CH = MMVAE_CATCE_CH;
    const int lane = threadIdx.x & 31, d = p.d, C = p.C, G = p.G;
    const int rsub = d < 32 ? lane / d : 0;
    const int jl = lane - rsub * d;
    const bool lane_ok = d >= 32 || rsub < G;
    const int wpc = blockDim.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * wpc;
    const int64_t groups = (p.rows + G - 1) / G;
    const TX* __restrict__ xg = reinterpret_cast<const TX*>(p.x);
    const TT* __restrict__ tg = reinterpret_cast<const TT*>(p.t);
    for (int64_t grp = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5); grp < groups; grp += nwarps) {
        const int64_t row = grp * G + rsub;
        const bool ok = lane_ok && row < p.rows;
        float acc = 0.f;
        if (ok) {
            const TX* xr = xg + row * p.ldx;
            const TT* tr = tg + (row % p.B) * p.ldt;
            float wl = 0.f;
            if (MODE != 0) wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;
            for (int j = jl; j < d; j += 32) {
                const TX* px = xr + j;
                const TT* pt = tr + j;
                float lse, ts, txs;
                col_stats<TX, TT, CH, true>(px, pt, C, d, lse, ts, txs);
                acc += txs - lse * ts;
                if (MODE == 0 && p.stats) {
                    p.stats[row * 2 * d + j] = lse;
                    p.stats[row * 2 * d + d + j] = ts;
                }
                if (MODE != 0)  // second pass: the row was just read, so these loads hit L1 / L2
                    col_grad<TX, TT, CH, true>(px, pt, reinterpret_cast<TX*>(p.g) + row * p.ldg + j, C, d, lse, ts, wl);
            }
        }
        if (MODE != 1) {
            float v = acc;
            if (d >= 32) {
                v = warp_sum(v);
            } else {  // segmented sum over the d lanes of a row
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float nb = __shfl_down_sync(0xffffffffu, v, o);
                    if (jl + o < d) v += nb;
                }
            }
            if (ok && jl == 0) p.out_rows[row] = p.lam * v;
        }
    }
}

template <typename TX, typename TT, int V>  // V = 16 / sizeof(TX): dense aligned slabs; V = 1: any stride / alignment
__global__ void __launch_bounds__(256, 3) catce_flat_bwd_kernel(const CatceParams p) {
    extern __shared__ __align__(16) float sm2[];
    constexpr int U = V >= 8 ? 3 : 6, T = 256;  // 24 data registers per tensor and batch
    const int d = p.d, n = p.C * p.d, R = p.R;
    float* s_nl = sm2;           // R*d: -logsumexp * log2(e)
    float* s_wts = sm2 + R * d;  // R*d: -w*lam * sum_c t
    float* s_wl = s_wts + R * d; // R:   w*lam
    const int64_t row0 = (int64_t)blockIdx.x * R;
    const int nrows = (int)min((int64_t)R, p.rows - row0);
    const int total = nrows * n;
    const int rb0 = (int)(row0 % p.B);
    const TX* __restrict__ xg = reinterpret_cast<const TX*>(p.x);
    const TT* __restrict__ tg = reinterpret_cast<const TT*>(p.t);
    TX* __restrict__ gg = reinterpret_cast<TX*>(p.g);

    auto split = [&](int off, int& r, int& ee) {  // off = r*n + ee, exact for R*n < 2^21 (host checks)
        r = (int)(((float)off + 0.5f) * p.inv_n);
        ee = off - r * n;
    };
    float xv[U][V], tv[U][V];
    auto load = [&](int e0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int off = e0 + u * T * V;
            if (off < total) {
                if (V > 1) {
                    load_vec<TX, V>(xg + row0 * n + off, xv[u], true);
                    load_vec<TT, V>(tg + (int64_t)rb0 * n + off, tv[u], false);
                } else {
                    int r, ee;
                    split(off, r, ee);
                    xv[u][0] = Elem<TX>::load1(xg + (row0 + r) * p.ldx + ee);
                    tv[u][0] = Elem<TT>::load1(tg + (int64_t)((rb0 + r) % (int)p.B) * p.ldt + ee);
                }
            }
        }
    };
    int e0 = threadIdx.x * V;
    load(e0);  // in flight while the column statistics are staged
    for (int i = threadIdx.x; i < nrows * d; i += T) {
        const int r = (int)(((float)i + 0.5f) * p.inv_d), j = i - r * d;
        const float* st = p.stats + (row0 + r) * 2 * d;
        const float wl = __ldg(p.w_rows + row0 + r) * p.lam;
        s_nl[i] = -__ldg(st + j) * kLog2e;
        s_wts[i] = -wl * __ldg(st + d + j);
        if (j == 0) s_wl[r] = wl;
    }
    __syncthreads();
    while (true) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int off = e0 + u * T * V;
            if (off < total) {
                int r, ee;
                split(off, r, ee);
                int j = ee - (int)(((float)ee + 0.5f) * p.inv_d) * d;
                int rd = r * d;
                float wl = s_wl[r];
                float gv[V];
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    gv[i] = fmaf(ex2_ftz(fmaf(xv[u][i], kLog2e, s_nl[rd + j])), s_wts[rd + j], wl * tv[u][i]);
                    if (V > 1) {
                        ++j; ++ee;
                        if (j == d) j = 0;
                        if (ee == n && i + 1 < V) { ee = 0; j = 0; ++r; rd += d; wl = s_wl[r]; }
                    }
                }
                if constexpr (V > 1) stg_stream(gg + row0 * n + off, Elem<TX>::pack(gv));
                else Elem<TX>::store1(gg + (row0 + r) * p.ldg + ee, gv[0]);
            }
        }
        e0 += U * T * V;
        if (e0 >= total) break;
        load(e0);
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Ring kernel (forward / fused, default when the slabs are TMA-able): persistent CTAs, producer / consumer.
//
// The one-shot staged kernel above spends its life waiting: every CTA does TMA round trip -> short compute -> exit, and
// 2.6 waves of that cost 18.7 us for 37 MB (C2 text).  Here a CTA owns a contiguous range of ITEMS (R consecutive
// rows each) and keeps S slabs in flight:
//   * warp NW (one elected lane) is the producer: waits for a stage to drain (`empty` mbarrier, one arrival per compute
//     warp), [fused mode: sends the gradient written in place over that stage out with one bulk store], then issues
//     the TMA bulk load of the next item into it (`full` mbarrier, complete_tx);
//   * warps 0..NW-1 compute: wait `full`, walk their rows (lanes over columns, two smem passes over the class axis:
//     max, then exp-sum / target sums), write row value + column statistics [+ gradient in place], arrive on `empty`.
//   * items are enumerated TARGET-major (item = bgroup*K + k, rows k*B + bgroup*R ...): consecutive items of a CTA
//     share their R target rows, so the target slab is loaded once per bgroup instead of once per item (K = 30 at
//     C2: the L2 -> SM traffic of the targets drops from 1x the reconstruction to ~1/6 of it).  Target slabs live in
//     their own TS slots (2 when K >= S, else S).
// ---------------------------------------------------------------------------------------------------------------
#ifndef MMVAE_CATCE_RING
// Measured r1 (profiles/r1_tune_catce.txt, profiles/r1_ncu_catce_notes.txt): correct, but SLOWER than the one-shot
// staged kernel on every benchmark shape (C2 text forward 22.5 vs 18.7 us, fused 34 vs 31 us): the kernel is not
// memory bound -- a row costs ~900 warp instructions at 2 consumer warps per scheduler (97 KB of smem per CTA caps
// the SM at 8 consumer warps), while the one-shot kernel keeps 20 warps per SM resident.  Kept as an opt-in
// experiment (-DMMVAE_CATCE_RING=1); off by default.
#define MMVAE_CATCE_RING 0
#endif
#ifndef MMVAE_CATCE_RING_XB
#define MMVAE_CATCE_RING_XB (24 * 1024)  // preferred bytes of one x stage
#endif
#ifndef MMVAE_CATCE_RING_SMEM
#define MMVAE_CATCE_RING_SMEM (110 * 1024)  // preferred smem per CTA (2 CTAs per SM)
#endif

template <typename TX, typename TT, int MODE>  // 0 fwd (+ stats), 2 fused
__global__ void __launch_bounds__(288) catce_ring_kernel(const CatceParams p) {
    extern __shared__ __align__(128) unsigned char smraw[];
    const int n = p.C * p.d, d = p.d, C = p.C, R = p.R, S = p.S, TS = p.TS, G = p.G;
    const uint32_t xb = (uint32_t)((size_t)R * n * sizeof(TX)), tb = (uint32_t)((size_t)R * n * sizeof(TT));
    unsigned char* sx = smraw;
    unsigned char* stg = smraw + (size_t)S * xb;
    uint64_t* full = reinterpret_cast<uint64_t*>(smraw + (size_t)S * xb + (size_t)TS * tb);
    uint64_t* empty = full + S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NW = (blockDim.x >> 5) - 1;
    const int64_t K = p.rows / p.B;
    const int64_t items = p.rows / R;
    const int64_t q = items / gridDim.x, rem = items % gridDim.x;
    const int64_t i0 = (int64_t)blockIdx.x * q + min((int64_t)blockIdx.x, rem);
    const int cnt = (int)(q + ((int64_t)blockIdx.x < rem ? 1 : 0));
    const TX* __restrict__ xg = reinterpret_cast<const TX*>(p.x);
    const TT* __restrict__ tg = reinterpret_cast<const TT*>(p.t);
    TX* __restrict__ gg = reinterpret_cast<TX*>(p.g);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, NW);
        }
    }
    __syncthreads();

    // item -> (target group, sample k) -> first row; one 64-bit division per CTA, then incremental
    struct Cursor {
        int64_t bg, k;
    };
    const Cursor c0{i0 / K, i0 % K};
    auto advance = [&](Cursor& c) {
        if (++c.k == K) {
            c.k = 0;
            ++c.bg;
        }
    };
    auto first_row = [&](const Cursor& c) { return c.k * p.B + c.bg * R; };

    if (warp == NW) {  // ---- producer ----
        if (lane == 0) {
            int64_t prev_bg = -1;
            Cursor cl = c0, cs = c0;  // load cursor (item li), store cursor (item li - S)
            int s = 0, ph = 0;        // stage and use count parity of item li
            for (int li = 0; li < cnt; ++li) {
                if (li >= S) {
                    mbar_wait(empty + s, ph ^ 1);
                    if (MODE == 2) {
                        bulk_s2g(gg + first_row(cs) * n, sx + (size_t)s * xb, xb);
                        bulk_wait_read();  // the stage is overwritten next
                    }
                    advance(cs);
                }
                const bool newt = cl.bg != prev_bg;
                prev_bg = cl.bg;
                mbar_expect_tx(full + s, xb + (newt ? tb : 0u));
                bulk_g2s(sx + (size_t)s * xb, xg + first_row(cl) * n, xb, full + s);
                if (newt) bulk_g2s(stg + (size_t)(cl.bg % TS) * tb, tg + cl.bg * R * n, tb, full + s);
                advance(cl);
                if (++s == S) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (MODE == 2) {  // drain: gradients of the last min(cnt, S) items (cs points at item max(0, cnt - S))
                for (int li = max(0, cnt - S); li < cnt; ++li) {
                    const int s2 = li % S;
                    mbar_wait(empty + s2, (li / S) & 1);
                    bulk_s2g(gg + first_row(cs) * n, sx + (size_t)s2 * xb, xb);
                    advance(cs);
                }
                bulk_wait_read();  // smem must outlive the reads of the bulk stores
            }
        }
        return;
    }

    // ---- consumers ----
    const int rsub = d < 32 ? lane / d : 0;
    const int jl = lane - rsub * d;
    const bool lane_ok = d >= 32 || rsub < G;
    Cursor cc = c0;
    int s = 0, ph = 0;
    for (int li = 0; li < cnt; ++li) {
        if (lane == 0) mbar_wait(full + s, ph);
        __syncwarp();
        const int64_t xrow0 = first_row(cc);
        TX* bx = reinterpret_cast<TX*>(sx + (size_t)s * xb);
        const TT* bt = reinterpret_cast<const TT*>(stg + (size_t)(cc.bg % TS) * tb);
        for (int rg = warp; rg * G < R; rg += NW) {
            const int rl = rg * G + rsub;
            const bool ok = lane_ok && rl < R;
            float acc = 0.f;
            if (ok) {
                const int64_t row = xrow0 + rl;
                float wl = 0.f;
                if (MODE != 0) wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;
                for (int j = jl; j < d; j += 32) {
                    TX* px = bx + (size_t)rl * n + j;
                    const TT* pt = bt + (size_t)rl * n + j;
                    float lse, ts, txs;
                    col_stats<TX, TT, MMVAE_CATCE_SCH, false>(px, pt, C, d, lse, ts, txs);
                    acc += txs - lse * ts;
                    if (MODE == 0 && p.stats) {
                        p.stats[row * 2 * d + j] = lse;
                        p.stats[row * 2 * d + d + j] = ts;
                    }
                    // gradient in place over the staged reconstruction
                    if (MODE != 0) col_grad<TX, TT, MMVAE_CATCE_SCH, false>(px, pt, px, C, d, lse, ts, wl);
                }
            }
            float v = acc;
            if (d >= 32) {
                v = warp_sum(v);
            } else {  // segmented sum over the d lanes of a row
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float nb = __shfl_down_sync(0xffffffffu, v, o);
                    if (jl + o < d) v += nb;
                }
            }
            if (ok && jl == 0) p.out_rows[xrow0 + rl] = p.lam * v;
        }
        if (MODE == 2) fence_async_smem();  // generic-proxy gradient writes -> visible to the bulk store
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
        advance(cc);
        if (++s == S) {
            s = 0;
            ph ^= 1;
        }
    }
}

// R, S, TS, NW for the ring kernel; false when the shape is not ring-able (unaligned / strided / slabs too large)
template <typename TX, typename TT>
static bool ring_plan(int mode, CatceParams& p, int* nw, size_t* smem, int* grid) {
    const int n = p.C * p.d, sx = (int)sizeof(TX), stt = (int)sizeof(TT);
    if (mode == 1) return false;
    const bool dense = p.ldx == n && p.ldt == n && (mode == 0 || p.ldg == n) && aligned16(p.x) && aligned16(p.t) &&
                       (mode == 0 || aligned16(p.g));
    if (!dense || p.rows % p.B != 0) return false;
    int u = 1;  // alignment unit: rows per 16-byte aligned slab
    while (u <= 16 && (((size_t)u * n * sx) % 16 != 0 || ((size_t)u * n * stt) % 16 != 0)) u <<= 1;
    if (u > 16 || p.B % u != 0) return false;
    const int G = p.d < 32 ? 32 / p.d : 1;
    const int64_t K = p.rows / p.B;
    // R: the largest u*2^k dividing B with an x stage <= the preferred size and <= 4 row groups per compute warp,
    // reduced again while the launch would have fewer than 2 items per SM
    int64_t R = u;
    while (R * 2 <= p.B && p.B % (R * 2) == 0 && (size_t)(R * 2) * n * sx <= MMVAE_CATCE_RING_XB && R * 2 <= 8 * G * 4) R *= 2;
    while (R > u && p.rows / R < 2 * kNumSMs) R /= 2;
    const size_t xb = (size_t)R * n * sx, tb = (size_t)R * n * stt;
    int S = 0, TS = 0;
    size_t need = 0;
    for (size_t budget : {(size_t)MMVAE_CATCE_RING_SMEM, (size_t)216 * 1024}) {
        for (int s = 4; s >= 2 && !S; --s) {
            const int ts = K >= s ? 2 : s;
            const size_t b = s * xb + ts * tb + 2 * s * sizeof(uint64_t) + 16;
            if (b <= budget) { S = s; TS = ts; need = b; }
        }
        if (S) break;
    }
    if (!S) return false;
    p.R = (int)R; p.S = S; p.TS = TS; p.G = G;
    const int groups = (int)((R + G - 1) / G);
    *nw = groups < 8 ? groups : 8;
    *smem = need;
    int cps = (int)((size_t)227 * 1024 / (need + 1024));
    if (cps > 4) cps = 4;
    if (cps < 1) cps = 1;
    const int64_t items = p.rows / R;
    *grid = (int)(items < (int64_t)kNumSMs * cps ? items : (int64_t)kNumSMs * cps);
    return true;
}

static size_t catce_smem(int R, int W, int n, int d, int sx, int st) {
    size_t off = up16((size_t)R * n * sx) + up16((size_t)R * n * st);
    off = up16(off + (size_t)(2 * R * d + 4 * R * W * d) * 4);
    return off + 16;
}


static int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

template <typename TX, typename TT>
static int launch_catce_v2(int mode, CatceParams p, cudaStream_t st) {
    const int n = p.C * p.d;
    if (mode == 1 && p.stats) {
        constexpr int V = 16 / (int)sizeof(TX);
        constexpr int kSlab = 256 * 24;  // elements a CTA keeps in flight per tensor (U*V = 24 per thread)
        p.inv_n = 1.0f / (float)n;
        p.inv_d = 1.0f / (float)p.d;
        const bool dense = p.ldx == n && p.ldt == n && p.ldg == n && aligned16(p.x) && aligned16(p.t) && aligned16(p.g);
        int R = 0;
        bool vec = false;
        if (dense && p.B < (1LL << 30)) {  // R rows: R*n a multiple of V, R | B (a slab never straddles the target wrap)
            const int u = V / gcd_i(n, V);
            for (int64_t r = u; r <= p.B && r * p.d <= 4096 && (r * n <= kSlab || r == u); r *= 2)
                if (p.B % r == 0) R = (int)r;
            vec = R > 0;
        }
        if (!vec && p.d <= 4096 && p.B < (1LL << 30)) {
            int64_t r = kSlab / n;
            if (r > 4096 / p.d) r = 4096 / p.d;
            R = (int)(r < 1 ? 1 : r);
        }
        if (R > 0) {
            p.R = R;
            const size_t smem = (size_t)(2 * R * p.d + R) * sizeof(float);
            const int64_t grid = (p.rows + R - 1) / R;
            if (grid > 0x7fffffffLL) return MMVAE_E_LIMIT;
            if (vec) catce_flat_bwd_kernel<TX, TT, V><<<(unsigned)grid, 256, smem, st>>>(p);
            else catce_flat_bwd_kernel<TX, TT, 1><<<(unsigned)grid, 256, smem, st>>>(p);
            MMVAE_LAUNCH_CHECK();
            return 0;
        }
        // column statistics wider than the smem table: recompute them in the column kernel
    }
    p.G = p.d < 32 ? 32 / p.d : 1;
    const int64_t groups = (p.rows + p.G - 1) / p.G;
    int64_t grid = (groups + 7) / 8;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;  // grid-stride over row groups beyond that
    if (mode == 0) catce_cols_kernel<TX, TT, 0><<<(unsigned)grid, 256, 0, st>>>(p);
    else if (mode == 1) catce_cols_kernel<TX, TT, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    else catce_cols_kernel<TX, TT, 2><<<(unsigned)grid, 256, 0, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Text-decoder tail fused in (reference decoders.py:722, "zero for padded area": output * mask[:, :, None]): the
// reconstruction arrives UNMASKED together with the (B, C) padding mask, the kernel evaluates the loss on
// x_eff[r, c, j] = mask[r % B, c] ? x[r, c, j] : 0 and returns the gradient with respect to the unmasked tensor
// (zero at masked positions).  A warp per row, lanes over the columns j (coalesced d-element runs), the class axis is
// walked with an online softmax; the gradient pass re-reads the row (L1 / L2).  Any row stride / alignment / dtype.
// (The transformer text decoder drops the K axis -- SURVEY N3 -- so these tensors are B x T x 27: latency, not bandwidth.)
// ---------------------------------------------------------------------------------------------------------------
template <typename TX, typename TT, int MODE>  // 0 forward (+ statistics), 1 backward, 2 fused
__global__ void __launch_bounds__(256) catce_masked_kernel(const CatceParams p) {
    const int lane = threadIdx.x & 31, d = p.d, C = p.C;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const TX* __restrict__ xg = reinterpret_cast<const TX*>(p.x);
    const TT* __restrict__ tg = reinterpret_cast<const TT*>(p.t);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < p.rows; row += nwarps) {
        const int64_t b = row % p.B;
        const TX* xr = xg + row * p.ldx;
        const TT* tr = tg + b * p.ldt;
        const unsigned char* mk = p.mask + b * p.ldm;
        float wl = 0.f;
        if (MODE != 0) wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;
        float acc = 0.f;
        for (int j = lane; j < d; j += 32) {
            float m = -INFINITY, se = 0.f, ts = 0.f, txs = 0.f;
            for (int c = 0; c < C; ++c) {
                const float xv = __ldg(mk + c) ? Elem<TX>::load1(xr + (int64_t)c * d + j) : 0.f;
                const float tv = Elem<TT>::load1(tr + (int64_t)c * d + j);
                const float nm = fmaxf(m, xv);
                se = se * expf(m - nm) + expf(xv - nm);
                m = nm;
                ts += tv;
                txs = fmaf(tv, xv, txs);
            }
            const float lse = m + logf(se);
            acc += txs - lse * ts;
            if (MODE == 0 && p.stats) {
                p.stats[row * 2 * d + j] = lse;
                p.stats[row * 2 * d + d + j] = ts;
            }
            if (MODE != 0) {
                TX* gr = reinterpret_cast<TX*>(p.g) + row * p.ldg;
                for (int c = 0; c < C; ++c) {
                    const bool keep = __ldg(mk + c) != 0;
                    const float xv = keep ? Elem<TX>::load1(xr + (int64_t)c * d + j) : 0.f;
                    const float tv = Elem<TT>::load1(tr + (int64_t)c * d + j);
                    Elem<TX>::store1(gr + (int64_t)c * d + j, keep ? wl * (tv - expf(xv - lse) * ts) : 0.f);
                }
            }
        }
        if (MODE != 1) {
            acc = warp_sum(acc);
            if (lane == 0) p.out_rows[row] = p.lam * acc;
        }
    }
}

template <typename TX, typename TT>
static int launch_catce_masked(int mode, const CatceParams& p, cudaStream_t st) {
    int64_t grid = (p.rows + 7) / 8;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    if (mode == 0) catce_masked_kernel<TX, TT, 0><<<(unsigned)grid, 256, 0, st>>>(p);
    else if (mode == 1) catce_masked_kernel<TX, TT, 1><<<(unsigned)grid, 256, 0, st>>>(p);
    else catce_masked_kernel<TX, TT, 2><<<(unsigned)grid, 256, 0, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    return 0;
}

template <typename TX, typename TT>
static int launch_catce(int mode, CatceParams p, cudaStream_t st) {
#if MMVAE_CATCE_RING
    {
        int nw, grid;
        size_t smem;
        CatceParams q = p;
        if (ring_plan<TX, TT>(mode, q, &nw, &smem, &grid)) {
            auto k = mode == 0 ? catce_ring_kernel<TX, TT, 0> : catce_ring_kernel<TX, TT, 2>;
            if (smem > 48 * 1024) {
                cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return (int)e;
            }
            k<<<(unsigned)grid, (nw + 1) * 32, smem, st>>>(q);
            MMVAE_LAUNCH_CHECK();
            return 0;
        }
    }
#endif
#if MMVAE_CATCE_IMPL == 2
    return launch_catce_v2<TX, TT>(mode, p, st);
#elif MMVAE_CATCE_IMPL == 1
    // measured r1 (tools/tune_catce.sh, profiles/r1_tune_catce.txt): the flat streaming backward beats the staged one
    // everywhere (C2 text 27.9 -> 22.9 us, C5 bf16 captions 65.7 -> 55.2 us); the column kernel only ties the staged
    // forward on long rows and loses the fused pass, so those stay on the TMA-staged kernel
    if (mode == 1 && p.stats) return launch_catce_v2<TX, TT>(mode, p, st);
    // rows of a few hundred bytes (actions, attributes, MNIST-sized label maps): the column kernel needs no staging
    // round trip (K = 50, 10 x 12: forward 37 -> 26 us, fused 57 -> 41 us)
    if ((size_t)p.C * p.d * sizeof(TX) <= 512) return launch_catce_v2<TX, TT>(mode, p, st);
#endif
    const int n = p.C * p.d, sx = (int)sizeof(TX), stt = (int)sizeof(TT);
    const bool fast_bwd = (mode == 1 && p.stats != nullptr);
    // warps per row split the class axis (short dependency chains); the cached-statistics backward has no per-row
    // reduction and simply strides all warps over the staged class rows
    // (measured r1, C = 45: W = 1 -> 17 us, W = 4 -> 27 us: the merge costs more than the shorter chains save)
#ifndef MMVAE_CATCE_W_LONG
#define MMVAE_CATCE_W_LONG 2
#endif
    int W = fast_bwd ? 1 : (p.C >= 96 ? MMVAE_CATCE_W_LONG : 1);
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
    return mmvae_catce_rows_masked(mode, recon, ld_recon, dtype_recon, target, ld_target, dtype_target, rows, B, C, d, lam,
                                   w_rows, w_const, out_rows, grad_recon, ld_grad, stats, nullptr, 0, stream);
}

extern "C" int mmvae_catce_rows_masked(int mode, const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                       int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t C, int64_t d,
                                       float lam, const float* w_rows, float w_const, float* out_rows, void* grad_recon,
                                       int64_t ld_grad, float* stats, const unsigned char* mask, int64_t ld_mask,
                                       void* stream) {
    if (!recon || !target || rows <= 0 || B <= 0 || C <= 0 || d <= 0) return MMVAE_E_ARG;
    if (mask && ld_mask < C) return MMVAE_E_ARG;
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
    p.mask = mask; p.ldm = ld_mask;
    cudaStream_t st = (cudaStream_t)stream;
    if (mask) {  // text-decoder tail fused in: one kernel for every shape / stride / dtype pair
        if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_F32) return launch_catce_masked<float, float>(mode, p, st);
        if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_F32) return launch_catce_masked<__nv_bfloat16, float>(mode, p, st);
        if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_BF16) return launch_catce_masked<__nv_bfloat16, __nv_bfloat16>(mode, p, st);
        if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_BF16) return launch_catce_masked<float, __nv_bfloat16>(mode, p, st);
        return MMVAE_E_ENUM;
    }
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_F32) return launch_catce<float, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_F32) return launch_catce<__nv_bfloat16, float>(mode, p, st);
    if (dtype_recon == MMVAE_BF16 && dtype_target == MMVAE_BF16) {
        int rc = 0;
        if (launch_catce_resident(mode, p, st, &rc)) return rc;
        if (launch_catce_pairs(mode, p, st, &rc)) return rc;
        return launch_catce<__nv_bfloat16, __nv_bfloat16>(mode, p, st);
    }
    if (dtype_recon == MMVAE_F32 && dtype_target == MMVAE_BF16) return launch_catce<float, __nv_bfloat16>(mode, p, st);
    return MMVAE_E_ENUM;
}
