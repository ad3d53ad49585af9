// Likelihood row reductions (element-wise families) -- the kernel that carries >95% of the bytes of every image
// configuration.  HBM-bound streaming reduction: 128-bit no-allocate loads of the reconstruction,
// cached loads of the K-times reused target, per-thread fp32 partials, warp-shuffle + smem block reduction, and
// (bwd / fused modes) a 128-bit streaming store of the gradient.  No tensor cores: there is no contraction here.
//
// Replaces reference objectives.py:30-52 / :103-125 / :389-458 (recon_loss_fn, reshape_for_loss, ReconLoss.bce,
// lprob, mse, l1) together with the "* llik_scaling).sum(-1)" of every call site in mmvae_models.py.
#include <cstdlib>

#include "common.cuh"

namespace mmvae {

enum { MODE_FWD = 0, MODE_BWD = 1, MODE_FUSED = 2 };

constexpr int kThreads = 256;
// 128-bit loads in flight per thread and tensor.
// tuning knobs (tools/tune_loglik.sh compiles variants with -D and times them on the box)
#ifndef MMVAE_FWD_UNROLL
#define MMVAE_FWD_UNROLL 4
#endif
#ifndef MMVAE_FWD_MINBLOCKS
#define MMVAE_FWD_MINBLOCKS 5  // r1: 77.6 -> 71.8 us for the C2 forward (5 CTAs of 256 threads per SM, <= 51 registers)
#endif
#ifndef MMVAE_BCE_FAST
#define MMVAE_BCE_FAST 1  // unclamped BCE forward with one finiteness test per vector (exact slow path behind it)
#endif
template <int MODE>
struct Tune {
    // measured r1 (see profiles/r1_tune_loglik_fwd.txt)
    static constexpr int kUnroll = (MODE == 0) ? MMVAE_FWD_UNROLL : 4;
    static constexpr int kMinBlocks = (MODE == 0) ? MMVAE_FWD_MINBLOCKS : 4;
};
constexpr int kUnrollMax = 4;

struct LoglikParams {
    const void* x;
    const void* t;
    void* g;
    const float* w_rows;
    float* out_rows;
    float* ws;
    int64_t ldx, ldt, ldg;
    int64_t rows, B, P;
    int64_t chunk;  // elements of a row handled by one CTA (multiple of the vector width)
    int cpr;        // CTAs per row (long, few rows) ...
    int tpr;        // ... or threads per row (32..256, power of two): short rows share a CTA, 256/tpr rows each
    int tile;       // batch rows per L2 tile: CTAs sweep all K samples of `tile` targets before moving on
    float scale, lam, w_const;
};

// MUFU.LG2 without the compiler's per-call denormal rescue (3 extra instructions per log).  Sub-normal inputs are
// handled exactly by the caller's vector-level guard (slow path below), everything else is identical.
__device__ __forceinline__ float lg2_ftz(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// MUFU.RCP without __fdividef's range rescue (2 FSETP + 2 predicated FMUL per call: ncu/SASS r1, bf16 fused pass at 32
// instructions per element).  Callers guarantee a normal, positive argument.
__device__ __forceinline__ float rcp_ftz(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

#ifndef MMVAE_BCE_BF16_NEWTON
#define MMVAE_BCE_BF16_NEWTON 0
#endif
// r2, measured and rejected (build option): 1/v for v in [1e-12, 0.25] WITHOUT the MUFU pipe, for gradients that are
// rounded to bf16 anyway -- magic-constant initial guess + two Newton steps on the FMA pipe, relative error 7e-6.
// Hypothesis: the bf16 fused BCE pass moves 6 bytes per element and spends 3 MUFU operations on it (two LG2, one RCP);
// at the HBM rate that is 3e12 MUFU/s of the 4.5e12 the 148 SMs have, so the math pipe, not memory, would be the
// ceiling of its 5.0 TB/s (77 %).  Measured on C5 (B = 4096): 5.03 TB/s with the Newton form, 5.09 TB/s with MUFU.RCP:
// the MUFU pipe is not the limit.
__device__ __forceinline__ float rcp_newton2(float v) {
    float r = __uint_as_float(0x7EF311C7u - __float_as_uint(v));
    r = fmaf(r, fmaf(-v, r, 1.0f), r);
    r = fmaf(r, fmaf(-v, r, 1.0f), r);
    return r;
}

#ifndef MMVAE_LOGLIK_TILE_MIN_K
#define MMVAE_LOGLIK_TILE_MIN_K 2
#endif

template <int LT>
struct LogP {
    // the accumulated values are multiplied by this once per row (BCE accumulates in log2 units)
    static constexpr float kValueScale =
        (LT == MMVAE_LT_BCE || LT == MMVAE_LT_BCE_LOGITS) ? 0.69314718055994530942f : 1.0f;
    float c_inv, c_const;  // family specific constants
    __device__ __forceinline__ explicit LogP(float scale) {
        if (LT == MMVAE_LT_LPROB_NORMAL) {
            c_inv = 1.0f / (2.0f * scale * scale);
            c_const = -logf(scale) - 0.91893853320467274178f;  // log sqrt(2 pi)
        } else if (LT == MMVAE_LT_LPROB_LAPLACE) {
            c_inv = 1.0f / scale;
            c_const = -logf(2.0f * scale);
        } else {
            c_inv = 0.f;
            c_const = 0.f;
        }
    }
    // exact value for a sub-normal reconstruction 0 < x < 2^-126 (lg2.approx.ftz would flush it to log 0)
    __device__ __noinline__ static float bce_value_subnormal(float x, float t) {
        const float l1 = fmaxf(__log2f(1.0f - x), -144.26950408889634f);
        const float l0 = fmaxf(__log2f(x), -144.26950408889634f);
        return fmaf(t, l0 - l1, l1);
    }
    // value of log p(t | x) and its derivative w.r.t. x
    // LOWP: the derivative is rounded to bf16 by the caller (cheaper reciprocal, see rcp_newton2)
    template <bool NEED_V, bool NEED_D, bool LOWP = false>
    __device__ __forceinline__ void eval(float x, float t, float& v, float& d) const {
        if (LT == MMVAE_LT_BCE) {
            // F.binary_cross_entropy: log terms clamped at -100; backward denominator clamped at 1e-12
            // MUFU.LG2-based __logf: abs error <= 2^-21.4 on [0.5,2], ~2 ulp elsewhere; the precise logf costs ~3x
            // the instructions and made this pass compute bound (2.1 TB/s); -inf / NaN behave identically under
            // the clamp.  Error on a 12288-element row sum ~4e-9 relative (tests/test_ops_gpu.py checks 1e-5).
            // evaluated in log2 units (the row sum is rescaled by ln 2 once, see kValueScale) and blended with one
            // FFMA: 2 MUFU.LG2 + FADD + 2 FMNMX + FADD + FFMA per element
            if (NEED_V) {
                // 1-x is never sub-normal (0 or >= 2^-24); a sub-normal x (0 < x < 2^-126) takes the exact path
                const float l1 = fmaxf(lg2_ftz(1.0f - x), -144.26950408889634f);  // -100 / ln 2
                const float l0 = fmaxf(lg2_ftz(x), -144.26950408889634f);
                v = fmaf(t, l0 - l1, l1);
            }
            if (NEED_D) {  // denominator in [1e-12, 0.25]
                const float den = fmaxf((1.0f - x) * x, 1e-12f);
                d = (t - x) * (LOWP && MMVAE_BCE_BF16_NEWTON ? rcp_newton2(den) : rcp_ftz(den));
            }
        } else if (LT == MMVAE_LT_BCE_LOGITS) {
            // x holds the decoder logit y.  s = sigmoid(y), xc = clamp(s, lo, hi) with lo = fp32(1e-6), hi = fp32(1-1e-6)
            // (reference decoders.py:96-97), value = t log xc + (1-t) log(1-xc).  log is monotonic, so the clamp moves
            // onto the logs: log2 s = -log2(1+e) - max(-y,0) log2(e), log2(1-s) = -log2(1+e) - max(y,0) log2(e),
            // e = exp(-|y|): one EX2 + one LG2 per element instead of a sigmoid pass plus two logs.
            const float kL2E = 1.44269504088896340736f;
            const float e = exp2f(-fabsf(x) * kL2E);
            if (NEED_V) {
                const float l1pe = lg2_ftz(1.0f + e);
                const float yl = x * kL2E;
                // bounds: log2(lo), log2(hi) for xc and log2(1-hi), log2(1-lo) for 1-xc, all with fp32 operands
                // lo = fp32(1e-6) = 9.99999997e-07, hi = fp32(1-1e-6) = 0.999998987, 1-hi = 1.0132789e-06, 1-lo = hi
                const float l0 = fminf(fmaxf(-l1pe - fmaxf(-yl, 0.f), -19.93156857296663f), -1.4618532729665813e-06f);
                const float l1 = fminf(fmaxf(-l1pe - fmaxf(yl, 0.f), -19.91253715874966f), -1.4618532729665813e-06f);
                v = fmaf(t, l0 - l1, l1);
            }
            if (NEED_D) {
                const float r = rcp_ftz(1.0f + e);  // 1 + e in [1, 2]
                const float sg = x >= 0.f ? r : e * r;
                // clamp passes the gradient only inside [lo, hi]; there x(1-x) >= 1e-6 >> 1e-12, so
                // (t-x)/max((1-x)x,1e-12) * s(1-s) = t - s
                d = (sg >= 9.99999997e-07f && sg <= 0.999998987f) ? (t - sg) : 0.f;
            }
        } else if (LT == MMVAE_LT_LPROB_NORMAL) {
            const float df = t - x;
            const float lp = -df * df * c_inv + c_const;
            // NaN -> 0 for the value (objectives.py:423).  Gradient of a masked entry as torch's autograd leaves it:
            // the zeroed upstream gradient times the NaN local derivative (t-x)/sigma^2 is NaN for the Normal family,
            // while the Laplace family goes through sgn(NaN) == 0 and yields 0.  (The reference itself raises on NaN
            // inputs under torch's default validate_args; only reachable with validation off.)
            const bool bad = (lp != lp);
            if (NEED_V) v = bad ? 0.f : lp;
            if (NEED_D) d = bad ? lp : 2.0f * df * c_inv;
        } else if (LT == MMVAE_LT_LPROB_LAPLACE) {
            const float df = t - x;
            const float lp = c_const - fabsf(df) * c_inv;
            const bool bad = (lp != lp);
            if (NEED_V) v = bad ? 0.f : lp;
            if (NEED_D) d = bad ? 0.f : (df > 0.f ? c_inv : (df < 0.f ? -c_inv : 0.f));
        } else if (LT == MMVAE_LT_LPROB_NORMAL_SELF) {
            // reference quirk (objectives.py:43-45): with padding masks recon_loss_fn overwrites the likelihood's SCALE
            // with its (cropped) loc, so lprob evaluates Normal(loc = x, scale = x).log_prob(t).  log(x) of a negative
            // mean is NaN -> the value is zeroed (:423); the gradient of a zeroed entry is autograd's 0 * (finite local
            // derivative) = 0 (NaN only where the local derivative itself is not finite, x == 0).
            const float ix = 1.0f / x, u = (t - x) * ix;
            const float lp = -0.5f * u * u - logf(x) - 0.91893853320467274178f;
            const bool bad = (lp != lp);
            if (NEED_V) v = bad ? 0.f : lp;
            if (NEED_D) {
                const float dl = (u * t * ix - 1.0f) * ix;  // d/dloc + d/dscale = (t-x) t / x^3 - 1/x
                d = bad ? 0.f * dl : dl;
            }
        } else if (LT == MMVAE_LT_LPROB_LAPLACE_SELF) {
            // Laplace(loc = x, scale = x).log_prob(t) = -log(2x) - |t-x|/x (same quirk)
            const float ix = 1.0f / x, df = t - x;
            const float lp = -logf(2.0f * x) - fabsf(df) * ix;
            const bool bad = (lp != lp);
            if (NEED_V) v = bad ? 0.f : lp;
            if (NEED_D) {
                const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
                const float dl = (sg - 1.0f + fabsf(df) * ix) * ix;  // sgn(t-x)/x - 1/x + |t-x|/x^2
                d = bad ? 0.f * dl : dl;
            }
        } else if (LT == MMVAE_LT_MSE) {
            const float df = t - x;
            if (NEED_V) v = -df * df;
            if (NEED_D) d = 2.0f * df;
        } else {  // L1
            const float df = t - x;
            if (NEED_V) v = -fabsf(df);
            if (NEED_D) d = (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        }
    }
};

template <typename TX, typename TT, int LT, int MODE, bool VECT, bool SHORT>
__global__ void __launch_bounds__(kThreads, Tune<MODE>::kMinBlocks) loglik_kernel(const LoglikParams p) {
    const int tpr = SHORT ? p.tpr : kThreads;  // compile-time 256 for long rows: keeps the hot variant lean
    constexpr int V = VECT ? Elem<TX>::kPer16B : 1;
    constexpr int kUnroll = (sizeof(TX) == 2) ? 2 : Tune<MODE>::kUnroll;  // bf16 vectors carry 8 elements
    constexpr bool NEED_V = (MODE != MODE_BWD);
    constexpr bool NEED_D = (MODE != MODE_FWD);
    __shared__ float red[32];

    // short rows: tpr threads per row, 256/tpr rows per CTA (cpr == 1); long rows: cpr CTAs per row (tpr == 256)
    const int rpc = kThreads / tpr;
    const int sub = threadIdx.x / tpr, tin = threadIdx.x - sub * tpr;
    // Logical index -> row.  Rows are k-major (row = k*B + b) and the K rows of a sample share one target row; walking
    // rows in memory order re-reads a target larger than L2 from DRAM K times.  So the launch walks b-tiles of
    // `tile` targets (sized to stay L2 resident) and, inside a tile, all K samples: li -> (tile, k, b in tile).
    const int64_t li = (int64_t)(blockIdx.x / p.cpr) * rpc + sub;
    const bool row_ok = li < p.rows;
    int64_t row_raw = row_ok ? li : p.rows - 1;
    if (p.tile < p.B) {
        const int64_t K = p.rows / p.B;
        const int64_t bt = row_raw / ((int64_t)p.tile * K);
        const int64_t b0 = bt * p.tile;
        const int64_t tb = min((int64_t)p.tile, p.B - b0);  // the last tile may be short
        const int64_t rem = row_raw - b0 * K;
        const int64_t k = rem / tb;
        row_raw = k * p.B + b0 + (rem - k * tb);
    }
    const int64_t row = row_raw;
    const int chunk_id = blockIdx.x % p.cpr;
    // 32-bit offsets inside a row (P < 2^30 is checked on the host): halves the address arithmetic
    const int c0 = chunk_id * (int)p.chunk;
    const int c1 = row_ok ? min((int)p.P, c0 + (int)p.chunk) : c0;

    const TX* __restrict__ x = reinterpret_cast<const TX*>(p.x) + row * p.ldx;
    const TT* __restrict__ t = reinterpret_cast<const TT*>(p.t) + (row % p.B) * p.ldt;
    TX* __restrict__ g = NEED_D ? reinterpret_cast<TX*>(p.g) + row * p.ldg : nullptr;

    const LogP<LT> f(p.scale);
    float wl = 0.f;
    if (NEED_D) wl = (p.w_rows ? __ldg(p.w_rows + row) : p.w_const) * p.lam;

    float acc = 0.f;
    const int kStep = tpr * V;

    auto process = [&](const float* xv, const float* tv, int i) {
        float gv[V];
        if (MMVAE_BCE_FAST && LT == MMVAE_LT_BCE && NEED_V && V > 1) {
            // BCE value fast path (forward and fused passes).  The clamps max(log x, -100), max(log(1-x), -100) of F.binary_cross_entropy can
            // only fire for x == 1 (1-x == 0), x == 0 or a sub-normal x (log x < -100 needs x < 3.7e-44; lg2.ftz flushes
            // those to log 0); in each of these cases -- and for NaN / out-of-range inputs -- an unclamped log is -inf
            // or NaN and so is the sum over the vector.  So: evaluate the vector unclamped (2 MUFU.LG2 + 2 FADD + FFMA +
            // FADD per element), test the running sum ONCE per vector for finiteness, and redo the vector with the exact clamped
            // expression in the (for decoder outputs, clamped to [1e-6, 1-1e-6] by the reference, never taken) rare case.
            float part = acc;  // same summation order as the clamped evaluation: fwd rows == fused rows bit for bit
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float l1 = lg2_ftz(1.0f - xv[e]);
                part += fmaf(tv[e], lg2_ftz(xv[e]) - l1, l1);
            }
            if (__builtin_expect(fabsf(part) < INFINITY, 1)) {
                acc = part;
            } else {
                for (int e = 0; e < V; ++e) {
                    const float x = xv[e];
                    if (x > 0.f && x < 1.17549435e-38f) {
                        acc += LogP<LT>::bce_value_subnormal(x, tv[e]);
                    } else {  // NaN / negative inputs propagate exactly like the clamped expression
                        float v = 0.f, d = 0.f;
                        f.template eval<true, false>(x, tv[e], v, d);
                        acc += v;
                    }
                }
            }
            if (NEED_D) {  // fused pass: the gradient does not depend on the clamps of the value
#pragma unroll
                for (int e = 0; e < V; ++e) {
                    float v = 0.f, d = 0.f;
                    f.template eval<false, true, sizeof(TX) == 2>(xv[e], tv[e], v, d);
                    gv[e] = wl * d;
                }
                stg_stream(g + i, Elem<TX>::pack(gv));
            }
        } else {
#pragma unroll
        for (int e = 0; e < V; ++e) {
            float v = 0.f, d = 0.f;
            f.template eval<NEED_V, NEED_D, sizeof(TX) == 2>(xv[e], tv[e], v, d);
            if (NEED_V) acc += v;
            if (NEED_D) gv[e] = wl * d;
        }
        if (LT == MMVAE_LT_BCE && NEED_V) {
            // vector-level guard for sub-normal reconstructions (never taken for decoder outputs, which the
            // reference clamps to [1e-6, 1-1e-6]): redo those elements exactly
            float mn = xv[0];
#pragma unroll
            for (int e = 1; e < V; ++e) mn = fminf(mn, xv[e]);
            if (__builtin_expect(mn < 1.17549435e-38f, 0)) {
                for (int e = 0; e < V; ++e)
                    if (xv[e] > 0.f && xv[e] < 1.17549435e-38f) {
                        float v = 0.f, d = 0.f;
                        f.template eval<true, false>(xv[e], tv[e], v, d);
                        acc += LogP<LT>::bce_value_subnormal(xv[e], tv[e]) - v;
                    }
            }
        }
        if (NEED_D) {
            if (V == 1)
                Elem<TX>::store1(g + i, gv[0]);
            else
                stg_stream(g + i, Elem<TX>::pack(gv));
        }
        }
    };

    int base = c0 + tin * V;
    // main body: kUnroll vectors per tensor in flight, no per-vector bounds predicates
    for (; base + (kUnroll - 1) * kStep < c1; base += kStep * kUnroll) {
        float xv[kUnroll][V], tv[kUnroll][V];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            load_vec<TX, V>(x + base + u * kStep, xv[u], true);
            load_vec<TT, V>(t + base + u * kStep, tv[u], false);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) process(xv[u], tv[u], base + u * kStep);
    }
    if (base < c1) {
        // epilogue: the remaining (< kUnroll) vectors, all loads issued before the first use (a serial tail with
        // one load in flight per iteration cost short rows ~35% of their bandwidth)
        float xv[kUnroll][V], tv[kUnroll][V];
#pragma unroll
        for (int u = 0; u < kUnroll - 1; ++u) {
            if (base + u * kStep < c1) {
                load_vec<TX, V>(x + base + u * kStep, xv[u], true);
                load_vec<TT, V>(t + base + u * kStep, tv[u], false);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll - 1; ++u)
            if (base + u * kStep < c1) process(xv[u], tv[u], base + u * kStep);
    }
    if (NEED_V) {
        // reduce over the tpr threads of a row: warp shuffles, then (tpr > 32) the row's warps through smem
        float tot = warp_sum(acc);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (tpr > 32) {
            if (lane == 0) red[wid] = tot;
            __syncthreads();
            const int wpr = tpr >> 5;  // warps per row
            if (tin < 32) {
                float v = tin < wpr ? red[sub * wpr + tin] : 0.f;
                tot = warp_sum(v);
            }
        }
        tot *= LogP<LT>::kValueScale;
        if (tin == 0 && row_ok) {
            if (p.cpr == 1)
                p.out_rows[row] = p.lam * tot;
            else
                p.ws[row * p.cpr + chunk_id] = tot;
        }
    }
}

// second stage of the deterministic row sum when a row was split over cpr CTAs
__global__ void loglik_finalize_kernel(const float* __restrict__ ws, float* __restrict__ out, int64_t rows, int cpr,
                                       float lam) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float s = 0.f;
    for (int c = 0; c < cpr; ++c) s += ws[r * cpr + c];
    out[r] = lam * s;
}

static int plan_tpr(int64_t P, int V, int cpr) {
    if (cpr > 1) return kThreads;
    // aim at >= kUnrollMax vectors per thread so the unrolled main loop is entered (measured r1, P = 3072:
    // 128 threads/row -> 4.7 TB/s fwd, 256 threads/row with a single round trip -> 3.6 TB/s)
    int64_t want = P / ((int64_t)V * kUnrollMax);
    int tpr = 32;
    while (tpr * 2 <= want && tpr < kThreads) tpr <<= 1;
    return tpr;
}

static void plan(int64_t rows, int64_t P, int V, int64_t* chunk, int* cpr) {
    const int64_t step = (int64_t)kThreads * V;            // elements per CTA sweep
    const int64_t min_chunk = step * kUnrollMax;           // one fully unrolled iteration
    const int64_t max_cpr = (P + min_chunk - 1) / min_chunk;
    int64_t want = ((int64_t)kNumSMs * 16 + rows - 1) / rows;  // aim at >= 16 CTAs per SM worth of work
    int64_t c = want < 1 ? 1 : want;
    if (c > max_cpr) c = max_cpr;
    if (c < 1) c = 1;
    int64_t ch = (P + c - 1) / c;
    ch = (ch + step - 1) / step * step;
    c = (P + ch - 1) / ch;
    *chunk = ch;
    *cpr = (int)c;
}

template <typename TX, typename TT, int LT, int MODE>
static int launch3(const LoglikParams& p, bool vect, cudaStream_t st) {
    const int rpc = kThreads / p.tpr;
    const int64_t grid = p.cpr > 1 ? p.rows * p.cpr : (p.rows + rpc - 1) / rpc;
    if (grid > 0x7fffffffLL) return MMVAE_E_LIMIT;
    const bool shortrows = p.tpr != kThreads;
    if (vect && !shortrows)
        loglik_kernel<TX, TT, LT, MODE, true, false><<<(unsigned)grid, kThreads, 0, st>>>(p);
    else if (vect)
        loglik_kernel<TX, TT, LT, MODE, true, true><<<(unsigned)grid, kThreads, 0, st>>>(p);
    else
        loglik_kernel<TX, TT, LT, MODE, false, true><<<(unsigned)grid, kThreads, 0, st>>>(p);
    MMVAE_LAUNCH_CHECK();
    if (MODE != MODE_BWD && p.cpr > 1) {
        loglik_finalize_kernel<<<(unsigned)((p.rows + 255) / 256), 256, 0, st>>>(p.ws, p.out_rows, p.rows, p.cpr, p.lam);
        MMVAE_LAUNCH_CHECK();
    }
    return 0;
}

template <typename TX, typename TT, int MODE>
static int launch2(const LoglikParams& p, int ltype, bool vect, cudaStream_t st) {
    switch (ltype) {
        case MMVAE_LT_BCE: return launch3<TX, TT, MMVAE_LT_BCE, MODE>(p, vect, st);
        case MMVAE_LT_LPROB_NORMAL: return launch3<TX, TT, MMVAE_LT_LPROB_NORMAL, MODE>(p, vect, st);
        case MMVAE_LT_LPROB_LAPLACE: return launch3<TX, TT, MMVAE_LT_LPROB_LAPLACE, MODE>(p, vect, st);
        case MMVAE_LT_MSE: return launch3<TX, TT, MMVAE_LT_MSE, MODE>(p, vect, st);
        case MMVAE_LT_L1: return launch3<TX, TT, MMVAE_LT_L1, MODE>(p, vect, st);
        case MMVAE_LT_BCE_LOGITS: return launch3<TX, TT, MMVAE_LT_BCE_LOGITS, MODE>(p, vect, st);
        case MMVAE_LT_LPROB_NORMAL_SELF: return launch3<TX, TT, MMVAE_LT_LPROB_NORMAL_SELF, MODE>(p, vect, st);
        case MMVAE_LT_LPROB_LAPLACE_SELF: return launch3<TX, TT, MMVAE_LT_LPROB_LAPLACE_SELF, MODE>(p, vect, st);
        default: return MMVAE_E_ENUM;
    }
}

template <int MODE>
static int launch(LoglikParams p, int dtx, int dtt, int ltype, cudaStream_t st) {
    if (!p.x || !p.t || p.rows <= 0 || p.B <= 0 || p.P <= 0) return MMVAE_E_ARG;
    if (p.P >= (1LL << 30)) return MMVAE_E_LIMIT;
    if (MODE != MODE_BWD && !p.out_rows) return MMVAE_E_ARG;
    if (MODE != MODE_FWD && !p.g) return MMVAE_E_ARG;
    if (MODE == MODE_BWD && !p.w_rows) return MMVAE_E_ARG;
    if (p.ldx < p.P || p.ldt < p.P || (MODE != MODE_FWD && p.ldg < p.P)) return MMVAE_E_ARG;
    const int sx = dtx == MMVAE_F32 ? 4 : 2, stt = dtt == MMVAE_F32 ? 4 : 2;
    const int V = 16 / sx;
    bool vect = (p.P % V == 0) && aligned16(p.x) && aligned16(p.t) && (p.ldx * sx) % 16 == 0 &&
                (p.ldt * stt) % 16 == 0;
    if (MODE != MODE_FWD) vect = vect && aligned16(p.g) && (p.ldg * sx) % 16 == 0;
    plan(p.rows, p.P, vect ? V : 1, &p.chunk, &p.cpr);
    p.tpr = plan_tpr(p.P, vect ? V : 1, p.cpr);
    {   // L2 tiling of the target: keep a tile of targets (<= 16 MB) resident while its K sample rows stream by
        const int64_t tile_rows = (16LL << 20) / (p.P * stt);
        // (tuning knob, read per call: the tiled order only pays when the target is re-read often enough)
        static const int tile_min_k = [] { const char* e = getenv("MMVAE_LOGLIK_TILE_MIN_K"); return e ? atoi(e) : MMVAE_LOGLIK_TILE_MIN_K; }();
        const bool tiled = p.rows > p.B && tile_rows < p.B && p.rows / p.B >= tile_min_k;
        p.tile = tiled ? (int)(tile_rows < 1 ? 1 : tile_rows) : (int)(p.B > 0x7fffffff ? 0x7fffffff : p.B);
    }
    if (MODE != MODE_BWD && p.cpr > 1 && !p.ws) return MMVAE_E_ARG;
    if (dtx == MMVAE_F32 && dtt == MMVAE_F32) return launch2<float, float, MODE>(p, ltype, vect, st);
    if (dtx == MMVAE_BF16 && dtt == MMVAE_BF16) return launch2<__nv_bfloat16, __nv_bfloat16, MODE>(p, ltype, vect, st);
    if (dtx == MMVAE_BF16 && dtt == MMVAE_F32) return launch2<__nv_bfloat16, float, MODE>(p, ltype, vect, st);
    if (dtx == MMVAE_F32 && dtt == MMVAE_BF16) return launch2<float, __nv_bfloat16, MODE>(p, ltype, vect, st);
    return MMVAE_E_ENUM;
}

}  // namespace mmvae

using namespace mmvae;

extern "C" int64_t mmvae_loglik_workspace_bytes(int64_t rows, int64_t P, int dtype_recon) {
    if (rows <= 0 || P <= 0) return 0;
    // worst case over the vector / scalar paths: the scalar path (V = 1) splits rows the finest
    int64_t chunk;
    int cpr_v, cpr_s;
    plan(rows, P, dtype_recon == MMVAE_F32 ? 4 : 8, &chunk, &cpr_v);
    plan(rows, P, 1, &chunk, &cpr_s);
    const int cpr = cpr_v > cpr_s ? cpr_v : cpr_s;
    return cpr > 1 ? rows * cpr * (int64_t)sizeof(float) : 0;
}

extern "C" int mmvae_loglik_rowreduce_fwd(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                          int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P,
                                          int ltype, float scale, float lam, float* out_rows, void* workspace,
                                          void* stream) {
    LoglikParams p{};
    p.x = recon; p.t = target; p.ldx = ld_recon; p.ldt = ld_target; p.rows = rows; p.B = B; p.P = P;
    p.scale = scale; p.lam = lam; p.out_rows = out_rows; p.ws = (float*)workspace;
    return launch<MODE_FWD>(p, dtype_recon, dtype_target, ltype, (cudaStream_t)stream);
}

extern "C" int mmvae_loglik_rowreduce_bwd(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                          int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P,
                                          int ltype, float scale, float lam, const float* w_rows, void* grad_recon,
                                          int64_t ld_grad, void* stream) {
    LoglikParams p{};
    p.x = recon; p.t = target; p.ldx = ld_recon; p.ldt = ld_target; p.rows = rows; p.B = B; p.P = P;
    p.scale = scale; p.lam = lam; p.w_rows = w_rows; p.g = grad_recon; p.ldg = ld_grad;
    return launch<MODE_BWD>(p, dtype_recon, dtype_target, ltype, (cudaStream_t)stream);
}

extern "C" int mmvae_loglik_rowreduce_fused(const void* recon, int64_t ld_recon, int dtype_recon, const void* target,
                                            int64_t ld_target, int dtype_target, int64_t rows, int64_t B, int64_t P,
                                            int ltype, float scale, float lam, const float* w_rows, float w_const,
                                            float* out_rows, void* grad_recon, int64_t ld_grad, void* workspace,
                                            void* stream) {
    LoglikParams p{};
    p.x = recon; p.t = target; p.ldx = ld_recon; p.ldt = ld_target; p.rows = rows; p.B = B; p.P = P;
    p.scale = scale; p.lam = lam; p.w_rows = w_rows; p.w_const = w_const; p.out_rows = out_rows;
    p.g = grad_recon; p.ldg = ld_grad; p.ws = (float*)workspace;
    return launch<MODE_FUSED>(p, dtype_recon, dtype_target, ltype, (cudaStream_t)stream);
}
