// Shared device helpers for the mmvae_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mmvae_b200.h"

#define MMVAE_LAUNCH_CHECK()                     \
    do {                                         \
        cudaError_t e__ = cudaGetLastError();    \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

namespace mmvae {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). Result valid in thread 0 (and warp 0).
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem /* >= 32 entries */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T r = (threadIdx.x < nw) ? smem[threadIdx.x] : T(0);
    if (wid == 0) r = warp_sum(r);
    __syncthreads();
    return r;
}

// ---- 128-bit streaming loads/stores -------------------------------------------------------------------
// Reconstructions and their gradients are touched exactly once per pass: bypass L1 allocation so that the
// (much smaller, K-times reused) targets keep the cache.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_keep(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t ldg_stream_u32(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_u32(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

template <typename T>
struct Elem;
template <>
struct Elem<float> {
    static constexpr int kPer16B = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, float* o) {
        o[0] = __uint_as_float(v.x);
        o[1] = __uint_as_float(v.y);
        o[2] = __uint_as_float(v.z);
        o[3] = __uint_as_float(v.w);
    }
    __device__ static __forceinline__ uint4 pack(const float* o) {
        return make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
    }
    __device__ static __forceinline__ float load1(const float* p) { return __ldg(p); }  // GLOBAL memory only
    __device__ static __forceinline__ float get(const float* p) { return *p; }          // any address space
    __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int kPer16B = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, float* o) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[2 * i] = __uint_as_float(w[i] << 16);
            o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static __forceinline__ uint4 pack(const float* o) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ float get(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// V elements of T -> fp32 registers: 128-bit loads (streaming or cached), 8-byte load for 4 bf16 next to fp32 data
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float* o, bool stream) {
    if (V == 1) {
        o[0] = Elem<T>::load1(p);
    } else {
        constexpr int per = Elem<T>::kPer16B;
        if (V >= per) {
#pragma unroll
            for (int i = 0; i < V / per; ++i) {
                const uint4 v = stream ? ldg_stream(p + i * per) : ldg_keep(p + i * per);
                Elem<T>::unpack(v, o + i * per);
            }
        } else {  // bf16 target next to an fp32 reconstruction: 4 x bf16 = 8 bytes
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
            o[0] = __uint_as_float(v.x << 16);
            o[1] = __uint_as_float(v.x & 0xffff0000u);
            o[2] = __uint_as_float(v.y << 16);
            o[3] = __uint_as_float(v.y & 0xffff0000u);
        }
    }
}

// Second stage of the deterministic batch reductions (learnable-prior gradients): ws holds `parts` partial vectors
// of length n_total (row-major); output element i = sum_q ws[q][i].  One CTA per output element, strided loads,
// fixed-shape block reduction -> bit-reproducible; out0 receives elements [0, n0), out1 the rest.
static __global__ void __launch_bounds__(128) partial_sum_kernel(const float* __restrict__ ws, int parts, int n_total,
                                                                 int n0, float* __restrict__ out0,
                                                                 float* __restrict__ out1) {
    __shared__ float red[32];
    const int i = blockIdx.x;
    float acc = 0.f;
    for (int q = threadIdx.x; q < parts; q += blockDim.x) acc += ws[(size_t)q * n_total + i];
    const float tot = block_sum(acc, red);
    if (threadIdx.x == 0) {
        if (i < n0) out0[i] = tot;
        else out1[i - n0] = tot;
    }
}

// exp(x - m) for softmax-style sums as ONE FFMA + ONE MUFU.EX2: ex2.approx.ftz(x*log2(e) - m*log2(e)).  (__expf
// expands to 6 instructions: scale, range test, two predicated fix-ups around MUFU for sub-normal RESULTS; results
// below 2^-126 are flushed to 0 here, which is irrelevant next to a sum whose largest term is 1.)
constexpr float kLog2e = 1.44269504088896340736f;
__device__ __forceinline__ float ex2_ftz(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float exp_shifted(float x, float neg_m_log2e) { return ex2_ftz(fmaf(x, kLog2e, neg_m_log2e)); }

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2) -------------------------------------------------------------
// sm_100 executes fp32 add / mul / fma on TWO values held in an aligned register pair with one instruction
// (PTX {add,sub,mul,fma}.f32x2).  The FMA pipe needs two cycles for it, so the FLOP rate does not change (r2
// microbenchmark tools/ffma2_bench.cu: 72 vs 74 TFLOP/s) -- but an issue-bound kernel gets the second issue slot back
// for its loads, MUFU and integer work.  ptxas folds the sign-bit tricks below (abs / neg) into operand modifiers.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 f2_pack(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 f2_bcast(float v) { return f2_pack(v, v); }
__device__ __forceinline__ void f2_unpack(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ float f2_lo(f32x2 v) { float a, b; f2_unpack(v, a, b); return a; }
__device__ __forceinline__ float f2_hi(f32x2 v) { float a, b; f2_unpack(v, a, b); return b; }
__device__ __forceinline__ f32x2 f2_fma(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 f2_add(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_sub(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_mul(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 f2_abs(f32x2 a) { return a & 0x7fffffff7fffffffull; }
__device__ __forceinline__ f32x2 f2_neg(f32x2 a) { return a ^ 0x8000000080000000ull; }

// ---- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__host__ __device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace mmvae
