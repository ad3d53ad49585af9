// r2 microbenchmark: does packed fp32 (FFMA2, fma.rn.f32x2) free issue slots on B200?
//   A: 8 independent scalar FFMA chains          B: 4 independent FFMA2 chains (same flops)
//   C: A + 8 independent integer LOP3/IADD chains  D: B + the same integer chains
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench tools/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) { unsigned r; asm volatile("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float m, float c) {
    float x[8]; u64 y[4]; unsigned q[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; q[i] = threadIdx.x + i; }
    for (int i = 0; i < 4; ++i) y[i] = pack(x[2 * i], x[2 * i + 1]);
    const u64 m2 = pack(m, m), c2 = pack(c, c);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fma1(x[i], m, c);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) y[i] = fma2(y[i], m2, c2);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = lop(q[i], 0x9e3779b9u + i);
        }
    }
    float s = 0; unsigned t = 0;
    for (int i = 0; i < 8; ++i) { s += x[i]; t += q[i]; }
    for (int i = 0; i < 4; ++i) { float a, b; unpack(y[i], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}
template <int MODE>
float run(float* out, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 4, 512>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(a);
    k<MODE><<<148 * 4, 512>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 512 * 4);
    const int iters = 20000;
    const double flops = 2.0 * 8 * iters * 148 * 4 * 512;
    float a = run<0>(out, iters), b = run<1>(out, iters), c = run<2>(out, iters), d = run<3>(out, iters);
    printf("A scalar FFMA x8      : %.3f ms  %.1f TFLOP/s\n", a, flops / a / 1e9);
    printf("B FFMA2 x4            : %.3f ms  %.1f TFLOP/s\n", b, flops / b / 1e9);
    printf("C scalar FFMA x8 + 8 LOP3: %.3f ms\n", c);
    printf("D FFMA2 x4 + 8 LOP3      : %.3f ms\n", d);
    return 0;
}
