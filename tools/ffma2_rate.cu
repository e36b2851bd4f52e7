// Microbenchmark: issue rate of the packed FP32 FMA of sm_100a (fma.rn.f32x2 -> SASS FFMA2, scalar weight broadcast) against
// the scalar FFMA pair it replaces in the trilinear blends (acc.x += w*v.x; acc.y += w*v.y).
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float2& acc, float w, float2 v) {
    unsigned long long a, b, c;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1,%1};" : "=l"(c) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a) : "l"(c), "l"(b));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(a));
}

template <int PACKED>
__global__ void __launch_bounds__(512) k(float2* out, int iters, float w0) {
    float2 acc[8], v[8];
    for (int i = 0; i < 8; i++) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); v[i] = make_float2(1.0f + i * 1e-3f, 0.999f - i * 1e-3f); }
    float w = w0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (PACKED) fma2(acc[i], w, v[i]);
            else { acc[i].x = __fmaf_rn(w, v[i].x, acc[i].x); acc[i].y = __fmaf_rn(w, v[i].y, acc[i].y); }
        }
    }
    float2 s = make_float2(0, 0);
    for (int i = 0; i < 8; i++) { s.x += acc[i].x; s.y += acc[i].y; }
    out[blockIdx.x * 512 + threadIdx.x] = s;
}

int main() {
    float2* out; cudaMalloc(&out, 148 * 512 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; packed++) {
        const int iters = 20000;
        if (packed) k<1><<<148, 512>>>(out, 100, 0.999f); else k<0><<<148, 512>>>(out, 100, 0.999f);
        cudaEventRecord(e0);
        if (packed) k<1><<<148, 512>>>(out, iters, 0.999f); else k<0><<<148, 512>>>(out, iters, 0.999f);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double fmas = 148.0 * 512 * iters * 16;   // scalar FMAs' worth of work
        printf("%s: %.3f ms, %.1f TFLOP/s (fp32 FMA = 2 flop), %.2f cycles per warp-level pair of FMAs per SMSP\n", packed ? "FFMA2 (f32x2)" : "FFMA x2     ", ms,
               2 * fmas / ms / 1e9, ms * 1e-3 * 1.965e9 / (iters * 8.0 * 4));
    }
    return 0;
}
