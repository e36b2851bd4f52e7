// Microbenchmark: issue rate of legacy warp-level mma.sync on sm_100a (tf32 m16n8k8, bf16 m16n8k16).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

template <int KIND, int ILP>
__global__ void __launch_bounds__(512) k(float* out, int iters) {
    float c[ILP][4];
    uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 0x3f800000u, 0x3f000000u}, b[2] = {0x3f800000u, threadIdx.x};
#pragma unroll
    for (int i = 0; i < ILP; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int ILP>
void run(const char* name, int warps) {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, ILP><<<148, warps * 32>>>(out, 100);
    cudaEventRecord(e0);
    k<KIND, ILP><<<148, warps * 32>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double n = (double)148 * warps * iters * ILP;
    double flop = n * (KIND == 0 ? 2048.0 : 4096.0);
    printf("%s warps/SM=%d ILP=%d: %.3f ms, %.1f TFLOP/s, %.2f cycles@1.965GHz per mma per SMSP\n", name, warps, ILP, ms, flop / ms / 1e9,
           ms * 1e-3 * 1.965e9 / (iters * ILP * warps / 4.0));
    cudaFree(out);
}

int main() {
    run<0, 4>("tf32 m16n8k8 ", 4); run<0, 8>("tf32 m16n8k8 ", 16); run<0, 8>("tf32 m16n8k8 ", 8);
    run<1, 4>("bf16 m16n8k16", 4); run<1, 8>("bf16 m16n8k16", 16);
    return 0;
}
