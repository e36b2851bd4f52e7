// Microbenchmark: L1 throughput of divergent 8-byte / 16-byte gathers on sm_100a as a function of how the 32 lanes of a
// request spread over sectors / 128-B lines.  Working set 64 KB per SM (L1 resident) so the L1 data pipe is isolated.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

// mode: lines = number of distinct 128-B lines per request (32,16,8,4,1); spread: 0 = lanes of a line group adjacent rows
// (same sector when group<=4 for 8-B), 1 = lanes of a group in different sectors of the line
template <int BYTES>
__global__ void __launch_bounds__(512) k(const uint8_t* __restrict__ buf, uint32_t lines_per_req, uint32_t spread, int iters, float* out) {
    const int lane = threadIdx.x & 31;
    const uint32_t ws_lines = 512;  // 64 KB working set
    const uint32_t group = 32 / lines_per_req, gid = lane / group, within = lane % group;
    uint32_t s = (blockIdx.x * 512 + (threadIdx.x & ~31)) * 2654435761u + 12345u;  // per-warp stream (same for all lanes)
    float acc = 0.f;
    uint32_t offs[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {   // every lane draws the same numbers; lane's line = draw[gid]
        uint32_t ss = s + j * 977u, line = 0;
        for (uint32_t g = 0; g <= gid; g++) line = rng(ss);
        line %= ws_lines;
        const uint32_t slot = spread ? (within * (128 / BYTES / group)) : within;
        offs[j] = line * 128 + (slot % (128 / BYTES)) * BYTES;
    }
    for (int it = 0; it < iters; it++) {
        const uint32_t rot = (uint32_t)(it * 37) % ws_lines * 128;   // same rotation for all lanes: keeps the line structure
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t o = (offs[j] + rot) & (ws_lines * 128 - 1);
            if (BYTES == 8) { float2 v = __ldg(reinterpret_cast<const float2*>(buf + o)); acc += v.x + v.y; }
            else { float4 v = __ldg(reinterpret_cast<const float4*>(buf + o)); acc += v.x + v.w; }
        }
    }
    out[blockIdx.x * 512 + threadIdx.x] = acc;
}

template <int BYTES>
void run(const uint8_t* buf, float* out, uint32_t lines, uint32_t spread) {
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<BYTES><<<148, 512>>>(buf, lines, spread, 50, out);
    cudaEventRecord(e0);
    k<BYTES><<<148, 512>>>(buf, lines, spread, iters, out);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double req_per_sm = 16.0 * iters * 8;
    printf("LDG.%-3d lines/request=%2u %-22s : %6.2f cycles/request/SM (@1.965GHz)   %.3f ms\n", BYTES * 8, lines,
           spread ? "(lanes spread in line)" : "(lanes adjacent)", ms * 1e-3 * 1.965e9 / req_per_sm, ms);
}

int main() {
    uint8_t* buf; float* out;
    cudaMalloc(&buf, 1 << 20); cudaMemset(buf, 0, 1 << 20); cudaMalloc(&out, 148 * 512 * 4);
    for (uint32_t lines : {32u, 16u, 8u, 4u, 2u, 1u}) { run<8>(buf, out, lines, 0); if (lines < 32) run<8>(buf, out, lines, 1); }
    for (uint32_t lines : {32u, 16u, 8u, 4u, 1u}) { run<16>(buf, out, lines, 0); }
    return 0;
}
