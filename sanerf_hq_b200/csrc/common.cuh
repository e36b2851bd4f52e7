// common.cuh -- shared device helpers for libsanerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sanerf_b200.h"

#ifndef __CUDA_ARCH__
#define SANERF_ARCH_OK 1
#elif __CUDA_ARCH__ >= 1000
#define SANERF_ARCH_OK 1
#else
#error "libsanerf_b200 is written for sm_100a (B200) only"
#endif

namespace sanerf {

__host__ __device__ inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// The XOR-of-products spatial hash of the reference (gridencoder.cu:45-58): uint32 wrap-around.
template <uint32_t D>
__device__ __forceinline__ uint32_t spatial_hash(const uint32_t (&p)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (uint32_t i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

// Row index of a grid vertex inside one level (gridencoder.cu:61-79): dense stride sum while the
// running stride still fits the level's row count, otherwise the hash; always reduced mod rows.
template <uint32_t D>
__device__ __forceinline__ uint32_t vertex_row(uint32_t gridtype, uint32_t rows, uint32_t res, const uint32_t (&p)[D]) {
    uint32_t stride = 1, index = 0;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (stride <= rows) {
            index += p[d] * stride;
            stride *= res;
        }
    }
    if (gridtype == 0 && stride > rows) index = spatial_hash<D>(p);
    return index % rows;
}

// Kernel-side level resolution, evaluated on the device exactly like the reference (gridencoder.cu:133).
__device__ __forceinline__ uint32_t level_resolution(uint32_t level, float S, uint32_t H) {
    return (uint32_t)ceilf(exp2f((float)level * S) * (float)H);
}

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

}  // namespace sanerf
