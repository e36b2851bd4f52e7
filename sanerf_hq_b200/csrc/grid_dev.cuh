// grid_dev.cuh -- device-side description of a multiresolution hash grid + the gather primitives shared by the fused render
// kernel (render.cu) and the tensor-core heads (heads.cu).
#pragma once
#include "common.cuh"

#ifndef SANERF_FFMA2
#define SANERF_FFMA2 1
#endif

namespace sanerf {

constexpr unsigned kFullMask = 0xffffffffu;

// ---- device-side model description (kernel parameter, constant bank) -------------------------
struct GridDev {
    const float* emb;
    uint32_t L, C;
    uint32_t off[SANERF_MAX_LEVELS];    // row offset of the level
    uint32_t res[SANERF_MAX_LEVELS];    // kernel-side resolution
    uint32_t hmask[SANERF_MAX_LEVELS];  // rows-1 for hashed levels (rows is a power of two), 0 = dense
    const void* base[SANERF_MAX_LEVELS];  // emb + off[l]*C: first row of the level (saves the per-load offset add)
    float resf[SANERF_MAX_LEVELS];        // (float)res, (float)(res-1)
    float topf[SANERF_MAX_LEVELS];
};

// floor of a clamped grid coordinate 0 <= pos < 2^22 without the conversion unit (FRND / F2I run on the quarter-rate XU pipe):
// pos + 2^23 rounded toward zero has an ulp of 1, so its low mantissa bits ARE floor(pos); both results are exact.
__device__ __forceinline__ void floor_split(float pos, uint32_t& cell, float& frac) {
    const float t = __fadd_rz(pos, 8388608.0f);
    cell = __float_as_uint(t) & 0x007fffffu;
    frac = pos - (t - 8388608.0f);
}

// Packed FP32 FMA of sm_100a (fma.rn.f32x2 -> SASS FFMA2 with a scalar-broadcast operand): acc.{x,y} = fma(w, v.{x,y}, acc.{x,y}),
// bit-identical to the two scalar FMAs it replaces (IEEE fma per component).  The FMA pipe needs two cycles for it
// (tools/ffma2_rate.cu: 2.07 vs 2.16 cycles per pair), but it takes ONE issue slot instead of two -- the trilinear blends are where
// the render kernels spend a sixth of their issue slots.
struct Acc2 {
    unsigned long long bits;
    __device__ __forceinline__ Acc2() : bits(0ull) {}     // {0.f, 0.f}
    __device__ __forceinline__ void fma(float w, float2 v) {
#if SANERF_FFMA2
        unsigned long long wv, vv;
        asm("mov.b64 %0, {%1,%1};" : "=l"(wv) : "f"(w));
        asm("mov.b64 %0, {%1,%2};" : "=l"(vv) : "f"(v.x), "f"(v.y));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(bits) : "l"(wv), "l"(vv));
#else
        float a0, a1;
        get(a0, a1);
        a0 = __fmaf_rn(w, v.x, a0);
        a1 = __fmaf_rn(w, v.y, a1);
        asm("mov.b64 %0, {%1,%2};" : "=l"(bits) : "f"(a0), "f"(a1));
#endif
    }
    __device__ __forceinline__ void get(float& a0, float& a1) const { asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(bits)); }
};

// Packed pairs for code that runs the same scalar recipe on two independent items (the two sample chunks of a proposal round, the
// two passes of the x-paired final stage): one issue slot per pair of results, each component the same IEEE operation as the
// scalar code (add / sub / mul / fma .rn, add .rz).
struct F2 {
    unsigned long long b;
    __device__ __forceinline__ F2() {}
    __device__ __forceinline__ F2(float x, float y) { asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(x), "f"(y)); }
    __device__ __forceinline__ float x() const { return __uint_as_float((uint32_t)b); }
    __device__ __forceinline__ float y() const { return __uint_as_float((uint32_t)(b >> 32)); }
};
__device__ __forceinline__ F2 f2_fma(F2 a, F2 m, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.b) : "l"(a.b), "l"(m.b), "l"(c.b)); return r; }
__device__ __forceinline__ F2 f2_mul(F2 a, F2 m) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.b) : "l"(a.b), "l"(m.b)); return r; }
__device__ __forceinline__ F2 f2_sub(F2 a, F2 m) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.b) : "l"(a.b), "l"(m.b)); return r; }
__device__ __forceinline__ F2 f2_add_rz(F2 a, F2 m) { F2 r; asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r.b) : "l"(a.b), "l"(m.b)); return r; }

// ---- C=8 feature grids: quarter-row gathers -------------------------------------------------------------------------
// A C=8 row is 32 bytes.  "lane = sample, two LDG.128 per corner" costs 2 L1 tag cycles per distinct 128-byte line per
// request (tools/l1_gather.cu), with up to 32 lines per request.  Here 4 lanes share a sample, each fetching 8 bytes (2 of
// the 8 channels) of every corner row: a request covers 8 samples x 4 quarter rows = 8 lines at 1 cycle per line.
// One level of grid g at point x: this lane's channel pair (2*part, 2*part+1), blended in the reference kernel's corner order.
__device__ __forceinline__ void quarter_level(const GridDev& g, int l, const float (&x)[3], int part, float& o0, float& o1) {
    const uint32_t res = g.res[l];
    const uint32_t hmask = g.hmask[l];
    const float2* __restrict__ rows = reinterpret_cast<const float2*>(g.base[l]) + part;   // 4 float2 per row
    const float resf = g.resf[l], top = g.topf[l];
    uint32_t b0[3], b1[3];
    float f[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        float pos = fminf(fmaxf(__fmaf_rn(x[d], resf, -0.5f), 0.0f), top);
        floor_split(pos, b0[d], f[d]);
        b1[d] = min(b0[d] + 1, res - 1);
    }
    float2 v[8];
    if (hmask == 0) {
        const uint32_t y0 = b0[1] * res, y1 = b1[1] * res, z0 = b0[2] * res * res, z1 = b1[2] * res * res;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(rows + 4 * (((i & 1) ? b1[0] : b0[0]) + ((i & 2) ? y1 : y0) + ((i & 4) ? z1 : z0)));
    } else {
        const uint32_t x0 = b0[0] & hmask, x1 = b1[0] & hmask;
        const uint32_t y0 = (b0[1] * 2654435761u) & hmask, y1 = (b1[1] * 2654435761u) & hmask;
        const uint32_t z0 = (b0[2] * 805459861u) & hmask, z1 = (b1[2] * 805459861u) & hmask;
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(rows + 4 * (((i & 1) ? x1 : x0) ^ ((i & 2) ? y1 : y0) ^ ((i & 4) ? z1 : z0)));
    }
    Acc2 acc;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        float ww = (i & 1) ? f[0] : 1 - f[0];
        ww *= (i & 2) ? f[1] : 1 - f[1];
        ww *= (i & 4) ? f[2] : 1 - f[2];
        acc.fma(ww, v[i]);
    }
    acc.get(o0, o1);
}

// host: sanerf_grid_t (C ABI) -> GridDev; also decides dense vs hashed per level exactly like the reference kernel
inline int fill_grid(GridDev& g, const sanerf_grid_t& s, uint32_t C_expected) {
    g.emb = s.embeddings;
    g.L = s.num_levels;
    g.C = s.level_dim;
    if (!s.embeddings || s.num_levels == 0 || s.num_levels > SANERF_MAX_LEVELS || s.level_dim != C_expected) return SANERF_E_CONFIG;
    for (uint32_t l = 0; l < s.num_levels; l++) {
        const uint32_t rows = s.offset[l + 1] - s.offset[l], res = s.res[l];
        if (res < 2 || rows == 0) return SANERF_E_CONFIG;
        // the reference's dense-vs-hash decision (gridencoder.cu:61-79) for D=3, gridtype hash
        uint64_t stride = 1;
        for (int d = 0; d < 3 && stride <= rows; d++) stride *= res;
        g.off[l] = s.offset[l];
        g.res[l] = res;
        g.base[l] = s.embeddings + (size_t)s.offset[l] * s.level_dim;
        g.resf[l] = (float)res;
        g.topf[l] = (float)(res - 1);
        if (stride <= rows) {
            g.hmask[l] = 0;  // dense, index < res^3 <= rows
        } else {
            if (rows & (rows - 1)) return SANERF_E_CONFIG;  // hashed levels have 2^T rows (grid.py:129)
            g.hmask[l] = rows - 1;
        }
    }
    for (uint32_t l = s.num_levels; l < SANERF_MAX_LEVELS; l++) {
        g.off[l] = g.res[l] = g.hmask[l] = 0;
        g.base[l] = nullptr;
        g.resf[l] = g.topf[l] = 0.f;
    }
    return 0;
}


}  // namespace sanerf
