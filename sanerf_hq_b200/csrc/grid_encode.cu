// grid_encode.cu -- multiresolution hash / tiled grid encoder operators (sm_100a).
//
// Replaces the reference's gridencoder/src/gridencoder.cu kernels K1..K5 behind
// the C ABI of include/sanerf_b200.h.  Same arithmetic (corner order, fp32 FMA accumulation,
// uint32 hash, device-evaluated level resolution) so results match the reference kernels; the
// implementation differs: vectorised row gathers (LDG.64/128), all 2^D corner loads of a level
// issued before the blend (memory-level parallelism), the input point loaded once per thread,
// a [B, L*C]-direct variant that removes the permute copy, and vector red.global.add for the
// scatter in the backward pass.
#include "common.cuh"

namespace sanerf {

template <uint32_t C>
struct Row {
    float v[C];
};

template <uint32_t C>
__device__ __forceinline__ Row<C> load_row(const float* __restrict__ p) {
    Row<C> r;
    if constexpr (C == 1) {
        r.v[0] = __ldg(p);
    } else if constexpr (C == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p));
        r.v[0] = t.x; r.v[1] = t.y;
    } else {
#pragma unroll
        for (uint32_t i = 0; i < C; i += 4) {
            float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
            r.v[i] = t.x; r.v[i + 1] = t.y; r.v[i + 2] = t.z; r.v[i + 3] = t.w;
        }
    }
    return r;
}

template <uint32_t D>
struct Cell {
    float frac[D];     // interpolation weight along d (after optional smoothstep)
    float dfrac[D];    // d frac / d pos
    uint32_t base[D];  // lower vertex
};

// Position of a point inside one level (gridencoder.cu:140-160).
template <uint32_t D>
__device__ __forceinline__ Cell<D> locate(const float (&x)[D], uint32_t res, bool align_corners, uint32_t interp) {
    Cell<D> c;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        float pos;
        if (align_corners) {
            pos = x[d] * (float)(res - 1);
            c.base[d] = min((uint32_t)floorf(pos), res - 2);
        } else {
            pos = fminf(fmaxf(__fmaf_rn(x[d], (float)res, -0.5f), 0.0f), (float)(res - 1));
            c.base[d] = (uint32_t)floorf(pos);
        }
        pos -= (float)c.base[d];
        if (interp == 1) {
            c.dfrac[d] = 6 * pos * (1.0f - pos);
            pos = pos * pos * (3.0f - 2.0f * pos);
        } else {
            c.dfrac[d] = 1.0f;
        }
        c.frac[d] = pos;
    }
    return c;
}

template <uint32_t D>
__device__ __forceinline__ bool out_of_unit_cube(const float (&x)[D]) {
    bool oob = false;
#pragma unroll
    for (uint32_t d = 0; d < D; d++) oob |= (x[d] < 0 || x[d] > 1);
    return oob;
}

// 2^D-corner blend of one level (gridencoder.cu:170-195): corner idx, bit d set -> +1 along d;
// weight is the running product in dimension order; accumulation is one FMA per corner/channel.
template <uint32_t D, uint32_t C>
__device__ __forceinline__ void blend_level(const float* __restrict__ level_rows, uint32_t rows, uint32_t res,
                                            uint32_t gridtype, const Cell<D>& c, float (&out)[C]) {
    constexpr uint32_t NC = 1u << D;
    uint32_t row[NC];
    float w[NC];
#pragma unroll
    for (uint32_t idx = 0; idx < NC; idx++) {
        float ww = 1;
        uint32_t p[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) {
                ww *= 1 - c.frac[d];
                p[d] = c.base[d];
            } else {
                ww *= c.frac[d];
                p[d] = min(c.base[d] + 1, res - 1);
            }
        }
        w[idx] = ww;
        row[idx] = vertex_row<D>(gridtype, rows, res, p);
    }
    if constexpr (NC * C <= 64) {
        Row<C> r[NC];
#pragma unroll
        for (uint32_t idx = 0; idx < NC; idx++) r[idx] = load_row<C>(level_rows + (size_t)row[idx] * C);
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
#pragma unroll
        for (uint32_t idx = 0; idx < NC; idx++)
#pragma unroll
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = __fmaf_rn(w[idx], r[idx].v[ch], out[ch]);
    } else {
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) out[ch] = 0;
#pragma unroll 4
        for (uint32_t idx = 0; idx < NC; idx++) {
            Row<C> r = load_row<C>(level_rows + (size_t)row[idx] * C);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ch++) out[ch] = __fmaf_rn(w[idx], r.v[ch], out[ch]);
        }
    }
}

// d out / d x for one level (gridencoder.cu:205-248), layout [D][C].
template <uint32_t D, uint32_t C>
__device__ __forceinline__ void level_dy_dx(const float* __restrict__ level_rows, uint32_t rows, uint32_t res,
                                            uint32_t gridtype, bool align_corners, const Cell<D>& c,
                                            float* __restrict__ dy_dx /*[D*C]*/) {
#pragma unroll
    for (uint32_t gd = 0; gd < D; gd++) {
        float acc[C];
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) acc[ch] = 0;
#pragma unroll
        for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
            float w = (float)(align_corners ? res - 1 : res);
            uint32_t p[D];
#pragma unroll
            for (uint32_t nd = 0; nd < D - 1; nd++) {
                const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                if ((idx & (1u << nd)) == 0) {
                    w *= 1 - c.frac[d];
                    p[d] = c.base[d];
                } else {
                    w *= c.frac[d];
                    p[d] = min(c.base[d] + 1, res - 1);
                }
            }
            p[gd] = c.base[gd];
            Row<C> lo = load_row<C>(level_rows + (size_t)vertex_row<D>(gridtype, rows, res, p) * C);
            p[gd] = min(c.base[gd] + 1, res - 1);
            Row<C> hi = load_row<C>(level_rows + (size_t)vertex_row<D>(gridtype, rows, res, p) * C);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ch++) acc[ch] = __fmaf_rn(w * (hi.v[ch] - lo.v[ch]), c.dfrac[gd], acc[ch]);
        }
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) dy_dx[gd * C + ch] = acc[ch];
    }
}

template <uint32_t C>
__device__ __forceinline__ void store_vec(float* __restrict__ dst, const float (&v)[C]) {
    if constexpr (C == 1) {
        dst[0] = v[0];
    } else if constexpr (C == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (uint32_t i = 0; i < C; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
}

// K1.  One thread per (point, level); blockIdx.y = level so one level's table is hot at a time.
// FUSED=false: reference layout, inputs in [0,1], outputs [L,B,C], optional dy_dx [B,L,D,C].
// FUSED=true : inputs raw in [-bound,bound] (mapped in-kernel), outputs [B, L*C].
template <uint32_t D, uint32_t C, bool FUSED>
__global__ void __launch_bounds__(256) grid_forward_kernel(const float* __restrict__ inputs, float bound,
                                                            const float* __restrict__ embeddings,
                                                            const int32_t* __restrict__ offsets,
                                                            float* __restrict__ outputs, uint32_t B, uint32_t L, float S,
                                                            uint32_t H, float* __restrict__ dy_dx, uint32_t gridtype,
                                                            bool align_corners, uint32_t interp) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        float v = __ldg(inputs + (size_t)b * D + d);
        if constexpr (FUSED) { if (bound > 0.f) v = __fmul_rn(__fadd_rn(v, bound), __frcp_rn(2 * bound)); }  // grid.py:156 (tensor / Python scalar = multiply by the fp32 reciprocal in ATen); bound<=0: already in [0,1]
        x[d] = v;
    }
    float* out = FUSED ? outputs + (size_t)b * L * C + (size_t)level * C : outputs + ((size_t)level * B + b) * C;
    float* dd = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : nullptr;

    float r[C];
    if (out_of_unit_cube<D>(x)) {  // gridencoder.cu:105-130
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++) r[ch] = 0;
        store_vec<C>(out, r);
        if (dd) {
#pragma unroll
            for (uint32_t i = 0; i < D * C; i++) dd[i] = 0;
        }
        return;
    }
    const uint32_t o0 = (uint32_t)__ldg(offsets + level);
    const uint32_t rows = (uint32_t)__ldg(offsets + level + 1) - o0;
    const uint32_t res = level_resolution(level, S, H);
    const float* level_rows = embeddings + (size_t)o0 * C;

    const Cell<D> c = locate<D>(x, res, align_corners, interp);
    blend_level<D, C>(level_rows, rows, res, gridtype, c, r);
    store_vec<C>(out, r);
    if (dd) level_dy_dx<D, C>(level_rows, rows, res, gridtype, align_corners, c, dd);
}

// K2.  Scatter w*grad to the 2^D corners.  One thread per (point, level, channel-group of G).
// Uses vector reductions (red.global.add.v2/v4.f32 on sm_90+) when C allows.
template <uint32_t G>
__device__ __forceinline__ void red_add(float* dst, const float (&v)[G]) {
    if constexpr (G == 1) {
        atomicAdd(dst, v[0]);
    } else if constexpr (G == 2) {
        atomicAdd(reinterpret_cast<float2*>(dst), make_float2(v[0], v[1]));
    } else {
        static_assert(G == 4, "channel group must be 1, 2 or 4");
        atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
    }
}

template <uint32_t D, uint32_t C, uint32_t G, bool FUSED>
__global__ void __launch_bounds__(256) grid_backward_kernel(const float* __restrict__ grad, const float* __restrict__ inputs,
                                                             float bound, const int32_t* __restrict__ offsets,
                                                             float* __restrict__ grad_embeddings, uint32_t B, uint32_t L,
                                                             float S, uint32_t H, uint32_t gridtype, bool align_corners,
                                                             uint32_t interp) {
    constexpr uint32_t NG = C / G;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / NG;
    if (b >= B) return;
    const uint32_t ch0 = (t - b * NG) * G;
    const uint32_t level = blockIdx.y;

    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        float v = __ldg(inputs + (size_t)b * D + d);
        if constexpr (FUSED) { if (bound > 0.f) v = __fmul_rn(__fadd_rn(v, bound), __frcp_rn(2 * bound)); }
        x[d] = v;
    }
    if (out_of_unit_cube<D>(x)) return;  // gridencoder.cu:278-283

    const uint32_t o0 = (uint32_t)__ldg(offsets + level);
    const uint32_t rows = (uint32_t)__ldg(offsets + level + 1) - o0;
    const uint32_t res = level_resolution(level, S, H);
    float* gg = grad_embeddings + (size_t)o0 * C;
    const float* g = FUSED ? grad + (size_t)b * L * C + (size_t)level * C + ch0 : grad + ((size_t)level * B + b) * C + ch0;

    float gv[G];
#pragma unroll
    for (uint32_t i = 0; i < G; i++) gv[i] = __ldg(g + i);
    bool any = false;
#pragma unroll
    for (uint32_t i = 0; i < G; i++) any |= (gv[i] != 0.0f);
    if (!any) return;  // adding +-0 is a no-op; skips the atomics of frozen / unused outputs

    const Cell<D> c = locate<D>(x, res, align_corners, interp);
#pragma unroll
    for (uint32_t idx = 0; idx < (1u << D); idx++) {
        float w = 1;
        uint32_t p[D];
#pragma unroll
        for (uint32_t d = 0; d < D; d++) {
            if ((idx & (1u << d)) == 0) {
                w *= 1 - c.frac[d];
                p[d] = c.base[d];
            } else {
                w *= c.frac[d];
                p[d] = min(c.base[d] + 1, res - 1);
            }
        }
        const uint32_t row = vertex_row<D>(gridtype, rows, res, p);
        float v[G];
#pragma unroll
        for (uint32_t i = 0; i < G; i++) v[i] = w * gv[i];
        red_add<G>(gg + (size_t)row * C + ch0, v);
    }
}

// K3.  grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]   (gridencoder.cu:352-378)
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_input_backward_kernel(const float* __restrict__ grad, const float* __restrict__ dy_dx,
                                                                   float* __restrict__ grad_inputs, uint32_t B, uint32_t L) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* dd = dy_dx + (size_t)b * L * D * C;
    float r = 0;
    for (uint32_t l = 0; l < L; l++) {
#pragma unroll
        for (uint32_t ch = 0; ch < C; ch++)
            r = __fmaf_rn(__ldg(grad + ((size_t)l * B + b) * C + ch), __ldg(dd + l * D * C + d * C + ch), r);
    }
    grad_inputs[t] = r;
}

// K4.  Total-variation regulariser gradient (gridencoder.cu:525-631), bug-compatible: the +1
// neighbour is taken even at the last vertex (`cur_d < resolution` is always true there).
template <uint32_t D, uint32_t C>
__global__ void __launch_bounds__(256) grid_tv_kernel(const float* __restrict__ inputs, const float* __restrict__ embeddings,
                                                       float* __restrict__ grad, const int32_t* __restrict__ offsets, float weight,
                                                       uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                                       bool align_corners) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t level = blockIdx.y;
    float x[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) x[d] = __ldg(inputs + (size_t)b * D + d);
    if (out_of_unit_cube<D>(x)) return;

    const uint32_t o0 = (uint32_t)__ldg(offsets + level);
    const uint32_t rows = (uint32_t)__ldg(offsets + level + 1) - o0;
    const uint32_t res = level_resolution(level, S, H);
    const float* rowsp = embeddings + (size_t)o0 * C;
    float* gr = grad + (size_t)o0 * C;

    uint32_t p[D];
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        if (align_corners) {
            p[d] = min((uint32_t)floorf(x[d] * (float)(res - 1)), res - 2);
        } else {
            p[d] = (uint32_t)floorf(fminf(fmaxf(__fmaf_rn(x[d], (float)res, -0.5f), 0.0f), (float)(res - 1)));
        }
    }
    float sum[C], sq[C];
#pragma unroll
    for (uint32_t ch = 0; ch < C; ch++) sum[ch] = sq[ch] = 0;
    const uint32_t row = vertex_row<D>(gridtype, rows, res, p);
    const Row<C> centre = load_row<C>(rowsp + (size_t)row * C);
    const float w = weight / (2 * D);
#pragma unroll
    for (uint32_t d = 0; d < D; d++) {
        const uint32_t cur = p[d];
        if (cur < res) {
            p[d] = cur + 1;
            const Row<C> nb = load_row<C>(rowsp + (size_t)vertex_row<D>(gridtype, rows, res, p) * C);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ch++) {
                const float gvv = centre.v[ch] - nb.v[ch];
                sum[ch] += gvv;
                sq[ch] = __fmaf_rn(gvv, gvv, sq[ch]);
            }
        }
        if (cur > 0) {
            p[d] = cur - 1;
            const Row<C> nb = load_row<C>(rowsp + (size_t)vertex_row<D>(gridtype, rows, res, p) * C);
#pragma unroll
            for (uint32_t ch = 0; ch < C; ch++) {
                const float gvv = centre.v[ch] - nb.v[ch];
                sum[ch] += gvv;
                sq[ch] = __fmaf_rn(gvv, gvv, sq[ch]);
            }
        }
        p[d] = cur;
    }
#pragma unroll
    for (uint32_t ch = 0; ch < C; ch++) atomicAdd(gr + (size_t)row * C + ch, w * sum[ch] * rsqrtf(sq[ch] + 1e-9f));
}

// K5.  Level-wise mean weight decay (gridencoder.cu:670-703).
__global__ void __launch_bounds__(256) grid_wd_kernel(const float* __restrict__ embeddings, float* __restrict__ grad,
                                                       const int32_t* __restrict__ offsets, float weight, uint32_t B, uint32_t L,
                                                       uint32_t C) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * C) return;
    const uint32_t n = (uint32_t)(i / C);
    uint32_t level = 0, lo = 0, hi = L;
    while (lo < hi) {
        const uint32_t m = (lo + hi) / 2;
        if ((uint32_t)__ldg(offsets + m) <= n) { level = m; lo = m + 1; } else { hi = m; }
    }
    const uint32_t rows = (uint32_t)(__ldg(offsets + level + 1) - __ldg(offsets + level));
    grad[i] += 2 * weight * embeddings[i] / rows;
}

__global__ void level_res_kernel(uint32_t* out, uint32_t L, float S, uint32_t H) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l < L) out[l] = level_resolution(l, S, H);
}

// ---------------------------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------------------------
template <uint32_t D, uint32_t C, bool FUSED>
static int launch_forward(const float* in, float bound, const float* emb, const int32_t* off, float* out, uint32_t B, uint32_t L,
                          uint32_t max_level, float S, uint32_t H, float* dy_dx, uint32_t gridtype, bool ac, uint32_t interp,
                          cudaStream_t st) {
    if (B == 0 || max_level == 0) return 0;
    dim3 grid(div_up(B, 256), max_level);
    grid_forward_kernel<D, C, FUSED><<<grid, 256, 0, st>>>(in, bound, emb, off, out, B, L, S, H, dy_dx, gridtype, ac, interp);
    return check_launch();
}

template <uint32_t D, bool FUSED>
static int dispatch_forward_c(uint32_t C, const float* in, float bound, const float* emb, const int32_t* off, float* out, uint32_t B,
                              uint32_t L, uint32_t max_level, float S, uint32_t H, float* dy_dx, uint32_t gridtype, bool ac,
                              uint32_t interp, cudaStream_t st) {
    switch (C) {
        case 1: return launch_forward<D, 1, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 2: return launch_forward<D, 2, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 4: return launch_forward<D, 4, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 8: return launch_forward<D, 8, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 16: return launch_forward<D, 16, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 32: return launch_forward<D, 32, FUSED>(in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        default: return SANERF_E_CHANNELS;
    }
}

template <bool FUSED>
static int dispatch_forward(uint32_t D, uint32_t C, const float* in, float bound, const float* emb, const int32_t* off, float* out,
                            uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H, float* dy_dx, uint32_t gridtype,
                            bool ac, uint32_t interp, cudaStream_t st) {
    switch (D) {
        case 2: return dispatch_forward_c<2, FUSED>(C, in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 3: return dispatch_forward_c<3, FUSED>(C, in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 4: return dispatch_forward_c<4, FUSED>(C, in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        case 5: return dispatch_forward_c<5, FUSED>(C, in, bound, emb, off, out, B, L, max_level, S, H, dy_dx, gridtype, ac, interp, st);
        default: return SANERF_E_DIM;
    }
}

template <uint32_t D, uint32_t C, bool FUSED>
static int launch_backward(const float* grad, const float* in, float bound, const int32_t* off, float* ge, uint32_t B, uint32_t L,
                           uint32_t max_level, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t interp, cudaStream_t st) {
    if (B == 0 || max_level == 0) return 0;
    constexpr uint32_t G = C >= 4 ? 4 : C;
    dim3 grid(div_up(B * (C / G), 256), max_level);
    grid_backward_kernel<D, C, G, FUSED><<<grid, 256, 0, st>>>(grad, in, bound, off, ge, B, L, S, H, gridtype, ac, interp);
    return check_launch();
}

template <uint32_t D, bool FUSED>
static int dispatch_backward_c(uint32_t C, const float* grad, const float* in, float bound, const int32_t* off, float* ge, uint32_t B,
                               uint32_t L, uint32_t max_level, float S, uint32_t H, uint32_t gridtype, bool ac, uint32_t interp,
                               cudaStream_t st) {
    switch (C) {
        case 1: return launch_backward<D, 1, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 2: return launch_backward<D, 2, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 4: return launch_backward<D, 4, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 8: return launch_backward<D, 8, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 16: return launch_backward<D, 16, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 32: return launch_backward<D, 32, FUSED>(grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        default: return SANERF_E_CHANNELS;
    }
}

template <bool FUSED>
static int dispatch_backward(uint32_t D, uint32_t C, const float* grad, const float* in, float bound, const int32_t* off, float* ge,
                             uint32_t B, uint32_t L, uint32_t max_level, float S, uint32_t H, uint32_t gridtype, bool ac,
                             uint32_t interp, cudaStream_t st) {
    switch (D) {
        case 2: return dispatch_backward_c<2, FUSED>(C, grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 3: return dispatch_backward_c<3, FUSED>(C, grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 4: return dispatch_backward_c<4, FUSED>(C, grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        case 5: return dispatch_backward_c<5, FUSED>(C, grad, in, bound, off, ge, B, L, max_level, S, H, gridtype, ac, interp, st);
        default: return SANERF_E_DIM;
    }
}

template <uint32_t D, uint32_t C>
static int launch_input_backward(const float* grad, const float* dy_dx, float* gi, uint32_t B, uint32_t L, cudaStream_t st) {
    if (B == 0) return 0;
    grid_input_backward_kernel<D, C><<<div_up(B * D, 256), 256, 0, st>>>(grad, dy_dx, gi, B, L);
    return check_launch();
}

template <uint32_t D, uint32_t C>
static int launch_tv(const float* in, const float* emb, float* grad, const int32_t* off, float w, uint32_t B, uint32_t L, float S,
                     uint32_t H, uint32_t gridtype, bool ac, cudaStream_t st) {
    if (B == 0 || L == 0) return 0;
    dim3 grid(div_up(B, 256), L);
    grid_tv_kernel<D, C><<<grid, 256, 0, st>>>(in, emb, grad, off, w, B, L, S, H, gridtype, ac);
    return check_launch();
}

#define SANERF_DISPATCH_DC(D_, C_, CALL)                                                   \
    do {                                                                                   \
        if (C_ != 1 && C_ != 2 && C_ != 4 && C_ != 8 && C_ != 16 && C_ != 32) return SANERF_E_CHANNELS; \
        switch (D_) {                                                                      \
            case 2: switch (C_) { case 1: { constexpr uint32_t DD = 2, CC = 1; CALL; } case 2: { constexpr uint32_t DD = 2, CC = 2; CALL; } case 4: { constexpr uint32_t DD = 2, CC = 4; CALL; } case 8: { constexpr uint32_t DD = 2, CC = 8; CALL; } case 16: { constexpr uint32_t DD = 2, CC = 16; CALL; } default: { constexpr uint32_t DD = 2, CC = 32; CALL; } } \
            case 3: switch (C_) { case 1: { constexpr uint32_t DD = 3, CC = 1; CALL; } case 2: { constexpr uint32_t DD = 3, CC = 2; CALL; } case 4: { constexpr uint32_t DD = 3, CC = 4; CALL; } case 8: { constexpr uint32_t DD = 3, CC = 8; CALL; } case 16: { constexpr uint32_t DD = 3, CC = 16; CALL; } default: { constexpr uint32_t DD = 3, CC = 32; CALL; } } \
            case 4: switch (C_) { case 1: { constexpr uint32_t DD = 4, CC = 1; CALL; } case 2: { constexpr uint32_t DD = 4, CC = 2; CALL; } case 4: { constexpr uint32_t DD = 4, CC = 4; CALL; } case 8: { constexpr uint32_t DD = 4, CC = 8; CALL; } case 16: { constexpr uint32_t DD = 4, CC = 16; CALL; } default: { constexpr uint32_t DD = 4, CC = 32; CALL; } } \
            case 5: switch (C_) { case 1: { constexpr uint32_t DD = 5, CC = 1; CALL; } case 2: { constexpr uint32_t DD = 5, CC = 2; CALL; } case 4: { constexpr uint32_t DD = 5, CC = 4; CALL; } case 8: { constexpr uint32_t DD = 5, CC = 8; CALL; } case 16: { constexpr uint32_t DD = 5, CC = 16; CALL; } default: { constexpr uint32_t DD = 5, CC = 32; CALL; } } \
            default: return SANERF_E_DIM;                                                  \
        }                                                                                  \
    } while (0)

}  // namespace sanerf

using namespace sanerf;

extern "C" {

int sanerf_grid_encode_forward(const float* inputs, const float* embeddings, const int32_t* offsets, float* outputs, uint32_t B,
                               uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S, uint32_t H, float* dy_dx,
                               uint32_t gridtype, int align_corners, uint32_t interp, sanerf_stream_t stream) {
    if (B && (!inputs || !embeddings || !offsets || !outputs)) return SANERF_E_NULL;
    return dispatch_forward<false>(D, C, inputs, 0.f, embeddings, offsets, outputs, B, L, max_level, S, H, dy_dx, gridtype,
                                   align_corners != 0, interp, (cudaStream_t)stream);
}

int sanerf_grid_encode_forward_fused(const float* positions, float bound, const float* embeddings, const int32_t* offsets,
                                     float* outputs_BLC, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                                     float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                     sanerf_stream_t stream) {
    if (B && (!positions || !embeddings || !offsets || !outputs_BLC)) return SANERF_E_NULL;
    return dispatch_forward<true>(D, C, positions, bound, embeddings, offsets, outputs_BLC, B, L, max_level, S, H, nullptr, gridtype,
                                  align_corners != 0, interp, (cudaStream_t)stream);
}

int sanerf_grid_encode_backward(const float* grad, const float* inputs, const float* embeddings, const int32_t* offsets,
                                float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level, float S,
                                uint32_t H, const float* dy_dx, float* grad_inputs, uint32_t gridtype, int align_corners,
                                uint32_t interp, sanerf_stream_t stream) {
    (void)embeddings;
    if (B && (!grad || !inputs || !offsets || !grad_embeddings)) return SANERF_E_NULL;
    int rc = dispatch_backward<false>(D, C, grad, inputs, 0.f, offsets, grad_embeddings, B, L, max_level, S, H, gridtype,
                                      align_corners != 0, interp, (cudaStream_t)stream);
    if (rc != 0 || !dy_dx) return rc;
    if (!grad_inputs) return SANERF_E_NULL;
    SANERF_DISPATCH_DC(D, C, return (launch_input_backward<DD, CC>(grad, dy_dx, grad_inputs, B, L, (cudaStream_t)stream)));
}

int sanerf_grid_encode_backward_fused(const float* grad_BLC, const float* positions, float bound, const int32_t* offsets,
                                      float* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, uint32_t max_level,
                                      float S, uint32_t H, uint32_t gridtype, int align_corners, uint32_t interp,
                                      sanerf_stream_t stream) {
    if (B && (!grad_BLC || !positions || !offsets || !grad_embeddings)) return SANERF_E_NULL;
    return dispatch_backward<true>(D, C, grad_BLC, positions, bound, offsets, grad_embeddings, B, L, max_level, S, H, gridtype,
                                   align_corners != 0, interp, (cudaStream_t)stream);
}

int sanerf_grad_total_variation(const float* inputs, const float* embeddings, float* grad, const int32_t* offsets, float weight,
                                uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                int align_corners, sanerf_stream_t stream) {
    if (B && (!inputs || !embeddings || !grad || !offsets)) return SANERF_E_NULL;
    SANERF_DISPATCH_DC(D, C, return (launch_tv<DD, CC>(inputs, embeddings, grad, offsets, weight, B, L, S, H, gridtype,
                                                        align_corners != 0, (cudaStream_t)stream)));
}

int sanerf_grad_weight_decay(const float* embeddings, float* grad, const int32_t* offsets, float weight, uint32_t B, uint32_t C,
                             uint32_t L, sanerf_stream_t stream) {
    if (B && (!embeddings || !grad || !offsets)) return SANERF_E_NULL;
    if (B == 0) return 0;
    const size_t n = (size_t)B * C;
    grid_wd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(embeddings, grad, offsets, weight, B, L, C);
    return check_launch();
}

int sanerf_grid_level_resolutions(uint32_t* out_res, uint32_t L, float S, uint32_t H, sanerf_stream_t stream) {
    if (!out_res) return SANERF_E_NULL;
    if (L == 0) return 0;
    level_res_kernel<<<div_up(L, 32), 32, 0, (cudaStream_t)stream>>>(out_res, L, S, H);
    return check_launch();
}

}  // extern "C"
