/*
 * sanerf_oracle.c -- CPU restatement of the reference's CUDA-only encoder kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or
 * executed by the product path (sanerf_hq_b200/, gridencoder/, shencoder/,
 * freqencoder/, nerf/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker or
 * as the CPU arm being timed.
 *
 * The reference ships no CPU implementation of these kernels;
 * each function below restates one reference kernel in plain C, thread loop
 * replaced by a plain for loop, citing the lines it follows.
 *
 * Parity pinning: the reference holds no golden vectors for this path
 * (SURVEY.md section 4).  This file is pinned two ways:
 *   (1) tests/golden/ fixtures generated in the build container by importing
 *       the reference's own Python (nerf/renderer.py, nerf/network.py) on top
 *       of these kernels (tests/golden/make_golden.py), and
 *   (2) on the GPU box against the reference's own CUDA kernels compiled
 *       verbatim into oracle/_ref (tests/test_ref_cuda_gpu.py).
 *
 * FMA contraction: nvcc (default -fmad=true) contracts a*b+c in the reference
 * kernels; the places where that happens are written with fmaf() here so the
 * rounding matches.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* Threading: no OpenMP (a second OpenMP runtime next to torch's is fragile).  The hot
 * forward kernels take a point range [b0,b1) so the Python wrapper (oracle/kernels.py)
 * can run disjoint ranges on a thread pool; ctypes drops the GIL during the call. */

#define MAX_D 5
#define MAX_C 32

/* gridencoder/src/gridencoder.cu:45-58  fast_hash: uint32 wrap-around multiply, XOR fold */
static uint32_t fast_hash(const uint32_t *pos_grid, uint32_t D)
{
    static const uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u,
                                       2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
    for (uint32_t i = 0; i < D; ++i) r ^= pos_grid[i] * primes[i];
    return r;
}

/* gridencoder/src/gridencoder.cu:61-79  get_grid_index.  NB the dense stride
 * loop stops as soon as stride > hashmap_size (condition is part of the loop
 * test), and the result is always reduced modulo hashmap_size. */
static uint32_t grid_index(uint32_t gridtype, uint32_t ch, uint32_t hashmap_size,
                           uint32_t resolution, const uint32_t *pos_grid, uint32_t D, uint32_t C)
{
    uint32_t stride = 1, index = 0;
    for (uint32_t d = 0; d < D && stride <= hashmap_size; d++) {
        index += pos_grid[d] * stride;
        stride *= resolution;
    }
    if (gridtype == 0 && stride > hashmap_size) index = fast_hash(pos_grid, D);
    return (index % hashmap_size) * C + ch;
}

/* gridencoder/src/gridencoder.cu:133 (and :275, :551): kernel-side level resolution, fp32 */
uint32_t oracle_level_resolution(uint32_t level, float S, uint32_t H)
{
    return (uint32_t)ceilf(exp2f((float)level * S) * (float)H);
}

static float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
static float smoothstep_d(float v) { return 6 * v * (1.0f - v); }

/* position within a level: gridencoder.cu:140-160 */
static void locate(const float *x, uint32_t D, uint32_t resolution, int align_corners, uint32_t interp,
                   float *pos, float *pos_deriv, uint32_t *pos_grid)
{
    for (uint32_t d = 0; d < D; d++) {
        if (align_corners) {
            pos[d] = x[d] * (float)(resolution - 1);
            uint32_t f = (uint32_t)floorf(pos[d]);
            pos_grid[d] = f < resolution - 2 ? f : resolution - 2;
        } else {
            /* nvcc contracts x*res - 0.5f into one FMA */
            pos[d] = fminf(fmaxf(fmaf(x[d], (float)resolution, -0.5f), 0.0f), (float)(resolution - 1));
            pos_grid[d] = (uint32_t)floorf(pos[d]);
        }
        pos[d] -= (float)pos_grid[d];
        if (interp == 1) {
            pos_deriv[d] = smoothstep_d(pos[d]);
            pos[d] = smoothstep_f(pos[d]);
        } else {
            pos_deriv[d] = 1.0f;
        }
    }
}

/*
 * K1  kernel_grid  gridencoder/src/gridencoder.cu:82-249
 * inputs [B,D] in [0,1]; embeddings [sO,C]; offsets [L+1]; outputs [L,B,C];
 * dy_dx [B, L*D*C] or NULL.  Levels >= max_level are left untouched (the
 * Python caller zero-fills them, gridencoder/grid.py:50-51).
 */
void oracle_grid_encode_forward(const float *inputs, const float *embeddings, const int32_t *offsets,
                                float *outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                uint32_t max_level, float S, uint32_t H, float *dy_dx,
                                uint32_t gridtype, int align_corners, uint32_t interp,
                                uint32_t b0, uint32_t b1)
{
    for (uint32_t level = 0; level < max_level; level++) {
        const float *grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const uint32_t resolution = oracle_level_resolution(level, S, H);
        for (int64_t b = b0; b < (int64_t)b1; b++) {
            const float *x = inputs + (size_t)b * D;
            float *out = outputs + ((size_t)level * B + b) * C;
            float *dd = dy_dx ? dy_dx + (size_t)b * D * L * C + (size_t)level * D * C : 0;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (x[d] < 0 || x[d] > 1) oob = 1;          /* :105-111; NaN compares false -> not oob */
            if (oob) {
                for (uint32_t c = 0; c < C; c++) out[c] = 0;
                if (dd) for (uint32_t i = 0; i < D * C; i++) dd[i] = 0;
                continue;
            }
            float pos[MAX_D], pos_deriv[MAX_D];
            uint32_t pos_grid[MAX_D], loc[MAX_D];
            locate(x, D, resolution, align_corners, interp, pos, pos_deriv, pos_grid);

            float res[MAX_C];
            for (uint32_t c = 0; c < C; c++) res[c] = 0;
            for (uint32_t idx = 0; idx < (1u << D); idx++) {  /* :170-195 corner order, bit d -> +1 along d */
                float w = 1;
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) {
                        w *= 1 - pos[d];
                        loc[d] = pos_grid[d];
                    } else {
                        w *= pos[d];
                        loc[d] = pos_grid[d] + 1 < resolution - 1 ? pos_grid[d] + 1 : resolution - 1;
                    }
                }
                uint32_t index = grid_index(gridtype, 0, hashmap_size, resolution, loc, D, C);
                for (uint32_t c = 0; c < C; c++) res[c] = fmaf(w, grid[index + c], res[c]);
            }
            for (uint32_t c = 0; c < C; c++) out[c] = res[c];

            if (dd) {                                         /* :205-248 */
                for (uint32_t gd = 0; gd < D; gd++) {
                    float rg[MAX_C];
                    for (uint32_t c = 0; c < C; c++) rg[c] = 0;
                    for (uint32_t idx = 0; idx < (1u << (D - 1)); idx++) {
                        float w = (float)(align_corners ? resolution - 1 : resolution);
                        for (uint32_t nd = 0; nd < D - 1; nd++) {
                            const uint32_t d = (nd >= gd) ? nd + 1 : nd;
                            if ((idx & (1u << nd)) == 0) {
                                w *= 1 - pos[d];
                                loc[d] = pos_grid[d];
                            } else {
                                w *= pos[d];
                                loc[d] = pos_grid[d] + 1 < resolution - 1 ? pos_grid[d] + 1 : resolution - 1;
                            }
                        }
                        loc[gd] = pos_grid[gd];
                        uint32_t il = grid_index(gridtype, 0, hashmap_size, resolution, loc, D, C);
                        loc[gd] = pos_grid[gd] + 1 < resolution - 1 ? pos_grid[gd] + 1 : resolution - 1;
                        uint32_t ir = grid_index(gridtype, 0, hashmap_size, resolution, loc, D, C);
                        for (uint32_t c = 0; c < C; c++)
                            rg[c] = fmaf(w * (grid[ir + c] - grid[il + c]), pos_deriv[gd], rg[c]);
                    }
                    for (uint32_t c = 0; c < C; c++) dd[gd * C + c] = rg[c];
                }
            }
        }
    }
}

/*
 * K2  kernel_grid_backward  gridencoder/src/gridencoder.cu:252-349
 * grad [L,B,C]; grad_embeddings [sO,C] (caller zero-fills, grid.py:83).
 * The reference scatters with atomicAdd in nondeterministic order; this
 * restatement accumulates in point order (serial), so parity is to fp32
 * summation-order tolerance, not bit-exact.
 */
void oracle_grid_encode_backward(const float *grad, const float *inputs, const int32_t *offsets,
                                 float *grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L,
                                 uint32_t max_level, float S, uint32_t H, uint32_t gridtype,
                                 int align_corners, uint32_t interp)
{
    (void)L;
    for (int64_t level = 0; level < (int64_t)max_level; level++) {
        float *gg = grad_embeddings + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const uint32_t resolution = oracle_level_resolution((uint32_t)level, S, H);
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            const float *g = grad + ((size_t)level * B + b) * C;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            float pos[MAX_D], pos_deriv[MAX_D];
            uint32_t pos_grid[MAX_D], loc[MAX_D];
            locate(x, D, resolution, align_corners, interp, pos, pos_deriv, pos_grid);
            for (uint32_t idx = 0; idx < (1u << D); idx++) {
                float w = 1;
                for (uint32_t d = 0; d < D; d++) {
                    if ((idx & (1u << d)) == 0) {
                        w *= 1 - pos[d];
                        loc[d] = pos_grid[d];
                    } else {
                        w *= pos[d];
                        loc[d] = pos_grid[d] + 1 < resolution - 1 ? pos_grid[d] + 1 : resolution - 1;
                    }
                }
                uint32_t index = grid_index(gridtype, 0, hashmap_size, resolution, loc, D, C);
                for (uint32_t c = 0; c < C; c++) gg[index + c] += w * g[c];
            }
        }
    }
}

/* K3  kernel_input_backward  gridencoder.cu:352-378 : grad_inputs[b,d] = sum_{l,c} grad[l,b,c]*dy_dx[b,l,d,c] */
void oracle_grid_input_backward(const float *grad, const float *dy_dx, float *grad_inputs,
                                uint32_t B, uint32_t D, uint32_t C, uint32_t L)
{
    for (int64_t t = 0; t < (int64_t)B * D; t++) {
        const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (int64_t)b * D);
        const float *dd = dy_dx + (size_t)b * L * D * C;
        float r = 0;
        for (uint32_t l = 0; l < L; l++)
            for (uint32_t c = 0; c < C; c++)
                r = fmaf(grad[((size_t)l * B + b) * C + c], dd[l * D * C + d * C + c], r);
        grad_inputs[t] = r;
    }
}

/*
 * K4  kernel_grad_tv  gridencoder.cu:525-631.  In-place on grad.  Bug-compatible:
 * `cur_d < resolution` is always true, so the +1 neighbour may index `resolution`
 * (wrapped by the modulo / hash).  Serial accumulation order (reference: atomics).
 */
void oracle_grad_total_variation(const float *inputs, const float *embeddings, float *grad,
                                 const int32_t *offsets, float weight, uint32_t B, uint32_t D,
                                 uint32_t C, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                                 int align_corners)
{
    for (uint32_t level = 0; level < L; level++) {
        const float *grid = embeddings + (size_t)(uint32_t)offsets[level] * C;
        float *gr = grad + (size_t)(uint32_t)offsets[level] * C;
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        const uint32_t resolution = oracle_level_resolution(level, S, H);
        for (uint32_t b = 0; b < B; b++) {
            const float *x = inputs + (size_t)b * D;
            int oob = 0;
            for (uint32_t d = 0; d < D; d++)
                if (x[d] < 0 || x[d] > 1) oob = 1;
            if (oob) continue;
            uint32_t pos_grid[MAX_D];
            for (uint32_t d = 0; d < D; d++) {
                float p;
                if (align_corners) {
                    p = x[d] * (float)(resolution - 1);
                    uint32_t f = (uint32_t)floorf(p);
                    pos_grid[d] = f < resolution - 2 ? f : resolution - 2;
                } else {
                    p = fminf(fmaxf(fmaf(x[d], (float)resolution, -0.5f), 0.0f), (float)(resolution - 1));
                    pos_grid[d] = (uint32_t)floorf(p);
                }
            }
            float results[MAX_C], idelta[MAX_C];
            for (uint32_t c = 0; c < C; c++) results[c] = idelta[c] = 0;
            uint32_t index = grid_index(gridtype, 0, hashmap_size, resolution, pos_grid, D, C);
            float w = weight / (2 * D);
            for (uint32_t d = 0; d < D; d++) {
                uint32_t cur = pos_grid[d];
                if (cur < resolution) {
                    pos_grid[d] = cur + 1;
                    uint32_t ir = grid_index(gridtype, 0, hashmap_size, resolution, pos_grid, D, C);
                    for (uint32_t c = 0; c < C; c++) {
                        float gv = grid[index + c] - grid[ir + c];
                        results[c] += gv;
                        idelta[c] = fmaf(gv, gv, idelta[c]);
                    }
                }
                if (cur > 0) {
                    pos_grid[d] = cur - 1;
                    uint32_t il = grid_index(gridtype, 0, hashmap_size, resolution, pos_grid, D, C);
                    for (uint32_t c = 0; c < C; c++) {
                        float gv = grid[index + c] - grid[il + c];
                        results[c] += gv;
                        idelta[c] = fmaf(gv, gv, idelta[c]);
                    }
                }
                pos_grid[d] = cur;
            }
            for (uint32_t c = 0; c < C; c++)
                gr[index + c] += w * results[c] * (1.0f / sqrtf(idelta[c] + 1e-9f));
        }
    }
}

/* K5  kernel_grad_wd  gridencoder.cu:670-703 : grad += 2*weight*param / rows(level) */
void oracle_grad_weight_decay(const float *embeddings, float *grad, const int32_t *offsets,
                              float weight, uint32_t B, uint32_t C, uint32_t L)
{
    for (int64_t b = 0; b < (int64_t)B * C; b++) {
        const uint32_t n = (uint32_t)(b / C);
        uint32_t level = 0, l = 0, r = L;
        while (l < r) {
            uint32_t m = (l + r) / 2;
            if ((uint32_t)offsets[m] <= n) { level = m; l = m + 1; } else { r = m; }
        }
        const uint32_t hashmap_size = (uint32_t)(offsets[level + 1] - offsets[level]);
        grad[b] += 2 * weight * embeddings[b] / hashmap_size;
    }
}

/*
 * K6  kernel_sh  shencoder/src/shencoder.cu:27-123 (values) for degree 1..4 with the
 * reference's fp32 constants and expression shapes (:49-68).  Degrees 5..8 are
 * evaluated in double from the same closed forms (:69-121) and rounded once, so
 * they are an accurate (<=1 ulp-ish) restatement rather than a rounding-faithful one.
 * inputs [B,3] (already normalised by the Python caller, sphere_harmonics.py:79-82).
 */
static void sh_high(double x, double y, double z, uint32_t C, float *o)
{
    double xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    double x4 = x2 * x2, y4 = y2 * y2, z4 = z2 * z2, x6 = x4 * x2, y6 = y4 * y2, z6 = z4 * z2;
    const double PI_ = 3.14159265358979323846;
    const double isp = 1.0 / sqrt(PI_);
    if (C <= 4) return;
    o[16] = (float)(0.75 * sqrt(35.) * isp * xy * (x2 - y2));
    o[17] = (float)(0.375 * sqrt(70.) * isp * yz * (-3 * x2 + y2));
    o[18] = (float)(0.75 * sqrt(5.) * isp * xy * (7 * z2 - 1));
    o[19] = (float)(0.375 * sqrt(10.) * isp * yz * (3 - 7 * z2));
    o[20] = (float)(3. / 16 * isp * (-30 * z2 + 35 * z4 + 3));
    o[21] = (float)(0.375 * sqrt(10.) * isp * xz * (3 - 7 * z2));
    o[22] = (float)(0.375 * sqrt(5.) * isp * (x2 - y2) * (7 * z2 - 1));
    o[23] = (float)(0.375 * sqrt(70.) * isp * xz * (-x2 + 3 * y2));
    o[24] = (float)(3. / 16 * sqrt(35.) * isp * (-6 * x2 * y2 + x4 + y4));
    if (C <= 5) return;
    o[25] = (float)(3. / 32 * sqrt(154.) * isp * y * (10 * x2 * y2 - 5 * x4 - y4));
    o[26] = (float)(0.75 * sqrt(385.) * isp * xy * z * (x2 - y2));
    o[27] = (float)(-1. / 32 * sqrt(770.) * isp * y * (3 * x2 - y2) * (9 * z2 - 1));
    o[28] = (float)(0.25 * sqrt(1155.) * isp * xy * z * (3 * z2 - 1));
    o[29] = (float)(1. / 16 * sqrt(165.) * isp * y * (14 * z2 - 21 * z4 - 1));
    o[30] = (float)(1. / 16 * sqrt(11.) * isp * z * (-70 * z2 + 63 * z4 + 15));
    o[31] = (float)(1. / 16 * sqrt(165.) * isp * x * (14 * z2 - 21 * z4 - 1));
    o[32] = (float)(0.125 * sqrt(1155.) * isp * z * (x2 - y2) * (3 * z2 - 1));
    o[33] = (float)(-1. / 32 * sqrt(770.) * isp * x * (x2 - 3 * y2) * (9 * z2 - 1));
    o[34] = (float)(3. / 16 * sqrt(385.) * isp * z * (-6 * x2 * y2 + x4 + y4));
    o[35] = (float)(3. / 32 * sqrt(154.) * isp * x * (10 * x2 * y2 - x4 - 5 * y4));
    if (C <= 6) return;
    o[36] = (float)(1. / 32 * sqrt(6006.) * isp * xy * (-10 * x2 * y2 + 3 * x4 + 3 * y4));
    o[37] = (float)(3. / 32 * sqrt(2002.) * isp * yz * (10 * x2 * y2 - 5 * x4 - y4));
    o[38] = (float)(0.375 * sqrt(91.) * isp * xy * (x2 - y2) * (11 * z2 - 1));
    o[39] = (float)(-1. / 32 * sqrt(2730.) * isp * yz * (3 * x2 - y2) * (11 * z2 - 3));
    o[40] = (float)(1. / 32 * sqrt(2730.) * isp * xy * (-18 * z2 + 33 * z4 + 1));
    o[41] = (float)(1. / 16 * sqrt(273.) * isp * yz * (30 * z2 - 33 * z4 - 5));
    o[42] = (float)(1. / 32 * sqrt(13.) * isp * (105 * z2 - 315 * z4 + 231 * z6 - 5));
    o[43] = (float)(1. / 16 * sqrt(273.) * isp * xz * (30 * z2 - 33 * z4 - 5));
    o[44] = (float)(1. / 64 * sqrt(2730.) * isp * (x2 - y2) * (11 * z2 * (3 * z2 - 1) - 7 * z2 + 1));
    o[45] = (float)(-1. / 32 * sqrt(2730.) * isp * xz * (x2 - 3 * y2) * (11 * z2 - 3));
    o[46] = (float)(3. / 32 * sqrt(91.) * isp * (11 * z2 - 1) * (-6 * x2 * y2 + x4 + y4));
    o[47] = (float)(3. / 32 * sqrt(2002.) * isp * xz * (10 * x2 * y2 - x4 - 5 * y4));
    o[48] = (float)(1. / 64 * sqrt(6006.) * isp * (15 * x2 * y4 - 15 * x4 * y2 + x6 - y6));
    if (C <= 7) return;
    o[49] = (float)(3. / 64 * sqrt(715.) * isp * y * (-21 * x2 * y4 + 35 * x4 * y2 - 7 * x6 + y6));
    o[50] = (float)(3. / 32 * sqrt(10010.) * isp * xy * z * (-10 * x2 * y2 + 3 * x4 + 3 * y4));
    o[51] = (float)(-3. / 64 * sqrt(385.) * isp * y * (13 * z2 - 1) * (-10 * x2 * y2 + 5 * x4 + y4));
    o[52] = (float)(0.375 * sqrt(385.) * isp * xy * z * (x2 - y2) * (13 * z2 - 3));
    o[53] = (float)(-3. / 64 * sqrt(35.) * isp * y * (3 * x2 - y2) * (13 * z2 * (11 * z2 - 3) - 27 * z2 + 3));
    o[54] = (float)(3. / 32 * sqrt(70.) * isp * xy * z * (-110 * z2 + 143 * z4 + 15));
    o[55] = (float)(1. / 64 * sqrt(105.) * isp * y * (-135 * z2 + 495 * z4 - 429 * z6 + 5));
    o[56] = (float)(1. / 32 * sqrt(15.) * isp * z * (315 * z2 - 693 * z4 + 429 * z6 - 35));
    o[57] = (float)(1. / 64 * sqrt(105.) * isp * x * (-135 * z2 + 495 * z4 - 429 * z6 + 5));
    o[58] = (float)(1. / 64 * sqrt(70.) * isp * z * (x2 - y2) * (143 * z2 * (3 * z2 - 1) - 187 * z2 + 45));
    o[59] = (float)(-3. / 64 * sqrt(35.) * isp * x * (x2 - 3 * y2) * (13 * z2 * (11 * z2 - 3) - 27 * z2 + 3));
    o[60] = (float)(3. / 32 * sqrt(385.) * isp * z * (13 * z2 - 3) * (-6 * x2 * y2 + x4 + y4));
    o[61] = (float)(-3. / 64 * sqrt(385.) * isp * x * (13 * z2 - 1) * (-10 * x2 * y2 + x4 + 5 * y4));
    o[62] = (float)(3. / 64 * sqrt(10010.) * isp * z * (15 * x2 * y4 - 15 * x4 * y2 + x6 - y6));
    o[63] = (float)(3. / 64 * sqrt(715.) * isp * x * (-35 * x2 * y4 + 21 * x4 * y2 - x6 + 7 * y6));
}

void oracle_sh_encode_forward(const float *inputs, float *outputs, uint32_t B, uint32_t C,
                              uint32_t b0, uint32_t b1)
{
    const uint32_t C2 = C * C;
    (void)B;
    for (int64_t b = b0; b < (int64_t)b1; b++) {
        const float x = inputs[b * 3 + 0], y = inputs[b * 3 + 1], z = inputs[b * 3 + 2];
        float *o = outputs + (size_t)b * C2;
        const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
        o[0] = 0.28209479177387814f;
        if (C <= 1) continue;
        o[1] = -0.48860251190291987f * y;
        o[2] = 0.48860251190291987f * z;
        o[3] = -0.48860251190291987f * x;
        if (C <= 2) continue;
        o[4] = 1.0925484305920792f * xy;
        o[5] = -1.0925484305920792f * yz;
        o[6] = fmaf(0.94617469575755997f, z2, -0.31539156525251999f);
        o[7] = -1.0925484305920792f * xz;
        o[8] = fmaf(0.54627421529603959f, x2, -(0.54627421529603959f * y2));
        if (C <= 3) continue;
        o[9] = 0.59004358992664352f * y * fmaf(-3.0f, x2, y2);
        o[10] = 2.8906114426405538f * xy * z;
        o[11] = 0.45704579946446572f * y * fmaf(-5.0f, z2, 1.0f);
        o[12] = 0.3731763325901154f * z * fmaf(5.0f, z2, -3.0f);
        o[13] = 0.45704579946446572f * x * fmaf(-5.0f, z2, 1.0f);
        o[14] = 1.4453057213202769f * z * (x2 - y2);
        o[15] = 0.59004358992664352f * x * fmaf(3.0f, y2, -x2);
        sh_high(x, y, z, C, o);
    }
}

/* K8  kernel_freq  freqencoder/src/freqencoder.cu:30-58.  The device uses the fast-math
 * __sinf(scalbnf(x,f) + phase); here sinf -- parity to ~1e-6 abs for small |2^f x| only. */
void oracle_freq_encode_forward(const float *inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                float *outputs)
{
    (void)deg;
    const float half_pi = 3.141592653589793f / 2;
    for (int64_t t = 0; t < (int64_t)B * C; t++) {
        const uint32_t b = (uint32_t)(t / C), c = (uint32_t)(t - (int64_t)b * C);
        const float *x = inputs + (size_t)b * D;
        if (c < D) {
            outputs[t] = x[c];
        } else {
            const uint32_t col = c / D - 1, d = c % D, freq = col / 2;
            const float phase = (float)(col % 2) * half_pi;
            outputs[t] = sinf(scalbnf(x[d], (int)freq) + phase);
        }
    }
}

/* K9  kernel_freq_backward  freqencoder.cu:63-94 */
void oracle_freq_encode_backward(const float *grad, const float *outputs, uint32_t B, uint32_t D,
                                 uint32_t deg, uint32_t C, float *grad_inputs)
{
    for (int64_t t = 0; t < (int64_t)B * D; t++) {
        const uint32_t b = (uint32_t)(t / D), d = (uint32_t)(t - (int64_t)b * D);
        const float *g = grad + (size_t)b * C, *o = outputs + (size_t)b * C;
        float r = g[d];
        g += D; o += D;
        for (uint32_t f = 0; f < deg; f++) {
            r += scalbnf(1.0f, (int)f) * (g[d] * o[D + d] - g[D + d] * o[d]);
            g += 2 * D; o += 2 * D;
        }
        grad_inputs[t] = r;
    }
}

