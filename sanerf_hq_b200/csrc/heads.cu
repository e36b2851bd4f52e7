// heads.cu -- the object (instance-mask) head of the render path on the 5th-gen tensor cores (sm_100a).
//
// Reference: nerf/renderer.py:376-385 + nerf/network.py:119-123, 31-66 (SkipConnMLP, no skip, no bias, leaky_relu 0.01):
//     point_masks = mask_mlp(cat[m_grid(x) (128), geo_feat (15)])         per SAMPLE, 143 -> 256 -> 256 -> n_inst
//     instance_mask_logits = sum_samples weights * point_masks             per ray
// 6.57 MFLOP per ray, 4.2 TFLOP per 800x800 frame: the one dense contraction of the path.
//
// One CTA owns the tensor memory of its SM and walks over tiles of 128 samples (= 4 rays); 8 warps run the epilogues, 16
// producer warps gather the feature grid for the next tile, one thread issues the MMAs (mask_head_kernel below):
//   * activations live in TMEM as the A operand (bf16 hi | bf16 lo, two K values per 32-bit column), one row per TMEM lane;
//     warps w and w+4 share the 32 lanes of sub-partition w and split the columns between them;
//   * weights are pre-split into bf16 hi / lo operand images (K-major, no swizzle) by a prepare kernel and streamed from L2
//     through a four-stage ring of 32 KB K-chunks filled by TMA bulk copies (cp.async.bulk + mbarrier expect_tx; the whole
//     MLP is 410 KB of operands -- it does not fit in shared memory);
//   * D[128,256] fp32 accumulates in one 256-column TMEM region while A is read from the other; a layer is issued as two
//     128-column halves, and the epilogue of a half (leaky_relu, bf16 hi/lo split) rewrites it IN PLACE as half of the next
//     layer's A operand while the tensor cores work on the other half -- the two regions swap roles every layer;
//   * split precision: D = Ah*Wh + Ah*Wl + Al*Wh (the dropped Al*Wl is 2^-18 relative), fp32 accumulation;
//   * the last layer (N padded to 16) is composited with the sample weights by a warp reduction (one warp = one ray).
//
// Input layout: the render kernel writes one 18-float record (point in [0,1]^3, geo_feat) per sample "tile-transposed",
// [tile][k][128 rows], so that both its stores and the loads here are coalesced (row r = ray*32 + sample; tile = r / 128); the
// 143-wide MLP input [m_grid(x) (128) | geo_feat (15)] is assembled in shared memory by the producer warps.
#include <cuda_bf16.h>

#include <cstdint>

#include "common.cuh"
#include "grid_dev.cuh"
#include "tc.cuh"

namespace sanerf {

constexpr int kHeadThreads = 256;
constexpr int kMaskK0 = 143, kMaskK0P = 144, kMaskH = 256, kMaskNOut = 16;
constexpr int kStageBytes = 2 * kMaskH * 64 * 2;   // SAM head ring stage: a [256 x 64] chunk, bf16 hi + lo images = 65536 B
// object head: every layer is issued as two output halves of 128 columns; K chunks of 48 (layer 0) / 64 (layer 1)
constexpr int kHalfN = 128, kKc0 = 48, kKc0N = 3, kKc1 = 64, kKc1N = 4;
constexpr int kMaskChunks0 = 2 * kKc0N, kMaskChunks = kMaskChunks0 + 2 * kKc1N;   // 6 + 8 chunks per tile
constexpr int kImgM0 = 2 * kHalfN * kKc0, kImgM1 = 2 * kHalfN * kKc1;             // bf16 elements (hi + lo image) per chunk
constexpr int kImg2 = 2 * kMaskNOut * kMaskH;                                     // resident layer-2 image
constexpr int kOffM1 = kMaskChunks0 * kImgM0, kOffM2 = kOffM1 + 2 * kKc1N * kImgM1, kImgTotal = kOffM2 + kImg2;
static_assert(kImgTotal == 2 * (kMaskH * kMaskK0P + kMaskH * kMaskH + kMaskNOut * kMaskH), "hi + lo image of every weight");
constexpr int kMaskStages = 4, kMaskStageBytes = kImgM1 * 2;                      // 4 x 32 KB
constexpr uint32_t kA0Col = 16;   // first TMEM column of the layer-0 input inside its region (the 16 before it hold the last layer's D)
constexpr uint32_t kColsAlo = 128, kColsD = 256;

using tc::idesc_bf16;
using tc::mma_bf16_ts;
using tc::pack_split16;
using tc::split_bf16;
__host__ __device__ constexpr int img_index(int n, int k, int N) { return tc::bf16_img_index(n, k, N); }

// ---- prepare: nn.Linear weights -> chunked operand images -------------------------------------------------------
// Stream order of one tile (kMaskChunks chunks, each an [128 x KC] K-major image, bf16 hi followed by bf16 lo):
//   layer 0: output half h = 0,1 x three 48-wide K chunks;  layer 1: output half h = 0,1 x four 64-wide K chunks;
//   then the resident layer-2 image [16 x 256].
__global__ void mask_prepare_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2, uint32_t n_inst,
                                    __nv_bfloat16* __restrict__ img) {
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kMaskH * kMaskK0P; i += stride) {   // layer 0 [256,143(+1)]
        const int n = i / kMaskK0P, k = i % kMaskK0P, c = (n / kHalfN) * kKc0N + k / kKc0, kk = k % kKc0, nn = n % kHalfN;
        __nv_bfloat16 h, l;
        split_bf16(k < kMaskK0 ? w0[n * kMaskK0 + k] : 0.f, h, l);
        img[c * kImgM0 + img_index(nn, kk, kHalfN)] = h;
        img[c * kImgM0 + kHalfN * kKc0 + img_index(nn, kk, kHalfN)] = l;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kMaskH * kMaskH; i += stride) {     // layer 1 [256,256]
        const int n = i / kMaskH, k = i % kMaskH, c = (n / kHalfN) * kKc1N + k / kKc1, kk = k % kKc1, nn = n % kHalfN;
        __nv_bfloat16 h, l;
        split_bf16(w1[n * kMaskH + k], h, l);
        img[kOffM1 + c * kImgM1 + img_index(nn, kk, kHalfN)] = h;
        img[kOffM1 + c * kImgM1 + kHalfN * kKc1 + img_index(nn, kk, kHalfN)] = l;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kMaskNOut * kMaskH; i += stride) {  // layer 2 [n_inst,256] -> 16 rows
        const int n = i / kMaskH, k = i % kMaskH;
        __nv_bfloat16 h, l;
        split_bf16(n < (int)n_inst ? w2[n * kMaskH + k] : 0.f, h, l);
        img[kOffM2 + img_index(n, k, kMaskNOut)] = h;
        img[kOffM2 + kMaskNOut * kMaskH + img_index(n, k, kMaskNOut)] = l;
    }
}

// one K-chunk of a layer: for every 16-wide k-step the three split-precision products
//   a_col: first A column of the chunk (hi part; the lo part sits kColsAlo columns further)
template <int N, int KC>
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, uint32_t a_col, uint32_t img_saddr, uint32_t first_accumulate) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    constexpr uint32_t kstep_bytes = 2 * N * 16, lo_off = N * KC * 2;
#pragma unroll
    for (int j = 0; j < KC / 16; j++) {
        const uint64_t bh = tc::smem_desc_kmajor(img_saddr + j * kstep_bytes, N * 16, 128);
        const uint64_t bl = tc::smem_desc_kmajor(img_saddr + lo_off + j * kstep_bytes, N * 16, 128);
        mma_bf16_ts(d_tmem, a_col + 8 * j, bh, idesc, (j > 0) ? 1u : first_accumulate);
        mma_bf16_ts(d_tmem, a_col + 8 * j, bl, idesc, 1u);
        mma_bf16_ts(d_tmem, a_col + kColsAlo + 8 * j, bh, idesc, 1u);
    }
}

// The object head keeps the A operand INTERLEAVED: k-step j (16 k values) owns 16 columns, bf16 hi pairs in the first 8 and
// bf16 lo pairs in the last 8 -- exactly the footprint of the 16 fp32 accumulator columns it is computed from, so that an
// epilogue can convert D to the next layer's A in place.  STEPS k-steps of an [N x KC] image starting at img_saddr.
template <int N, int STEPS>
__device__ __forceinline__ void issue_ksteps(uint32_t d_tmem, uint32_t a_col, uint32_t img_saddr, uint32_t lo_off, uint32_t first_accumulate) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    constexpr uint32_t kstep_bytes = 2 * N * 16;
#pragma unroll
    for (int j = 0; j < STEPS; j++) {
        const uint64_t bh = tc::smem_desc_kmajor(img_saddr + j * kstep_bytes, N * 16, 128);
        const uint64_t bl = tc::smem_desc_kmajor(img_saddr + lo_off + j * kstep_bytes, N * 16, 128);
        mma_bf16_ts(d_tmem, a_col + 16 * j, bh, idesc, (j > 0) ? 1u : first_accumulate);
        mma_bf16_ts(d_tmem, a_col + 16 * j, bl, idesc, 1u);
        mma_bf16_ts(d_tmem, a_col + 16 * j + 8, bh, idesc, 1u);
    }
}

// Three roles in one CTA per SM (25 warps):
//   * warps 0..7, the EPILOGUE side: warps w and w+4 share the 32 TMEM lanes of sub-partition w; "part" 0 (warps 0..3) owns
//     output columns [0,128) of every layer, part 1 (warps 4..7) columns [128,256);
//   * warps 8..23, the PRODUCERS: while the MLP of tile i runs they gather m_grid (16 levels x 8 corners of 32-byte rows,
//     quarter-row loads) for the 128 samples of tile i+1 straight into the shared-memory input tile [143][128], from the
//     18-float records (point, geo_feat) the render kernel left per sample -- the 572-byte per-sample input never exists in HBM;
//   * warp 24, one thread: the ISSUER -- feeds the weight ring (TMA bulk copies, four 32 KB stages) and issues every tcgen05.mma.
// Each layer is issued as two output halves.  As soon as half 0 of a layer has been committed, part 0 converts it (leaky_relu,
// bf16 hi/lo split) IN PLACE to the first half of the next layer's A operand while the tensor cores are busy with half 1;
// the next layer then starts on the K range that is ready while part 1 converts the second half.  The tensor pipe only waits
// for an epilogue at the end of a tile.  TMEM: two 256-column regions that swap roles (A / D) from layer to layer.
// Measured and rejected: x-paired producer gathers (8 lanes per sample, the x-neighbours of a corner fetched by adjacent lanes:
// 1/3 fewer L1 line lookups, 40 % more producer instructions): 13.1 -> 15.7 ms per frame -- the producers are issue-bound as
// much as L1-bound.  (The same pairing pays in the render kernel's final stage, whose C=2 gathers are L1-bound: render.cu.)
// Also without effect (round 2): sharing the cell lookups among the four lanes of a sample (one lookup + 9 SHFL per level instead
// of four lookups, -25 % producer instructions): 13.45 -> 13.42 ms; the same with four levels = 32 loads in flight per lane: 13.28.
// Neither fewer producer instructions nor more loads in flight move the kernel: per tile the tensor pipe is busy 5.2 of 12.2 us
// and the rest is the serial MMA <-> epilogue chain of ONE tile -- tensor memory (2 x 256 columns) holds a single tile in flight.
#ifndef SANERF_MASK_PRODUCER_WARPS
#define SANERF_MASK_PRODUCER_WARPS 16
#endif
constexpr int kProdWarps = SANERF_MASK_PRODUCER_WARPS;   // one sample per lane quad and tile (8 warps x two samples: 2 ms slower)
static_assert(kProdWarps == 16, "128 samples per tile = producer warps x 8 samples");
constexpr int kProdThreads = 32 * kProdWarps;
constexpr int kIssuerWarp = kHeadThreads / 32 + kProdWarps;
constexpr int kMaskThreads = kHeadThreads + kProdThreads + 64, kRecK = 18;

__global__ void __launch_bounds__(kMaskThreads, 1)
    mask_head_kernel(const float* __restrict__ rec, const float* __restrict__ weights, const __grid_constant__ GridDev mg,
                     const __nv_bfloat16* __restrict__ img, float* __restrict__ logits, uint32_t n_tiles, uint32_t n_rays, uint32_t n_inst) {
    extern __shared__ __align__(128) uint8_t smem[];   // [4 stages x 32 KB][layer-2 image 16 KB][input tile 143 x 128 fp32]
    __shared__ __align__(8) uint64_t bar_full[kMaskStages], bar_free[kMaskStages], bar_ain[2], bar_q[4], bar_d[2], bar_l2, bar_w2, bar_in_full[2], bar_in_free[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, part = (warp >> 2) & 1;
    const uint32_t stage_saddr = tc::smem_u32(smem), w2_saddr = stage_saddr + kMaskStages * kMaskStageBytes;

    // barriers + tensor memory; the resident layer-2 image arrives by TMA bulk copy
    if (tid == 0) {
        for (int i = 0; i < kMaskStages; i++) {
            tc::mbar_init(&bar_full[i], 1);
            tc::mbar_init(&bar_free[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&bar_ain[i], kHeadThreads / 2);   // the 128 threads of one part: "my half of the layer-0 input is in TMEM"
            tc::mbar_init(&bar_d[i], 1);                    // tcgen05.commit: "output half i of the layer is complete"
        }
        for (int i = 0; i < 4; i++) tc::mbar_init(&bar_q[i], kHeadThreads / 2);   // "columns [64 i, 64 i + 64) are the next layer's A"

        tc::mbar_init(&bar_l2, 1);
        tc::mbar_init(&bar_w2, 1);
        for (int i = 0; i < 2; i++) {                    // the input tile is handed over in two halves: rows k < 64 | k >= 64
            tc::mbar_init(&bar_in_full[i], kProdThreads);       // every producer thread arrives (+ the geo_feat bytes on half 1)
            tc::mbar_init(&bar_in_free[i], kHeadThreads / 2);   // every thread of the part that reads the half arrives
        }
        tc::fence_mbar_init();
        tc::mbar_expect_tx(&bar_w2, kImg2 * 2);
        tc::tma_load_1d(w2_saddr, img + kOffM2, kImg2 * 2, &bar_w2);
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_base_s;
    const uint32_t r0 = tm, r1 = tm + 256;   // the two regions
    auto wait = [](uint64_t* bar, uint32_t parity) { tc::mbar_wait(bar, parity); };
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total_chunks = my_tiles * kMaskChunks;
    auto load_chunk = [&](uint32_t g) {   // chunk g of this CTA's stream -> stage g % kMaskStages (whole warp calls, one lane issues)
        const uint32_t c = g % kMaskChunks, s = g % kMaskStages;
        const void* src = c < kMaskChunks0 ? img + c * kImgM0 : img + kOffM1 + (c - kMaskChunks0) * kImgM1;
        const uint32_t bytes = (c < kMaskChunks0 ? kImgM0 : kImgM1) * 2;
        if (tc::elect_one()) {
            tc::mbar_expect_tx(&bar_full[s], bytes);
            tc::tma_load_1d(stage_saddr + s * kMaskStageBytes, src, bytes, &bar_full[s]);
        }
        __syncwarp();
    };
    float* xin = reinterpret_cast<float*>(smem + kMaskStages * kMaskStageBytes + kImg2 * 2);   // [143][128] input tile, filled by the producers

    if (warp == kIssuerWarp) {
        // ---- issuer: one warp owns the weight ring and the tensor pipe (warp-uniform control flow, one elected lane issues) ----
        uint32_t ph_full = 0, ph_ain = 0, ph_q = 0;   // one parity bit per barrier
        auto wait_bit = [&](uint64_t* bars, uint32_t& ph, int i) {
            wait(&bars[i], (ph >> i) & 1);
            ph ^= 1u << i;
        };
        uint32_t g = 0;
        for (uint32_t t = 0; t < my_tiles; t++) {
            const uint32_t X = (t & 1) ? r1 : r0, Y = (t & 1) ? r0 : r1;   // this tile: A0 in X, D0 -> Y, D1 -> X, D2 -> Y
#pragma unroll 1
            for (int c = 0; c < kMaskChunks; c++, g++) {
                // layer 0, K chunk 0 reads input rows k < 48 (part 0's), the later chunks part 1's too;
                // layer 1, K chunk kc of output half 0 is the first to read activation quarter kc
                if (c < 2) wait_bit(bar_ain, ph_ain, c);
                else if (c >= kMaskChunks0 && c < kMaskChunks0 + kKc1N) wait_bit(bar_q, ph_q, c - kMaskChunks0);
                const uint32_t s = g % kMaskStages, saddr = stage_saddr + s * kMaskStageBytes;
                wait(&bar_full[s], (ph_full >> s) & 1);
                ph_full ^= 1u << s;
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    if (c < kMaskChunks0) {          // layer 0
                        const int h = c / kKc0N, kc = c % kKc0N;
                        issue_ksteps<kHalfN, kKc0 / 16>(Y + h * kHalfN, X + kA0Col + kc * kKc0, saddr, kHalfN * kKc0 * 2, kc > 0);
                    } else {                         // layer 1
                        const int cc = c - kMaskChunks0, h = cc / kKc1N, kc = cc % kKc1N;
                        issue_ksteps<kHalfN, kKc1 / 16>(X + h * kHalfN, Y + kc * kKc1, saddr, kHalfN * kKc1 * 2, kc > 0);
                    }
                    tc::mma_commit(&bar_free[s]);
                    if (c == kKc0N - 1 || c == kMaskChunks0 + kKc1N - 1) tc::mma_commit(&bar_d[0]);
                    if (c == kMaskChunks0 - 1 || c == kMaskChunks - 1) tc::mma_commit(&bar_d[1]);
                }
                __syncwarp();
            }
            // layer 2: 256 -> n_inst (16 output columns), resident image, one activation quarter (4 k-steps) at a time
            if (t == 0) wait(&bar_w2, 0);
#pragma unroll 1
            for (int qd = 0; qd < 4; qd++) {
                wait_bit(bar_q, ph_q, qd);
                tc::fence_after_sync();
                if (tc::elect_one()) {
                    issue_ksteps<kMaskNOut, 4>(Y, X + 64 * qd, w2_saddr + qd * 4 * (2 * kMaskNOut * 16), kMaskNOut * kMaskH * 2, qd > 0);
                    if (qd == 3) tc::mma_commit(&bar_l2);
                }
                __syncwarp();
            }
        }
    } else if (warp == kIssuerWarp + 1) {
        // ---- weight loader: keeps the ring full on its own, so that the issuer never waits for a stage to drain ---------
        for (uint32_t g = 0; g < total_chunks; g++) {
            if (g >= kMaskStages) wait(&bar_free[g % kMaskStages], ((g / kMaskStages) - 1) & 1);
            load_chunk(g);
        }
    } else if (warp >= kHeadThreads / 32) {
        // ---- producer warps: build the input tile of every tile of this CTA, one tile ahead of the tensor cores ----------
        const int pw = warp - kHeadThreads / 32, s8 = lane >> 2, qp = lane & 3, ptid = tid - kHeadThreads;
        const int rowA = 8 * pw + s8;
        uint32_t ph_in = 0;
        bool first = true;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const float* r = rec + (size_t)tile * kRecK * 128;
            float xa[3];
#pragma unroll
            for (int d = 0; d < 3; d++) xa[d] = __ldg(r + d * 128 + rowA);
            bool ina = true;   // points outside [0,1]^3 give zero features (gridencoder.cu:105-130)
#pragma unroll
            for (int d = 0; d < 3; d++) ina &= !(xa[d] < 0.f || xa[d] > 1.f);
            if (!first) tc::mbar_wait(&bar_in_free[0], ph_in);   // part 0 has copied rows k < 64 of the previous tile out of the buffer
#pragma unroll 1
            for (int l = 0; l < 16; l += 2) {
                if (l == 8) {
                    // levels 0..7 = rows k < 64 are complete: part 0 converts them while levels 8..15 are gathered
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_in_full[0])) : "memory");
                    if (!first) tc::mbar_wait(&bar_in_free[1], ph_in);
                    if (ptid == 0) {   // geo_feat rows [128,143): 7680 contiguous bytes on both sides, one TMA bulk copy
                        asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(tc::smem_u32(&bar_in_full[1])),
                                     "r"(15u * 128u * 4u)
                                     : "memory");
                        tc::tma_load_1d(tc::smem_u32(xin + 128 * 128), r + 3 * 128, 15 * 128 * 4, &bar_in_full[1]);
                    }
                }
                // two levels per iteration: 16 independent loads in flight per lane
                float a0, a1, b0, b1;
                quarter_level(mg, l, xa, qp, a0, a1);
                quarter_level(mg, l + 1, xa, qp, b0, b1);
                float* d0 = xin + (8 * l + 2 * qp) * 128;
                d0[rowA] = ina ? a0 : 0.f;
                d0[128 + rowA] = ina ? a1 : 0.f;
                d0[8 * 128 + rowA] = ina ? b0 : 0.f;
                d0[9 * 128 + rowA] = ina ? b1 : 0.f;
            }
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_in_full[1])) : "memory");
            if (!first) ph_in ^= 1;
            first = false;
        }
    } else {
        // ---- epilogue side ----------------------------------------------------------------------------------------------
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        uint32_t ph_in = 0, ph_d = 0, ph_l2 = 0;
        // layer-0 input of the next tile: this thread's row, k in [0,64) (part 0) or [64,144) (part 1), -> columns [16,160) of X
        auto convert_input = [&](uint32_t X) {
            wait(&bar_in_full[part], ph_in);
            ph_in ^= 1;
            const float* src = xin + q * 32 + lane;
            const int k_begin = part ? 64 : 0, n_grp = part ? 5 : 4;
#pragma unroll 1
            for (int grp = 0; grp < n_grp; grp++) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int k = k_begin + grp * 16 + i;
                    v[i] = k < kMaskK0 ? src[k * 128] : 0.f;
                }
                uint32_t hi[8], lo[8];
                pack_split16(v, hi, lo);
                tc::tmem_st8(X + lane_base + kA0Col + k_begin + grp * 16, hi);
                tc::tmem_st8(X + lane_base + kA0Col + k_begin + grp * 16 + 8, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_ain[part])) : "memory");
            // this thread has read its half of the input row: hand that half of the buffer back to the producers
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_in_free[part])) : "memory");
        };
        // Output half h of a layer is complete in region `reg`: all eight warps convert it -- this warp the 64 columns
        // [128 h + 64 part, +64) of its 32 rows -- leaky_relu -> bf16 hi/lo, written over the same columns (activation quarter 2h+part)
        auto epilogue_in_place = [&](uint32_t reg, int h) {
            wait(&bar_d[h], (ph_d >> h) & 1);
            ph_d ^= 1u << h;
            tc::fence_after_sync();
            const uint32_t base = reg + lane_base + h * kHalfN + part * 64;
            // two tcgen05.ld in flight per step: the TMEM read latency of one 16-column group hides behind the other's math
#pragma unroll 1
            for (int grp = 0; grp < 4; grp += 2) {
                uint32_t t0[16], t1[16];
                tc::tmem_ld16(base + grp * 16, t0);
                tc::tmem_ld16(base + grp * 16 + 16, t1);
                tc::tmem_ld_wait();
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; i++) {
                        const float x = __uint_as_float(hh ? t1[i] : t0[i]);
                        v[i] = x > 0.f ? x : 0.01f * x;   // F.leaky_relu default slope (network.py:66)
                    }
                    uint32_t hi[8], lo[8];
                    pack_split16(v, hi, lo);
                    tc::tmem_st8(base + (grp + hh) * 16, hi);
                    tc::tmem_st8(base + (grp + hh) * 16 + 8, lo);
                }
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(&bar_q[2 * h + part])) : "memory");
        };
        if (my_tiles) convert_input(r0);
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const uint32_t X = (it & 1) ? r1 : r0, Y = (it & 1) ? r0 : r1;
#pragma unroll 1
            for (int e = 0; e < 4; e++) epilogue_in_place(e < 2 ? Y : X, e & 1);   // layer 0 (D in Y), then layer 1 (D in X)
            // every layer-1 MMA has completed: Y (A of layer 1) is dead except for the 16 columns layer 2 writes -> stage the
            // next tile's input there now, so that its layer 0 follows this tile's layer 2 without a gap
            if (tile + gridDim.x < n_tiles) convert_input(Y);
            if (part != 0) continue;
            wait(&bar_l2, ph_l2);
            ph_l2 ^= 1;
            tc::fence_after_sync();
            // ---- composite: logits[ray] = sum_samples w * point_masks (renderer.py:384); one warp = one ray ---------------
            {
                uint32_t t[16];
                tc::tmem_ld16(Y + lane_base, t);
                tc::tmem_ld_wait();
                const uint32_t ray = tile * 4 + q;
                const float w = ray < n_rays ? __ldg(weights + (size_t)ray * 32 + lane) : 0.f;
#pragma unroll
                for (int c = 0; c < kMaskNOut; c++) {
                    if (c < (int)n_inst) {   // uniform
                        float s = __fmul_rn(w, __uint_as_float(t[c]));
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                        if (lane == 0 && ray < n_rays) logits[(size_t)ray * n_inst + c] = s;
                    }
                }
            }
            // these columns are next written by layer 1 of the next tile, which waits for this warp's next epilogue (bar_q)
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}


// =====================================================================================================================
// SAM feature head: samvit = LayerNorm(SkipConnMLP(f)),  f = cat[f_sam(128), f_image(31), image(3), depth(1)]  (163)
// Reference: nerf/renderer.py:361-369, nerf/network.py:113-116, 31-66: five Linear layers WITH bias, width 256,
// leaky_relu(0.01) after all but the last, the 163-d input re-concatenated (hidden first) in front of layer 2
// (weight [256, 256+163]), then nn.LayerNorm(256, eps=1e-5).  Runs once per RAY (0.69 MFLOP each).
// Same machinery as the object head: 128 rays per tile, activations in TMEM (bf16 hi | lo), weight K-chunks streamed through
// the TMA-fed ring.  Layer 2 is two accumulating phases: A = hidden (K=256), then A = the input tile again (K=163 -> 176).
// The input tile (row-major [128,163] fp32, 83 KB) is staged once in shared memory and used by both phases.
// =====================================================================================================================
constexpr int kSamIn = 163, kSamW = 256;   // (the input is padded to 176 = 11 k-steps of 16)
struct SamChunk {
    uint32_t img_off;   // bf16 elements from the start of the image workspace
    uint16_t kc;        // K of the chunk (48 or 64)
    uint16_t a_col;     // first A column (hi part)
    uint8_t first;      // 1: first chunk of an accumulation (D is overwritten)
    uint8_t last;       // 1: last chunk before the activations are read back
    uint8_t layer;      // source weight matrix 0..4
    uint8_t pad;
    uint16_t k0;        // first source column of the chunk in that matrix
    uint16_t kvalid;    // number of real (non-padding) columns in the chunk
};
constexpr int kSamChunks = 22;
__constant__ SamChunk c_sam_chunks[kSamChunks];

__global__ void sam_prepare_kernel(const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                                   const float* __restrict__ w3, const float* __restrict__ w4, __nv_bfloat16* __restrict__ img) {
    const float* ws[5] = {w0, w1, w2, w3, w4};
    const int fan_in[5] = {kSamIn, kSamW, kSamW + kSamIn, kSamW, kSamW};
    for (int c = blockIdx.y; c < kSamChunks; c += gridDim.y) {
        const SamChunk ch = c_sam_chunks[c];
        const float* w = ws[ch.layer];
        const int fi = fan_in[ch.layer];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kSamW * ch.kc; i += gridDim.x * blockDim.x) {
            const int n = i / ch.kc, kk = i % ch.kc;
            __nv_bfloat16 h, l;
            split_bf16(kk < ch.kvalid ? w[(size_t)n * fi + ch.k0 + kk] : 0.f, h, l);
            img[ch.img_off + img_index(n, kk, kSamW)] = h;
            img[ch.img_off + kSamW * ch.kc + img_index(n, kk, kSamW)] = l;
        }
    }
}

template <int KC>
__device__ __forceinline__ void issue_chunk_rt(uint32_t d_tmem, uint32_t a_col, uint32_t img_saddr, uint32_t first_accumulate) {
    issue_chunk<kSamW, KC>(d_tmem, a_col, img_saddr, first_accumulate);
}

__global__ void __launch_bounds__(kHeadThreads, 1)
    samvit_mlp_kernel(const float* __restrict__ sam_in, const __nv_bfloat16* __restrict__ img, const float* __restrict__ b0,
                      const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ b3, const float* __restrict__ b4,
                      const float* __restrict__ ln_w, const float* __restrict__ ln_b, float* __restrict__ out, uint32_t n_tiles, uint32_t n_rays,
                      uint32_t out_nchw) {
    extern __shared__ __align__(128) uint8_t smem[];   // [2 stages x 64 KB][input tile 128 x 163 fp32]
    __shared__ __align__(8) uint64_t bar_full[2], bar_free[2], bar_done, bar_in;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = warp & 3, part = warp >> 2;
    const uint32_t stage_saddr = tc::smem_u32(smem);
    const float* xt = reinterpret_cast<const float*>(smem + 2 * kStageBytes);
    const uint32_t xt_saddr = stage_saddr + 2 * kStageBytes;
    constexpr uint32_t kInBytes = 128 * kSamIn * sizeof(float);
    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            tc::mbar_init(&bar_full[i], 1);
            tc::mbar_init(&bar_free[i], 1);
        }
        tc::mbar_init(&bar_done, 1);
        tc::mbar_init(&bar_in, 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tm = tmem_base_s, a_mma = tm, d_mma = tm + kColsD;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t a_rw = a_mma + lane_base, d_rw = d_mma + lane_base;
    uint32_t ph_full[2] = {0, 0}, ph_free[2] = {0, 0}, ph_done = 0;   // ph_full / ph_free are only used by warp 0
    const uint32_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total_chunks = my_tiles * kSamChunks;
    const int row = q * 32 + lane;

    auto load_chunk = [&](uint32_t g) {   // one thread: TMA bulk copy of chunk g into stage g & 1
        const SamChunk ch = c_sam_chunks[g % kSamChunks];
        const uint32_t bytes = 2u * kSamW * ch.kc * 2u;
        tc::mbar_expect_tx(&bar_full[g & 1], bytes);
        tc::tma_load_1d(stage_saddr + (g & 1) * kStageBytes, img + ch.img_off, bytes, &bar_full[g & 1]);
    };
    uint32_t g = 0;
    // one accumulation phase of n_chunks chunks (see mask_head_kernel::run_chunks)
    auto run_chunks = [&](int n_chunks) {
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) {   // warp-uniform ring bookkeeping; one elected lane issues the MMAs, commits and bulk copies
            tc::fence_after_sync();
            for (int c = 0; c < n_chunks; c++, g++) {
                const SamChunk ch = c_sam_chunks[g % kSamChunks];
                tc::mbar_wait(&bar_full[g & 1], ph_full[g & 1]);
                ph_full[g & 1] ^= 1;
                const uint32_t saddr = stage_saddr + (g & 1) * kStageBytes;
                if (tc::elect_one()) {
                    tc::fence_after_sync();
                    if (ch.kc == 64) issue_chunk_rt<64>(d_mma, a_mma + ch.a_col, saddr, ch.first ? 0u : 1u);
                    else issue_chunk_rt<48>(d_mma, a_mma + ch.a_col, saddr, ch.first ? 0u : 1u);
                    tc::mma_commit(&bar_free[g & 1]);
                    if (c == n_chunks - 1) tc::mma_commit(&bar_done);
                }
                __syncwarp();
                if (g + 1 < total_chunks) {
                    if (g >= 1) {
                        tc::mbar_wait(&bar_free[(g + 1) & 1], ph_free[(g + 1) & 1]);
                        ph_free[(g + 1) & 1] ^= 1;
                    }
                    if (tc::elect_one()) load_chunk(g + 1);
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        tc::mbar_wait(&bar_done, ph_done);
        ph_done ^= 1;
        tc::fence_after_sync();
    };
    // the 163-d input row of this thread (from the shared-memory tile) -> A columns; part 0: k [0,96), part 1: k [96,176)
    auto stage_input = [&] {
        const float* xr = xt + row * kSamIn;
        const int k_begin = part ? 96 : 0, n_grp = part ? 5 : 6;
#pragma unroll 1
        for (int grp = 0; grp < n_grp; grp++) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int k = k_begin + grp * 16 + i;
                v[i] = k < kSamIn ? xr[k] : 0.f;
            }
            uint32_t hi[8], lo[8];
            pack_split16(v, hi, lo);
            tc::tmem_st8(a_rw + (k_begin >> 1) + grp * 8, hi);
            tc::tmem_st8(a_rw + kColsAlo + (k_begin >> 1) + grp * 8, lo);
        }
        tc::tmem_st_wait();
    };
    // D[:, part*128 .. +128) + bias -> leaky_relu -> bf16 hi/lo -> A columns of the next layer
    auto epilogue_to_a = [&](const float* __restrict__ bias) {
#pragma unroll 1
        for (int grp = 0; grp < 8; grp += 2) {
            uint32_t t0[16], t1[16];
            tc::tmem_ld16(d_rw + part * 128 + grp * 16, t0);
            tc::tmem_ld16(d_rw + part * 128 + grp * 16 + 16, t1);
            tc::tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float x = __uint_as_float(h ? t1[i] : t0[i]) + __ldg(bias + part * 128 + (grp + h) * 16 + i);
                    v[i] = x > 0.f ? x : 0.01f * x;
                }
                uint32_t hi[8], lo[8];
                pack_split16(v, hi, lo);
                tc::tmem_st8(a_rw + part * 64 + (grp + h) * 8, hi);
                tc::tmem_st8(a_rw + kColsAlo + part * 64 + (grp + h) * 8, lo);
            }
        }
        tc::tmem_st_wait();
    };

    // The input tile ([128,163] fp32, 83 KB, contiguous; the buffer is padded to whole tiles) is prefetched by TMA one tile
    // ahead: the copy of tile i+1 starts right after tile i's second (skip-connection) use of the buffer.
    uint32_t ph_in = 0;
    if (tid == 0 && total_chunks) {
        load_chunk(0);
        tc::mbar_expect_tx(&bar_in, kInBytes);
        tc::tma_load_1d(xt_saddr, sam_in + (size_t)blockIdx.x * 128 * kSamIn, kInBytes, &bar_in);
    }
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        tc::mbar_wait(&bar_in, ph_in);
        ph_in ^= 1;
        stage_input();
        run_chunks(3);          // layer 0: 163 -> 256
        epilogue_to_a(b0);
        run_chunks(4);          // layer 1
        epilogue_to_a(b1);
        run_chunks(4);          // layer 2, hidden part (columns 0..255 of the [256,419] weight); its MMAs have read A when this returns
        stage_input();
        run_chunks(3);          // layer 2, skip part (columns 256..418), accumulates onto D
        if (tid == 0 && tile + gridDim.x < n_tiles) {   // all threads passed run_chunks' barrier after their last read of xt
            tc::mbar_expect_tx(&bar_in, kInBytes);
            tc::tma_load_1d(xt_saddr, sam_in + (size_t)(tile + gridDim.x) * 128 * kSamIn, kInBytes, &bar_in);
        }
        epilogue_to_a(b2);
        run_chunks(4);          // layer 3
        epilogue_to_a(b3);
        run_chunks(4);          // layer 4 (no activation)
        // ---- bias + LayerNorm(256, eps 1e-5, affine) + store; one thread per ray (part 0), three passes over its TMEM row ----
        if (part == 0) {
            const uint32_t ray = tile * 128 + row;
            float mean = 0.f;
#pragma unroll 1
            for (int grp = 0; grp < 16; grp++) {
                uint32_t t[16];
                tc::tmem_ld16(d_rw + grp * 16, t);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) mean += __uint_as_float(t[i]) + __ldg(b4 + grp * 16 + i);
            }
            mean *= (1.0f / kSamW);
            float var = 0.f;
#pragma unroll 1
            for (int grp = 0; grp < 16; grp++) {
                uint32_t t[16];
                tc::tmem_ld16(d_rw + grp * 16, t);
                tc::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const float dlt = __uint_as_float(t[i]) + __ldg(b4 + grp * 16 + i) - mean;
                    var = __fmaf_rn(dlt, dlt, var);
                }
            }
            const float rstd = rsqrtf(var * (1.0f / kSamW) + 1e-5f);
#pragma unroll 1
            for (int grp = 0; grp < 16; grp++) {
                uint32_t t[16];
                tc::tmem_ld16(d_rw + grp * 16, t);
                tc::tmem_ld_wait();
                if (ray < n_rays) {
                    float* dst = out + (size_t)ray * kSamW + grp * 16;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        float y[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int c = grp * 16 + i + j;
                            y[j] = (__uint_as_float(t[i + j]) + __ldg(b4 + c) - mean) * rstd * __ldg(ln_w + c) + __ldg(ln_b + c);
                        }
                        if (!out_nchw) {
                            *reinterpret_cast<float4*>(dst + i) = make_float4(y[0], y[1], y[2], y[3]);
                        } else {
                            // channel-major [256][n_rays] = the [1,256,H,W] tensor of trainer.py:540-541 (reshape + permute +
                            // contiguous) written directly: the 32 lanes of a warp are 32 consecutive rays -> coalesced rows
#pragma unroll
                            for (int j = 0; j < 4; j++) out[(size_t)(grp * 16 + i + j) * n_rays + ray] = y[j];
                        }
                    }
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm, 512);
}

// Output consumer of the feature frame (nerf/trainer.py:540-546): `samvit.reshape(1,h,w,C).permute(0,3,1,2)
// .contiguous()` followed by `F.interpolate(..., (Ho,Wo), mode='bilinear')` (align_corners=False) in ONE pass over the taps that
// are actually needed: out[c,i,j] = bilinear blend of in[y0..y1, x0..x1, c], source index max((dst + 0.5) * in/out - 0.5, 0) like
// ATen's upsample_bilinear2d.  One block = 32 output columns of one output row; threads = channels on the (coalesced) read side,
// a shared-memory transpose makes the [C][Ho][Wo] writes coalesced too.
__global__ void __launch_bounds__(256) feature_resize_nchw_kernel(const float* __restrict__ in, uint32_t h, uint32_t w, uint32_t C, uint32_t Ho,
                                                                  uint32_t Wo, float* __restrict__ out) {
    __shared__ float tile[256][33];
    const uint32_t i = blockIdx.y, j0 = blockIdx.x * 32;
    const float rh = (float)h / (float)Ho, rw = (float)w / (float)Wo;
    const float ys = fmaxf(rh * ((float)i + 0.5f) - 0.5f, 0.f);
    const uint32_t y0 = (uint32_t)ys, yp = y0 < h - 1 ? 1u : 0u;
    const float ly1 = ys - (float)y0, ly0 = 1.f - ly1;
    for (uint32_t c0 = 0; c0 < C; c0 += 256) {
        const uint32_t c = c0 + threadIdx.x;
        for (uint32_t jj = 0; jj < 32 && j0 + jj < Wo; jj++) {
            const float xs = fmaxf(rw * ((float)(j0 + jj) + 0.5f) - 0.5f, 0.f);
            const uint32_t x0 = (uint32_t)xs, xp = x0 < w - 1 ? 1u : 0u;
            const float lx1 = xs - (float)x0, lx0 = 1.f - lx1;
            if (c < C) {
                const float* p = in + ((size_t)y0 * w + x0) * C + c;
                const float v00 = __ldg(p), v01 = __ldg(p + (size_t)xp * C), v10 = __ldg(p + (size_t)yp * w * C),
                            v11 = __ldg(p + ((size_t)yp * w + xp) * C);
                tile[threadIdx.x][jj] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
            }
        }
        __syncthreads();
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (uint32_t cc = warp; cc < 256 && c0 + cc < C; cc += 8)
            if (j0 + lane < Wo) out[((size_t)(c0 + cc) * Ho + i) * Wo + j0 + lane] = tile[cc][lane];
        __syncthreads();
    }
}

static size_t sam_build_chunks(SamChunk* out) {
    // (layer, k0 in the source matrix, number of valid columns) per accumulation phase; each is cut into 64/48-wide chunks
    struct Phase { int layer, k0, kvalid; };
    const Phase phases[6] = {{0, 0, kSamIn}, {1, 0, kSamW}, {2, 0, kSamW}, {2, kSamW, kSamIn}, {3, 0, kSamW}, {4, 0, kSamW}};
    size_t off = 0;
    int n = 0;
    for (const Phase& ph : phases) {
        const int kp = (ph.kvalid + 15) / 16 * 16;
        for (int k = 0; k < kp;) {
            const int kc = (kp - k) >= 64 ? 64 : (kp - k);   // 176 = 64 + 64 + 48
            SamChunk& c = out[n++];
            c.img_off = (uint32_t)off;
            c.kc = (uint16_t)kc;
            c.a_col = (uint16_t)(k / 2);
            // layer 2's second phase accumulates onto the first
            c.first = (k == 0 && !(ph.layer == 2 && ph.k0 != 0)) ? 1 : 0;
            c.last = (k + kc >= kp) ? 1 : 0;
            c.layer = (uint8_t)ph.layer;
            c.pad = 0;
            c.k0 = (uint16_t)(ph.k0 + k);
            c.kvalid = (uint16_t)((ph.kvalid - k) < kc ? (ph.kvalid - k > 0 ? ph.kvalid - k : 0) : kc);
            off += 2 * (size_t)kSamW * kc;
            k += kc;
        }
    }
    return n == kSamChunks ? off : 0;
}

}  // namespace sanerf

using namespace sanerf;

extern "C" {

size_t sanerf_mask_head_workspace_bytes(void) { return (size_t)kImgTotal * sizeof(__nv_bfloat16); }

int sanerf_mask_head(const float* records, const float* weights, const sanerf_grid_t* m_grid, const float* w0, const float* w1,
                     const float* w2, uint32_t n_inst, uint32_t n_rays, void* workspace, float* logits, sanerf_stream_t stream) {
    if (n_rays == 0) return 0;
    if (!records || !weights || !m_grid || !w0 || !w1 || !w2 || !workspace || !logits) return SANERF_E_NULL;
    if (n_inst == 0 || n_inst > (uint32_t)kMaskNOut || m_grid->num_levels != 16) return SANERF_E_CONFIG;
    if (((uintptr_t)records & 15) != 0) return SANERF_E_CONFIG;   // the geo_feat rows of a tile travel by TMA bulk copy
    GridDev mg;
    int rc = fill_grid(mg, *m_grid, 8);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(workspace);
    mask_prepare_kernel<<<64, 256, 0, st>>>(w0, w1, w2, n_inst, img);
    const size_t smem = (size_t)kMaskStages * kMaskStageBytes + (size_t)kImg2 * 2 + (size_t)kMaskK0 * 128 * sizeof(float);
    if (cudaFuncSetAttribute(mask_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return SANERF_E_SMEM;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_tiles = div_up(n_rays, 4u);
    mask_head_kernel<<<n_tiles < (uint32_t)sms ? n_tiles : (uint32_t)sms, kMaskThreads, smem, st>>>(records, weights, mg, img, logits, n_tiles, n_rays, n_inst);
    return check_launch();
}

size_t sanerf_samvit_mlp_workspace_bytes(void) {
    SamChunk tmp[kSamChunks + 8];
    return sam_build_chunks(tmp) * sizeof(__nv_bfloat16);
}

int sanerf_samvit_mlp(const float* sam_in, const float* const* w, const float* const* b, const float* ln_w, const float* ln_b, uint32_t n_rays,
                      void* workspace, float* out, sanerf_stream_t stream) {
    return sanerf_samvit_mlp_layout(sam_in, w, b, ln_w, ln_b, n_rays, workspace, out, 0, stream);
}

int sanerf_samvit_mlp_layout(const float* sam_in, const float* const* w, const float* const* b, const float* ln_w, const float* ln_b,
                             uint32_t n_rays, void* workspace, float* out, uint32_t out_nchw, sanerf_stream_t stream) {
    if (n_rays == 0) return 0;
    if (!sam_in || !w || !b || !ln_w || !ln_b || !workspace || !out) return SANERF_E_NULL;
    for (int i = 0; i < 5; i++)
        if (!w[i] || !b[i]) return SANERF_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    SamChunk chunks[kSamChunks + 8];
    if (sam_build_chunks(chunks) == 0) return SANERF_E_CONFIG;
    cudaError_t e = cudaMemcpyToSymbolAsync(c_sam_chunks, chunks, sizeof(SamChunk) * kSamChunks, 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(workspace);
    sam_prepare_kernel<<<dim3(16, kSamChunks), 256, 0, st>>>(w[0], w[1], w[2], w[3], w[4], img);
    const size_t smem = 2 * (size_t)kStageBytes + (size_t)128 * kSamIn * sizeof(float);
    if (cudaFuncSetAttribute(samvit_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return SANERF_E_SMEM;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t n_tiles = div_up(n_rays, 128u);
    samvit_mlp_kernel<<<n_tiles < (uint32_t)sms ? n_tiles : (uint32_t)sms, kHeadThreads, smem, st>>>(
        sam_in, img, b[0], b[1], b[2], b[3], b[4], ln_w, ln_b, out, n_tiles, n_rays, out_nchw);
    return check_launch();
}

int sanerf_feature_resize_nchw(const float* in_hwc, uint32_t h, uint32_t w, uint32_t C, uint32_t Ho, uint32_t Wo, float* out_chw,
                               sanerf_stream_t stream) {
    if (!in_hwc || !out_chw) return SANERF_E_NULL;
    if (h == 0 || w == 0 || C == 0 || Ho == 0 || Wo == 0) return 0;
    if (C > 1024) return SANERF_E_CONFIG;
    const dim3 grid(div_up(Wo, 32u), Ho);
    feature_resize_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in_hwc, h, w, C, Ho, Wo, out_chw);
    return check_launch();
}

}  // extern "C"
