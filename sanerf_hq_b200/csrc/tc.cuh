// tc.cuh -- thin inline-PTX layer over the Blackwell 5th-gen tensor core (tcgen05) for the tiny MLPs of the
// render path (sm_100a).
//
// Shape of the problem: a "group" of 4 warps (128 threads) owns 128 rows (= samples), one row per thread,
// which is exactly the TMEM datapath mapping: TMEM lane i <-> thread i of the group, and a warp may only touch
// the 32 lanes of its own sub-partition (warp_id % 4).  A layer  D[128,N] = A[128,K] * W[N,K]^T  is
//     tcgen05.st   (each thread writes its own A row into TMEM columns)      -- no shared-memory staging
//     tcgen05.mma  .kind::tf32, A from TMEM, B = W from shared memory (K-major, no swizzle), D in TMEM
//     tcgen05.ld   (each thread reads its own D row)
// fp32 accuracy comes from split-precision operands (3xTF32):  A = Ah + Al,  W = Wh + Wl  (each exactly
// representable in tf32),  D = Ah*Wh + Ah*Wl + Al*Wh  with fp32 accumulation in the tensor core; the dropped
// Al*Wl term is 2^-22 relative.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace sanerf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp, .sync.aligned) ----------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- ordering ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the tensor core's async proxy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_barrier(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// One lane of a converged warp.  Code under `if (elect_one())` is single-threaded AND known to be so by the compiler:
// tcgen05.mma / commit / bulk copies are then issued straight from uniform registers (under `if (lane == 0)` every one of
// them is wrapped in an ELECT + BRA.U.ANY loop, ~70 cycles of issue per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // suspend-time hint: sleep in hardware, do not spin
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(20000u)
            : "memory");
    } while (!done);
}
// the same without a suspend-time hint: for waits on the critical path of a hand-off chain
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// all previously issued tcgen05.mma of this thread arrive on `bar` when they complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA bulk copy (cp.async.bulk, the 1-D form of the tensor memory accelerator; SASS: UBLKCP) ----------------------------
// One thread arms the mbarrier with the byte count and issues the copy; the barrier completes when the bytes have landed in
// shared memory (written through the async proxy, which is also the proxy tcgen05.mma reads through: no proxy fence needed).
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_saddr), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- descriptors --------------------------------------------------------------------------------------------
// Shared-memory operand descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): core matrix = 8 rows x 16 B
// stored contiguously (128 B); SBO = byte distance between core matrices adjacent along N, LBO = along K.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::tf32 (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M x N.
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[tmem] * B[smem]^T ; one thread issues for the whole CTA
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---- TMEM <-> registers: shape 32x32b, thread i of the warp <-> lane (32*(warp%4) + i), 16 consecutive columns ---
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- split precision ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
// residual x - hi (exact in fp32), itself rounded to tf32 (the tensor core would otherwise truncate it): |x - hi - lo| <= 2^-23 |x|
__device__ __forceinline__ uint32_t tf32_lo(float x, uint32_t hi) { return tf32_hi(x - __uint_as_float(hi)); }

// ---- weights in shared memory ---------------------------------------------------------------------------------
// W [N,K] (nn.Linear layout, row n = output unit) is stored as the canonical K-major no-swizzle operand
//     float index(n,k) = ((k/4)*N + n)*4 + k%4      (core matrix = 8 consecutive n x 4 consecutive k = 128 B)
// so that LBO (next 4 k) = 16*N bytes and SBO (next 8 n) = 128 bytes; one MMA (K=8) starts at k-step ks:
//     base + ks*2*16*N bytes.
__host__ __device__ constexpr int w_index(int n, int k, int N) { return ((k >> 2) * N + n) * 4 + (k & 3); }

// Stage W (global, [N,K] row-major, fp32) into hi / lo operand images (each N*KP floats, KP = K rounded up to 8; zero padded).
template <int N, int K, int KP>
__device__ __forceinline__ void stage_split_weights(float* s_hi, float* s_lo, const float* __restrict__ w, int tid, int nthreads) {
    for (int i = tid; i < N * KP; i += nthreads) {
        const int n = i / KP, k = i % KP;
        const float v = k < K ? __ldg(w + n * K + k) : 0.f;
        const uint32_t h = tf32_hi(v);
        const uint32_t l = tf32_hi(v - __uint_as_float(h));
        s_hi[w_index(n, k, N)] = __uint_as_float(h);
        s_lo[w_index(n, k, N)] = __uint_as_float(l);
    }
}

// Issue D[128,N] (+)= A[128,8*KSTEPS] * W^T for one operand image.  a_tmem: first column of A (lane field 0).
template <int N, int KSTEPS>
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_tmem, const float* s_w, uint32_t first_accumulate) {
    constexpr uint32_t idesc = idesc_tf32(128, N);
    const uint32_t wbase = smem_u32(s_w);
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ks++) {
        const uint64_t bd = smem_desc_kmajor(wbase + ks * 2 * 16 * N, 16 * N, 128);
        mma_tf32_ts(d_tmem, a_tmem + 8 * ks, bd, idesc, (ks > 0) ? 1u : first_accumulate);
    }
}

}  // namespace tc
}  // namespace sanerf

// ================================================================================================================
// Group MLP: 4 warps = 128 rows, one row per thread.  TMEM budget per group: 128 columns
//     [0,64)   A operand (activations, tf32 hi [and lo when 2K <= 64])
//     [64,128) D accumulator (fp32)
// A layer with 2K <= 64 runs in one round (hi and lo side by side); a K = 64 layer runs in two rounds
// (hi: D = Ah*Wh + Ah*Wl, then lo over the same columns: D += Al*Wh).
// ================================================================================================================
namespace sanerf {
namespace tc {

constexpr uint32_t kGroupCols = 128, kACols = 64;

struct Group {
    uint32_t a_mma, d_mma;  // TMEM addresses (lane field 0) of the A / D regions for the MMA
    uint32_t a_rw, d_rw;    // same columns with this warp's lane base (32 * (warp % 4)) for tcgen05.ld / st
    uint64_t* bar;          // mbarrier the MMAs commit to
    uint32_t phase;         // parity of the next completion
    uint32_t bar_id;        // named barrier of the group (128 threads)
    bool issuer;            // the one thread of the group that takes / returns the TMEM slot
    bool lead;              // warp 0 of the group: one elected lane of it issues tcgen05.mma
};

__device__ __forceinline__ Group make_group(uint32_t tmem_base, int group, int warp_in_group, int lane, uint64_t* bar) {
    Group g;
    g.a_mma = tmem_base + group * kGroupCols;
    g.d_mma = g.a_mma + kACols;
    g.a_rw = g.a_mma + ((uint32_t)(warp_in_group * 32) << 16);
    g.d_rw = g.d_mma + ((uint32_t)(warp_in_group * 32) << 16);
    g.bar = bar;
    g.phase = 0;
    g.bar_id = 1 + group;
    g.issuer = (warp_in_group == 0 && lane == 0);
    g.lead = warp_in_group == 0;
    return g;
}

// ---- TMEM slots ---------------------------------------------------------------------------------------------------------
// A CTA with more than 4 groups time-shares the four 128-column slots of its tensor memory: a group holds a slot only for the
// duration of its tensor-core rounds (a small fraction of a ray's time).  `free_mask` (shared memory) has bit s set when slot
// s is free.  acquire / release are executed by all 128 threads of the group.
__device__ __forceinline__ void group_acquire(Group& g, unsigned int* free_mask, volatile int* slot_bcast, uint32_t tmem_base, int warp_in_group) {
    if (g.issuer) {
        int s;
        for (;;) {
            const unsigned int m = *reinterpret_cast<volatile unsigned int*>(free_mask);
            if (m) {
                s = __ffs(m) - 1;
                if (atomicAnd(free_mask, ~(1u << s)) & (1u << s)) break;
            } else {
                __nanosleep(100);
            }
        }
        *slot_bcast = s;
    }
    named_barrier(g.bar_id, 128);            // publishes the slot to the group (bar.sync orders shared-memory accesses)
    const int s = *slot_bcast;
    g.a_mma = tmem_base + s * kGroupCols;
    g.d_mma = g.a_mma + kACols;
    g.a_rw = g.a_mma + ((uint32_t)(warp_in_group * 32) << 16);
    g.d_rw = g.d_mma + ((uint32_t)(warp_in_group * 32) << 16);
    fence_after_sync();                      // the previous owner's TMEM reads were ordered before its release
}
__device__ __forceinline__ void group_release(Group& g, unsigned int* free_mask, volatile int* slot_bcast) {
    fence_before_sync();                     // this thread's tcgen05.ld have completed (tmem_ld_wait)
    named_barrier(g.bar_id, 128);
    if (g.issuer) atomicOr(free_mask, 1u << *slot_bcast);
}

// write hi (PART 0) or lo (PART 1) parts of a[K] into K consecutive TMEM columns, 8/16 columns per instruction so that
// only one chunk of converted values is live at a time
template <int K, int PART>
__device__ __forceinline__ void st_split(uint32_t taddr, const float (&a)[K]) {
    static_assert(K % 8 == 0, "A rows are written 8 or 16 columns at a time");
    constexpr int CH = (K % 16 == 0) ? 16 : 8;
#pragma unroll
    for (int c = 0; c < K; c += CH) {
        uint32_t t[CH];
#pragma unroll
        for (int i = 0; i < CH; i++) {
            const uint32_t h = tf32_hi(a[c + i]);
            t[i] = PART == 0 ? h : tf32_lo(a[c + i], h);
        }
        if constexpr (CH == 16) tmem_st16(taddr + c, t);
        else tmem_st8(taddr + c, t);
    }
}

// wait for every thread's A row, let the issuer launch the MMAs, wait for them to complete
template <class IssueFn>
__device__ __forceinline__ void group_round(Group& g, IssueFn&& issue) {
    tmem_st_wait();
    fence_before_sync();
    named_barrier(g.bar_id, 128);
    if (g.lead) {              // warp-uniform branch + elect: the MMAs are issued back to back from uniform registers
        if (elect_one()) {
            fence_after_sync();
            issue();
            mma_commit(g.bar);
        }
    }
    __syncwarp();
    mbar_wait(g.bar, g.phase);
    g.phase ^= 1;
    fence_after_sync();
}

// d[N] = (RELU?) a[K] . W[N,K]^T for the 128 rows of the group; s_hi / s_lo = operand images of W (stage_split_weights)
template <int K, int N, bool RELU>
__device__ __forceinline__ void group_layer(Group& g, const float* s_hi, const float* s_lo, const float (&a)[K], float (&d)[N]) {
    static_assert(K % 8 == 0 && K <= 64 && N % 16 == 0 && N <= 64, "layer shape");
    const uint32_t d_mma = g.d_mma, a_mma = g.a_mma;
    if constexpr (2 * K <= (int)kACols) {
        st_split<K, 0>(g.a_rw, a);
        st_split<K, 1>(g.a_rw + K, a);
        group_round(g, [&] {
            issue_layer<N, K / 8>(d_mma, a_mma, s_hi, 0u);
            issue_layer<N, K / 8>(d_mma, a_mma, s_lo, 1u);
            issue_layer<N, K / 8>(d_mma, a_mma + K, s_hi, 1u);
        });
    } else {
        // hi pass: write tf32(a) and keep the exact residual a - tf32(a) for the lo pass (3 instead of 4 conversions per value)
        float res[K];
#pragma unroll
        for (int c = 0; c < K; c += 16) {
            uint32_t t[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                t[i] = tf32_hi(a[c + i]);
                res[c + i] = a[c + i] - __uint_as_float(t[i]);
            }
            tmem_st16(g.a_rw + c, t);
        }
        group_round(g, [&] {
            issue_layer<N, K / 8>(d_mma, a_mma, s_hi, 0u);
            issue_layer<N, K / 8>(d_mma, a_mma, s_lo, 1u);
        });
#pragma unroll
        for (int c = 0; c < K; c += 16) {
            uint32_t t[16];
#pragma unroll
            for (int i = 0; i < 16; i++) t[i] = tf32_hi(res[c + i]);
            tmem_st16(g.a_rw + c, t);
        }
        group_round(g, [&] { issue_layer<N, K / 8>(d_mma, a_mma, s_hi, 1u); });
    }
#pragma unroll
    for (int c = 0; c < N; c += 16) {
        uint32_t t[16];
        tmem_ld16(g.d_rw + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float v = __uint_as_float(t[i]);
            d[c + i] = RELU ? fmaxf(v, 0.f) : v;
        }
    }
}

// Two independent 128-row tiles (each thread owns one row of each) through the same layer in ONE round: tile t uses A columns
// [t*2K, (t+1)*2K) (hi | lo) and D columns [t*N, (t+1)*N).  Needs 4K <= 64 and 2N <= 64.
template <int K, int N, bool RELU>
__device__ __forceinline__ void group_layer_x2(Group& g, const float* s_hi, const float* s_lo, const float (&a0)[K], const float (&a1)[K],
                                               float (&d0)[N], float (&d1)[N]) {
    static_assert(K % 8 == 0 && 4 * K <= (int)kACols && N % 16 == 0 && 2 * N <= 64, "layer shape (two tiles)");
    const uint32_t d_mma = g.d_mma, a_mma = g.a_mma;
    st_split<K, 0>(g.a_rw, a0);
    st_split<K, 1>(g.a_rw + K, a0);
    st_split<K, 0>(g.a_rw + 2 * K, a1);
    st_split<K, 1>(g.a_rw + 3 * K, a1);
    group_round(g, [&] {
#pragma unroll
        for (int t = 0; t < 2; t++) {
            issue_layer<N, K / 8>(d_mma + t * N, a_mma + t * 2 * K, s_hi, 0u);
            issue_layer<N, K / 8>(d_mma + t * N, a_mma + t * 2 * K, s_lo, 1u);
            issue_layer<N, K / 8>(d_mma + t * N, a_mma + t * 2 * K + K, s_hi, 1u);
        }
    });
#pragma unroll
    for (int c = 0; c < 2 * N; c += 16) {
        uint32_t t[16];
        tmem_ld16(g.d_rw + c, t);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float v = __uint_as_float(t[i]);
            if (c + i < N) d0[c + i] = RELU ? fmaxf(v, 0.f) : v;
            else d1[c + i - N] = RELU ? fmaxf(v, 0.f) : v;
        }
    }
}

// ================================================================================================================
// bf16 split-precision layers (kind::f16): a = hi + lo with two bf16 (16 significant bits, 2^-18 relative), two K values per
// 32-bit TMEM column, so a K-wide layer needs K/2 + K/2 columns of A: a 64-wide hidden layer fits the 64-column A region with
// BOTH parts resident -> one round per layer and no activation ever waits in registers.  Used for the hidden layers of
// grid_mlp (whose outputs feed exp() and the compositing, not the index buffers) and for the 256-wide heads (heads.cu).
// ================================================================================================================
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// 16 fp32 values -> 8 columns of bf16 hi pairs + 8 columns of bf16 lo pairs (even k in the low half of the column).
// Paired conversions (cvt.rn.bf16x2.f32): 6 instructions per pair of values.
__device__ __forceinline__ void pack_split16(const float (&v)[16], uint32_t (&hi)[8], uint32_t (&lo)[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);     // .x (low half) = v[2i]
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h);
        const float r0 = v[2 * i] - __uint_as_float(hb << 16);                       // bf16 -> fp32 is a shift
        const float r1 = v[2 * i + 1] - __uint_as_float(hb & 0xffff0000u);
        const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
        hi[i] = hb;
        lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
}
// element index (bf16 units) inside an [N x K] K-major no-swizzle operand image: core matrix = 8 n x 8 k (16 B per row)
__host__ __device__ constexpr int bf16_img_index(int n, int k, int N) { return ((k >> 3) * N + n) * 8 + (k & 7); }
__host__ __device__ constexpr uint32_t idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// W [N,K] fp32 (nn.Linear layout) -> bf16 hi / lo operand images in shared memory (each N*K bf16)
template <int N, int K>
__device__ __forceinline__ void stage_split_weights_bf16(__nv_bfloat16* s_hi, __nv_bfloat16* s_lo, const float* __restrict__ w, int tid, int nthreads) {
    static_assert(K % 16 == 0, "bf16 MMA K step");
    for (int i = tid; i < N * K; i += nthreads) {
        const int n = i / K, k = i % K;
        __nv_bfloat16 h, l;
        split_bf16(__ldg(w + i), h, l);
        s_hi[bf16_img_index(n, k, N)] = h;
        s_lo[bf16_img_index(n, k, N)] = l;
    }
}
// D[128,N] = A[128,K] * W^T with the three split products per 16-wide k-step; A hi at a_col, A lo at a_col + K/2
template <int N, int K>
__device__ __forceinline__ void issue_layer_bf16(uint32_t d_tmem, uint32_t a_col, const __nv_bfloat16* s_hi, const __nv_bfloat16* s_lo) {
    constexpr uint32_t idesc = idesc_bf16(128, N);
    const uint32_t wh = smem_u32(s_hi), wl = smem_u32(s_lo);
#pragma unroll 1
    for (int j = 0; j < K / 16; j++) {
        const uint64_t bh = smem_desc_kmajor(wh + j * 2 * 16 * N, 16 * N, 128);
        const uint64_t bl = smem_desc_kmajor(wl + j * 2 * 16 * N, 16 * N, 128);
        mma_bf16_ts(d_tmem, a_col + 8 * j, bh, idesc, j > 0 ? 1u : 0u);
        mma_bf16_ts(d_tmem, a_col + 8 * j, bl, idesc, 1u);
        mma_bf16_ts(d_tmem, a_col + K / 2 + 8 * j, bh, idesc, 1u);
    }
}
// Hidden layer fed from tensor memory: the previous layer's pre-activations sit in the D region (K fp32 columns per row);
// each thread streams its row 16 columns at a time through ReLU + bf16 split into the A region, then one round of MMAs
// overwrites the D region with this layer's pre-activations.  Nothing but a 16-value chunk is ever live in registers.
template <int K, int N>
__device__ __forceinline__ void group_layer_from_tmem(Group& g, const __nv_bfloat16* s_hi, const __nv_bfloat16* s_lo) {
    static_assert(K % 16 == 0 && K <= (int)kACols && N % 16 == 0 && N <= 64, "layer shape");
#pragma unroll 1
    for (int c = 0; c < K; c += 16) {
        uint32_t t[16];
        tmem_ld16(g.d_rw + c, t);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] = fmaxf(__uint_as_float(t[i]), 0.f);
        uint32_t hi[8], lo[8];
        pack_split16(v, hi, lo);
        tmem_st8(g.a_rw + c / 2, hi);
        tmem_st8(g.a_rw + K / 2 + c / 2, lo);
    }
    const uint32_t d_mma = g.d_mma, a_mma = g.a_mma;
    group_round(g, [&] { issue_layer_bf16<N, K>(d_mma, a_mma, s_hi, s_lo); });
}

}  // namespace tc
}  // namespace sanerf
