// mlp_tc.cu -- standalone entry point for the tcgen05 group MLP of tc.cuh (bias-free, ReLU between layers, none after
// the last: reference nerf/network.py:9-29 `MLP`), used to validate the tensor-core path in isolation:
//     out[M,16] = relu(relu(x[M,K] W0^T) W1^T) W2^T       W0 [H,K], W1 [H,H], W2 [16,H]  (nn.Linear layout)
// It is the grid_mlp of the fused render kernel (K = 2*levels, H = 64 or 16) lifted out of the ray loop.
#include "common.cuh"
#include "tc.cuh"

namespace sanerf {

constexpr int kMlpThreads = 512;  // 4 groups of 4 warps

template <int K, int H>
struct MlpSmem {
    static constexpr int w0 = 0;                    // hi image, then lo image
    static constexpr int w1 = w0 + 2 * H * K;
    static constexpr int w2 = w1 + 2 * H * H;
    static constexpr int total = w2 + 2 * 16 * H;   // floats
};

template <int K, int H>
__global__ void __launch_bounds__(kMlpThreads, 1)
    mlp3_tc_kernel(const float* __restrict__ x, const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                   float* __restrict__ out, uint32_t M) {
    using S = MlpSmem<K, H>;
    extern __shared__ __align__(128) float sm[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, group = warp >> 2, wq = warp & 3;

    tc::stage_split_weights<H, K, K>(sm + S::w0, sm + S::w0 + H * K, w0, tid, kMlpThreads);
    tc::stage_split_weights<H, H, H>(sm + S::w1, sm + S::w1 + H * H, w1, tid, kMlpThreads);
    tc::stage_split_weights<16, H, H>(sm + S::w2, sm + S::w2 + 16 * H, w2, tid, kMlpThreads);
    tc::fence_proxy_async_smem();
    if (tid == 0) {
        for (int i = 0; i < 4; i++) tc::mbar_init(&bars[i], 1);
        tc::fence_mbar_init();
    }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    tc::Group g = tc::make_group(tmem_base_s, group, wq, lane, &bars[group]);

    const uint32_t tiles = div_up(M, 128u);
    for (uint32_t tile = blockIdx.x * 4 + group; tile < tiles; tile += gridDim.x * 4) {
        const uint32_t row = tile * 128 + wq * 32 + lane;
        float a[K];
#pragma unroll
        for (int k = 0; k < K; k++) a[k] = row < M ? __ldg(x + (size_t)row * K + k) : 0.f;
        float h1[H];
        tc::group_layer<K, H, true>(g, sm + S::w0, sm + S::w0 + H * K, a, h1);
        float h2[H];
        tc::group_layer<H, H, true>(g, sm + S::w1, sm + S::w1 + H * H, h1, h2);
        float o[16];
        tc::group_layer<H, 16, false>(g, sm + S::w2, sm + S::w2 + 16 * H, h2, o);
        if (row < M) {
#pragma unroll
            for (int c = 0; c < 16; c += 4)
                *reinterpret_cast<float4*>(out + (size_t)row * 16 + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base_s, 512);
}

template <int K, int H>
static int launch_mlp3(const float* x, const float* w0, const float* w1, const float* w2, float* out, uint32_t M, cudaStream_t st) {
    const size_t smem = (size_t)MlpSmem<K, H>::total * sizeof(float);
    auto kfn = mlp3_tc_kernel<K, H>;
    if (cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return SANERF_E_SMEM;
    }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint32_t need = div_up(M, 512u);
    kfn<<<need < (uint32_t)sms ? need : (uint32_t)sms, kMlpThreads, smem, st>>>(x, w0, w1, w2, out, M);
    return check_launch();
}

}  // namespace sanerf

using namespace sanerf;

extern "C" int sanerf_mlp3_tc(const float* x, const float* w0, const float* w1, const float* w2, float* out, uint32_t M, uint32_t K,
                              uint32_t H, sanerf_stream_t stream) {
    if (M == 0) return 0;
    if (!x || !w0 || !w1 || !w2 || !out) return SANERF_E_NULL;
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 32 && H == 64) return launch_mlp3<32, 64>(x, w0, w1, w2, out, M, st);
    if (K == 8 && H == 16) return launch_mlp3<8, 16>(x, w0, w1, w2, out, M, st);
    if (K == 16 && H == 32) return launch_mlp3<16, 32>(x, w0, w1, w2, out, M, st);
    return SANERF_E_CONFIG;
}
