// freq_encode.cu -- NeRF sinusoidal (frequency) encoder (sm_100a).
//
// Replaces the reference's freqencoder/src/freqencoder.cu kernels K8/K9 behind the C ABI of
// include/sanerf_b200.h.  out[b,:] = [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(deg-1) x), cos(..)]
// with D-wide groups; like the reference (built with -use_fast_math, freqencoder/backend.py:9)
// cos is evaluated as the SFU sine of (x + pi/2): __sinf(scalbnf(x, f) + phase) (:52-56).
// One thread per input coordinate produces its 1+2*deg outputs (the reference uses one thread
// per output element and re-reads the input 1+2*deg times).
#include "common.cuh"

namespace sanerf {

__global__ void __launch_bounds__(256) freq_forward_kernel(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg,
                                                            uint32_t C, float* __restrict__ outputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float x = __ldg(inputs + t);
    float* out = outputs + (size_t)b * C;
    out[d] = x;
    const float half_pi = 3.141592653589793f / 2;
    for (uint32_t f = 0; f < deg; f++) {
        const float a = scalbnf(x, (int)f);
        out[D + (2 * f) * D + d] = __sinf(a + 0.0f * half_pi);
        out[D + (2 * f + 1) * D + d] = __sinf(a + 1.0f * half_pi);
    }
}

// K9: grad_inputs[b,d] = g[d] + sum_f 2^f (g_sin * cos - g_cos * sin)   (freqencoder.cu:63-94)
__global__ void __launch_bounds__(256) freq_backward_kernel(const float* __restrict__ grad, const float* __restrict__ outputs,
                                                             uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                                             float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * D) return;
    const uint32_t b = t / D, d = t - b * D;
    const float* g = grad + (size_t)b * C;
    const float* o = outputs + (size_t)b * C;
    float r = __ldg(g + d);
    for (uint32_t f = 0; f < deg; f++) {
        const uint32_t s = D + (2 * f) * D + d, c = s + D;
        r += scalbnf(1.0f, (int)f) * (__ldg(g + s) * __ldg(o + c) - __ldg(g + c) * __ldg(o + s));
    }
    grad_inputs[t] = r;
}

}  // namespace sanerf

using namespace sanerf;

extern "C" {

int sanerf_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                               sanerf_stream_t stream) {
    if (C != D + 2 * D * deg) return SANERF_E_CHANNELS;
    if (B == 0 || D == 0) return 0;
    if (!inputs || !outputs) return SANERF_E_NULL;
    freq_forward_kernel<<<div_up(B * D, 256), 256, 0, (cudaStream_t)stream>>>(inputs, B, D, deg, C, outputs);
    return check_launch();
}

int sanerf_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                float* grad_inputs, sanerf_stream_t stream) {
    if (C != D + 2 * D * deg) return SANERF_E_CHANNELS;
    if (B == 0 || D == 0) return 0;
    if (!grad || !outputs || !grad_inputs) return SANERF_E_NULL;
    freq_backward_kernel<<<div_up(B * D, 256), 256, 0, (cudaStream_t)stream>>>(grad, outputs, B, D, deg, C, grad_inputs);
    return check_launch();
}

}  // extern "C"
