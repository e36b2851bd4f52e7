// sh_encode.cu -- real spherical-harmonics direction encoder, degree 1..8 (sm_100a).
//
// Replaces the reference's shencoder/src/shencoder.cu kernels K6 (values + analytic dy/dx) and K7
// (input gradient) behind the C ABI of include/sanerf_b200.h.  The basis polynomials are the
// standard real SH in Cartesian form (same sign/ordering convention as the reference,
// shencoder.cu:49-121).  Instead of 3x64 hand-expanded derivative polynomials (:125-354) the
// derivatives come from evaluating the same polynomials on forward-mode dual numbers, so value
// and gradient cannot drift apart.
#include "common.cuh"

namespace sanerf {

struct Dual3 {
    float v, dx, dy, dz;
};
__device__ __forceinline__ Dual3 operator+(Dual3 a, Dual3 b) { return {a.v + b.v, a.dx + b.dx, a.dy + b.dy, a.dz + b.dz}; }
__device__ __forceinline__ Dual3 operator-(Dual3 a, Dual3 b) { return {a.v - b.v, a.dx - b.dx, a.dy - b.dy, a.dz - b.dz}; }
__device__ __forceinline__ Dual3 operator-(Dual3 a) { return {-a.v, -a.dx, -a.dy, -a.dz}; }
__device__ __forceinline__ Dual3 operator*(Dual3 a, Dual3 b) {
    return {a.v * b.v, a.dx * b.v + a.v * b.dx, a.dy * b.v + a.v * b.dy, a.dz * b.v + a.v * b.dz};
}
__device__ __forceinline__ Dual3 operator*(float s, Dual3 a) { return {s * a.v, s * a.dx, s * a.dy, s * a.dz}; }
__device__ __forceinline__ Dual3 operator*(Dual3 a, float s) { return s * a; }
__device__ __forceinline__ Dual3 operator+(Dual3 a, float s) { return {a.v + s, a.dx, a.dy, a.dz}; }
__device__ __forceinline__ Dual3 operator+(float s, Dual3 a) { return a + s; }
__device__ __forceinline__ Dual3 operator-(Dual3 a, float s) { return {a.v - s, a.dx, a.dy, a.dz}; }
__device__ __forceinline__ Dual3 operator-(float s, Dual3 a) { return {s - a.v, -a.dx, -a.dy, -a.dz}; }

__device__ __forceinline__ float constant_of(float, float c) { return c; }
__device__ __forceinline__ Dual3 constant_of(Dual3, float c) { return {c, 0.f, 0.f, 0.f}; }

// Degree-C real SH (C*C values) of (x,y,z) into o[].  T = float or Dual3.
template <typename T>
__device__ __forceinline__ void sh_basis(T x, T y, T z, uint32_t C, T* o) {
    o[0] = constant_of(x, 0.28209479177387814f);
    if (C <= 1) return;
    o[1] = -0.48860251190291987f * y;
    o[2] = 0.48860251190291987f * z;
    o[3] = -0.48860251190291987f * x;
    if (C <= 2) return;
    const T xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    o[4] = 1.0925484305920792f * xy;
    o[5] = -1.0925484305920792f * yz;
    o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    o[7] = -1.0925484305920792f * xz;
    o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    if (C <= 3) return;
    o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    o[10] = 2.8906114426405538f * xy * z;
    o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    o[14] = 1.4453057213202769f * z * (x2 - y2);
    o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
    const T x4 = x2 * x2, y4 = y2 * y2, z4 = z2 * z2;
    const T x6 = x4 * x2, y6 = y4 * y2, z6 = z4 * z2;
    if (C <= 4) return;
    o[16] = 2.5033429417967046f * (xy * (x2 - y2));
    o[17] = 1.7701307697799304f * (yz * (-3.0f * x2 + y2));
    o[18] = 0.9461746957575601f * (xy * (7.0f * z2 - 1.0f));
    o[19] = 0.6690465435572892f * (yz * (3.0f - 7.0f * z2));
    o[20] = 0.10578554691520431f * ((-30.0f * z2 + 35.0f * z4 + 3.0f));
    o[21] = 0.6690465435572892f * (xz * (3.0f - 7.0f * z2));
    o[22] = 0.47308734787878004f * ((x2 - y2) * (7.0f * z2 - 1.0f));
    o[23] = 1.7701307697799304f * (xz * (-x2 + 3.0f * y2));
    o[24] = 0.6258357354491761f * ((-6.0f * x2 * y2 + x4 + y4));
    if (C <= 5) return;
    o[25] = 0.6563820568401701f * (y * (10.0f * x2 * y2 - 5.0f * x4 - y4));
    o[26] = 8.302649259524165f * (xy * z * (x2 - y2));
    o[27] = -0.4892382994352504f * (y * (3.0f * x2 - y2) * (9.0f * z2 - 1.0f));
    o[28] = 4.793536784973324f * (xy * z * (3.0f * z2 - 1.0f));
    o[29] = 0.45294665119569694f * (y * (14.0f * z2 - 21.0f * z4 - 1.0f));
    o[30] = 0.1169503224534236f * (z * (-70.0f * z2 + 63.0f * z4 + 15.0f));
    o[31] = 0.45294665119569694f * (x * (14.0f * z2 - 21.0f * z4 - 1.0f));
    o[32] = 2.396768392486662f * (z * (x2 - y2) * (3.0f * z2 - 1.0f));
    o[33] = -0.4892382994352504f * (x * (x2 - 3.0f * y2) * (9.0f * z2 - 1.0f));
    o[34] = 2.075662314881041f * (z * (-6.0f * x2 * y2 + x4 + y4));
    o[35] = 0.6563820568401701f * (x * (10.0f * x2 * y2 - x4 - 5.0f * y4));
    if (C <= 6) return;
    o[36] = 1.3663682103838286f * (xy * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4));
    o[37] = 2.366619162231752f * (yz * (10.0f * x2 * y2 - 5.0f * x4 - y4));
    o[38] = 2.0182596029148963f * (xy * (x2 - y2) * (11.0f * z2 - 1.0f));
    o[39] = -0.9212052595149235f * (yz * (3.0f * x2 - y2) * (11.0f * z2 - 3.0f));
    o[40] = 0.9212052595149235f * (xy * (-18.0f * z2 + 33.0f * z4 + 1.0f));
    o[41] = 0.5826213625187313f * (yz * (30.0f * z2 - 33.0f * z4 - 5.0f));
    o[42] = 0.06356920226762842f * ((105.0f * z2 - 315.0f * z4 + 231.0f * z6 - 5.0f));
    o[43] = 0.5826213625187313f * (xz * (30.0f * z2 - 33.0f * z4 - 5.0f));
    o[44] = 0.46060262975746175f * ((x2 - y2) * (11.0f * z2 * (3.0f * z2 - 1.0f) - 7.0f * z2 + 1.0f));
    o[45] = -0.9212052595149235f * (xz * (x2 - 3.0f * y2) * (11.0f * z2 - 3.0f));
    o[46] = 0.5045649007287241f * ((11.0f * z2 - 1.0f) * (-6.0f * x2 * y2 + x4 + y4));
    o[47] = 2.366619162231752f * (xz * (10.0f * x2 * y2 - x4 - 5.0f * y4));
    o[48] = 0.6831841051919143f * ((15.0f * x2 * y4 - 15.0f * x4 * y2 + x6 - y6));
    if (C <= 7) return;
    o[49] = 0.7071627325245963f * (y * (-21.0f * x2 * y4 + 35.0f * x4 * y2 - 7.0f * x6 + y6));
    o[50] = 5.2919213236038f * (xy * z * (-10.0f * x2 * y2 + 3.0f * x4 + 3.0f * y4));
    o[51] = -0.5189155787202603f * (y * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + 5.0f * x4 + y4));
    o[52] = 4.151324629762082f * (xy * z * (x2 - y2) * (13.0f * z2 - 3.0f));
    o[53] = -0.15645893386229404f * (y * (3.0f * x2 - y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f));
    o[54] = 0.4425326924449826f * (xy * z * (-110.0f * z2 + 143.0f * z4 + 15.0f));
    o[55] = 0.0903316075825173f * (y * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f));
    o[56] = 0.06828427691200495f * (z * (315.0f * z2 - 693.0f * z4 + 429.0f * z6 - 35.0f));
    o[57] = 0.0903316075825173f * (x * (-135.0f * z2 + 495.0f * z4 - 429.0f * z6 + 5.0f));
    o[58] = 0.07375544874083044f * (z * (x2 - y2) * (143.0f * z2 * (3.0f * z2 - 1.0f) - 187.0f * z2 + 45.0f));
    o[59] = -0.15645893386229404f * (x * (x2 - 3.0f * y2) * (13.0f * z2 * (11.0f * z2 - 3.0f) - 27.0f * z2 + 3.0f));
    o[60] = 1.0378311574405206f * (z * (13.0f * z2 - 3.0f) * (-6.0f * x2 * y2 + x4 + y4));
    o[61] = -0.5189155787202603f * (x * (13.0f * z2 - 1.0f) * (-10.0f * x2 * y2 + x4 + 5.0f * y4));
    o[62] = 2.6459606618019f * (z * (15.0f * x2 * y4 - 15.0f * x4 * y2 + x6 - y6));
    o[63] = 0.7071627325245963f * (x * (-35.0f * x2 * y4 + 21.0f * x4 * y2 - x6 + 7.0f * y6));
}

// K6.  One thread per direction.  inputs [B,3]; outputs [B,C*C]; dy_dx [B,3,C*C] or NULL.
__global__ void __launch_bounds__(256) sh_forward_kernel(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B,
                                                          uint32_t C, float* __restrict__ dy_dx) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint32_t C2 = C * C;
    const float x = __ldg(inputs + (size_t)b * 3), y = __ldg(inputs + (size_t)b * 3 + 1), z = __ldg(inputs + (size_t)b * 3 + 2);
    float* out = outputs + (size_t)b * C2;
    if (!dy_dx) {
        float o[64];
        sh_basis<float>(x, y, z, C, o);
        for (uint32_t i = 0; i < C2; i++) out[i] = o[i];
    } else {
        Dual3 o[64];
        sh_basis<Dual3>({x, 1.f, 0.f, 0.f}, {y, 0.f, 1.f, 0.f}, {z, 0.f, 0.f, 1.f}, C, o);
        float* dx = dy_dx + (size_t)b * 3 * C2;
        for (uint32_t i = 0; i < C2; i++) {
            out[i] = o[i].v;
            dx[i] = o[i].dx;
            dx[C2 + i] = o[i].dy;
            dx[2 * C2 + i] = o[i].dz;
        }
    }
}

// K7.  grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]   (shencoder.cu:358-382)
__global__ void __launch_bounds__(256) sh_backward_kernel(const float* __restrict__ grad, uint32_t B, uint32_t D, uint32_t C,
                                                           const float* __restrict__ dy_dx, float* __restrict__ grad_inputs) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t b = t / D;
    if (b >= B) return;
    const uint32_t d = t - b * D, C2 = C * C;
    const float* g = grad + (size_t)b * C2;
    const float* dd = dy_dx + (size_t)b * D * C2 + (size_t)d * C2;
    float acc = grad_inputs[t];
    for (uint32_t ch = 0; ch < C2; ch++) acc = __fmaf_rn(__ldg(g + ch), __ldg(dd + ch), acc);
    grad_inputs[t] = acc;
}

}  // namespace sanerf

using namespace sanerf;

extern "C" {

int sanerf_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree, float* dy_dx,
                             sanerf_stream_t stream) {
    if (D != 3) return SANERF_E_DIM;
    if (degree < 1 || degree > 8) return SANERF_E_DEGREE;
    if (B == 0) return 0;
    if (!inputs || !outputs) return SANERF_E_NULL;
    sh_forward_kernel<<<div_up(B, 256), 256, 0, (cudaStream_t)stream>>>(inputs, outputs, B, degree, dy_dx);
    return check_launch();
}

int sanerf_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree, const float* dy_dx,
                              float* grad_inputs, sanerf_stream_t stream) {
    (void)inputs;
    if (D != 3) return SANERF_E_DIM;
    if (degree < 1 || degree > 8) return SANERF_E_DEGREE;
    if (B == 0) return 0;
    if (!grad || !dy_dx || !grad_inputs) return SANERF_E_NULL;
    sh_backward_kernel<<<div_up(B * D, 256), 256, 0, (cudaStream_t)stream>>>(grad, B, D, degree, dy_dx, grad_inputs);
    return check_launch();
}

}  // extern "C"
