// freqencoder.cu -- NeRF sin/cos positional encoding for sm_100a.
//
// Replaces the reference's freqencoder extension (freqencoder/src/freqencoder.cu: kernel_freq :30-58,
// kernel_freq_backward :63-94).  Output layout per row: [x (D), sin(2^0 x) (D), cos(2^0 x) (D), sin(2^1 x) ...].
// The reference builds this extension with -use_fast_math and calls __sinf(scalbnf(x,f) + phase) with phase in
// {0, pi/2}; we use the same intrinsic on the same argument so values agree bit for bit on the same GPU.
//
// Decomposition: the reference uses one thread per OUTPUT element (re-reading the input C/D times, integer div/mod
// per element).  Here one thread owns one (row, d) pair: reads x once, produces its 1 + 2*deg outputs.
#include "common.cuh"

namespace rn {
namespace {

__global__ void __launch_bounds__(256)
freq_forward_kernel(const float* __restrict__ inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                    float* __restrict__ outputs) {
    const uint32_t total = B * D;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t b = t / D, d = t - b * D;
        const float x = __ldg(inputs + t);
        float* o = outputs + (size_t)b * C + d;
        o[0] = x;
        o += D;
        for (uint32_t f = 0; f < deg; ++f) {
            const float a = scalbnf(x, (int)f);
            o[0] = __sinf(a + 0.0f);
            o[D] = __sinf(a + 1.5707963705062866f);  // (float)(PI/2), the reference's phase shift for the cos column
            o += 2 * D;
        }
    }
}

// g_x[d] = g[d] + sum_f 2^f (g_sin * out_cos - g_cos * out_sin)          (freqencoder.cu:63-94)
__global__ void __launch_bounds__(256)
freq_backward_kernel(const float* __restrict__ grad, const float* __restrict__ outputs, uint32_t B, uint32_t D,
                     uint32_t deg, uint32_t C, float* __restrict__ grad_inputs) {
    const uint32_t total = B * D;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t b = t / D, d = t - b * D;
        const float* g = grad + (size_t)b * C + d;
        const float* o = outputs + (size_t)b * C + d;
        float result = __ldg(g);
        g += D;
        o += D;
        for (uint32_t f = 0; f < deg; ++f) {
            const float v = __fmaf_rn(__ldg(g), __ldg(o + D), -__fmul_rn(__ldg(g + D), __ldg(o)));
            result = __fmaf_rn(scalbnf(1.0f, (int)f), v, result);
            g += 2 * D;
            o += 2 * D;
        }
        grad_inputs[t] = result;
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" int rn_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C,
                                      float* outputs, void* stream) {
    RN_REQUIRE(C == D + 2 * D * deg, "output_dim must equal D + 2*D*degree");
    if (B == 0 || D == 0) return RN_OK;
    RN_REQUIRE(inputs && outputs, "null pointer");
    freq_forward_kernel<<<wave_grid((uint64_t)B * D, 256, 8), 256, 0, (cudaStream_t)stream>>>(inputs, B, D, deg, C, outputs);
    return finish_launch("rn_freq_encode_forward");
}

extern "C" int rn_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                                       uint32_t C, float* grad_inputs, void* stream) {
    RN_REQUIRE(C == D + 2 * D * deg, "output_dim must equal D + 2*D*degree");
    if (B == 0 || D == 0) return RN_OK;
    RN_REQUIRE(grad && outputs && grad_inputs, "null pointer");
    freq_backward_kernel<<<wave_grid((uint64_t)B * D, 256, 8), 256, 0, (cudaStream_t)stream>>>(grad, outputs, B, D, deg, C,
                                                                                            grad_inputs);
    return finish_launch("rn_freq_encode_backward");
}
