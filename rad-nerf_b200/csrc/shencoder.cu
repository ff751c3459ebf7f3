// shencoder.cu -- real spherical-harmonics direction encoding (degree 1..8) for sm_100a.
//
// Replaces the reference's shencoder extension (shencoder/src/shencoder.cu: kernel_sh :27-356,
// kernel_sh_backward :358-383).  The basis polynomials are generated from the closed form by tools/gen_sh.py
// (sh_basis.inc) instead of being hand-listed; channel order and sign convention are the reference's.
// One thread evaluates one direction; the degree is a template parameter so the row lives in registers and is
// written with 16-byte stores when its length allows.
#include "common.cuh"

#include "sh.cuh"

namespace rn {
namespace {

template <int N>
__device__ __forceinline__ void store_row(float* __restrict__ dst, const float (&v)[N]) {
    if constexpr (N % 4 == 0) {
        if (((uintptr_t)dst & 15u) == 0) {
#pragma unroll
            for (int i = 0; i < N / 4; ++i)
                reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) dst[i] = v[i];
}

template <int DEG>
__global__ void __launch_bounds__(256)
sh_forward_kernel(const float* __restrict__ inputs, float* __restrict__ outputs, uint32_t B, uint32_t D,
                  float* __restrict__ dy_dx) {
    constexpr int C2 = DEG * DEG;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float x = __ldg(inputs + (size_t)b * D), y = __ldg(inputs + (size_t)b * D + 1),
                    z = __ldg(inputs + (size_t)b * D + 2);
        float Y[C2];
        if (dy_dx) {
            float gx[C2], gy[C2], gz[C2];
            sh_eval<DEG, true>(x, y, z, Y, gx, gy, gz);
            float* g = dy_dx + (size_t)b * D * C2;
            store_row<C2>(g, gx);
            store_row<C2>(g + C2, gy);
            store_row<C2>(g + 2 * C2, gz);
        } else {
            sh_eval<DEG, false>(x, y, z, Y, nullptr, nullptr, nullptr);
        }
        store_row<C2>(outputs + (size_t)b * C2, Y);
    }
}

// grad_inputs[b,d] += sum_ch grad[b,ch] * dy_dx[b,d,ch]                   (shencoder.cu:358-383)
__global__ void __launch_bounds__(256)
sh_backward_kernel(const float* __restrict__ grad, uint32_t B, uint32_t D, uint32_t C2,
                   const float* __restrict__ dy_dx, float* __restrict__ grad_inputs) {
    const uint32_t total = B * D;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t b = t / D;
        const float* g = grad + (size_t)b * C2;
        const float* j = dy_dx + (size_t)t * C2;
        float acc = grad_inputs[t];
        for (uint32_t ch = 0; ch < C2; ++ch) acc = __fmaf_rn(__ldg(g + ch), __ldg(j + ch), acc);
        grad_inputs[t] = acc;
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" int rn_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree,
                                    float* dy_dx, void* stream) {
    RN_REQUIRE(D == 3, "SH encoder only supports input dim == 3");
    if (degree < 1 || degree > 8) {
        set_error("rn_sh_encode_forward: SH encoder only supports degree in [1, 8] (got %u)", degree);
        return RN_E_UNSUPPORTED;
    }
    if (B == 0) return RN_OK;
    RN_REQUIRE(inputs && outputs, "null pointer");
    const uint32_t grid = wave_grid(B, 256, 8);
    cudaStream_t st = (cudaStream_t)stream;
    switch (degree) {
        case 1: sh_forward_kernel<1><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 2: sh_forward_kernel<2><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 3: sh_forward_kernel<3><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 4: sh_forward_kernel<4><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 5: sh_forward_kernel<5><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 6: sh_forward_kernel<6><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        case 7: sh_forward_kernel<7><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
        default: sh_forward_kernel<8><<<grid, 256, 0, st>>>(inputs, outputs, B, D, dy_dx); break;
    }
    return finish_launch("rn_sh_encode_forward");
}

extern "C" int rn_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree,
                                     const float* dy_dx, float* grad_inputs, void* stream) {
    (void)inputs;
    RN_REQUIRE(D == 3, "SH encoder only supports input dim == 3");
    RN_REQUIRE(degree >= 1 && degree <= 8, "degree must be in [1, 8]");
    if (B == 0) return RN_OK;
    RN_REQUIRE(grad && dy_dx && grad_inputs, "null pointer");
    sh_backward_kernel<<<wave_grid((uint64_t)B * D, 256, 8), 256, 0, (cudaStream_t)stream>>>(grad, B, D, degree * degree,
                                                                                          dy_dx, grad_inputs);
    return finish_launch("rn_sh_encode_backward");
}
