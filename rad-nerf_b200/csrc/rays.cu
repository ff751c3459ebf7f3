// rays.cu -- camera-ray generation on the device (absorbs the reference's per-frame get_rays torch-op chain,
// nerf/utils.py:248-333: meshgrid + 0.5, (i-cx)/fx, normalise, directions @ R^T, origin broadcast).
#include "common.cuh"

namespace rn {
namespace {

__global__ void __launch_bounds__(256)
get_rays_kernel(const float* __restrict__ pose, float fx, float fy, float cx, float cy, uint32_t W,
                const int32_t* __restrict__ pixel_ids, uint32_t n, float* __restrict__ rays_o, float* __restrict__ rays_d) {
    __shared__ float P[12];
    if (threadIdx.x < 12) P[threadIdx.x] = pose[threadIdx.x];  // rows 0..2 of the 4x4
    __syncthreads();
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t p = pixel_ids ? (uint32_t)__ldg(pixel_ids + k) : k;
        const uint32_t j = p / W, i = p - j * W;
        const float x = ((float)i + 0.5f - cx) / fx;
        const float y = ((float)j + 0.5f - cy) / fy;
        const float inv = 1.0f / sqrtf(x * x + y * y + 1.0f);
        const float dx = x * inv, dy = y * inv, dz = inv;
        float* o = rays_o + (size_t)k * 3;
        float* d = rays_d + (size_t)k * 3;
        o[0] = P[3]; o[1] = P[7]; o[2] = P[11];
        d[0] = P[0] * dx + P[1] * dy + P[2] * dz;
        d[1] = P[4] * dx + P[5] * dy + P[6] * dz;
        d[2] = P[8] * dx + P[9] * dy + P[10] * dz;
    }
}

// fp32 image in [0,1] -> uint8, 16 values per thread (one 16-byte store): the reference does this on the host after the
// device->host copy, `(pred * 255).astype(np.uint8)` (nerf/utils.py:952-960) -- fp32 multiply, truncation toward zero
__global__ void __launch_bounds__(256)
image_to_uint8_kernel(const float4* __restrict__ src, uint4* __restrict__ dst, uint32_t n16) {
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += gridDim.x * blockDim.x) {
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 v = __ldg(src + 4 * k + q);
            w[q] = (uint32_t)(uint8_t)__fmul_rn(v.x, 255.0f) | ((uint32_t)(uint8_t)__fmul_rn(v.y, 255.0f) << 8) |
                   ((uint32_t)(uint8_t)__fmul_rn(v.z, 255.0f) << 16) | ((uint32_t)(uint8_t)__fmul_rn(v.w, 255.0f) << 24);
        }
        dst[k] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" int rn_image_to_uint8(const float* image, uint8_t* out, uint64_t n_values, void* stream) {
    if (n_values == 0) return RN_OK;
    RN_REQUIRE(image && out, "null pointer");
    RN_REQUIRE(n_values % 16 == 0 && n_values < (1ull << 35) && ((uintptr_t)image & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "n_values must be a multiple of 16 and the buffers 16-byte aligned");
    const uint32_t n16 = (uint32_t)(n_values / 16);
    image_to_uint8_kernel<<<wave_grid(n16, 256, 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(image),
                                                                                  reinterpret_cast<uint4*>(out), n16);
    return finish_launch("rn_image_to_uint8");
}

extern "C" int rn_get_rays(const float* pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                           const int32_t* pixel_ids, uint32_t n, float* rays_o, float* rays_d, void* stream) {
    if (n == 0) return RN_OK;
    RN_REQUIRE(pose && rays_o && rays_d, "null pointer");
    RN_REQUIRE(pixel_ids || n == H * W, "n must be H*W when no pixel list is given");
    get_rays_kernel<<<wave_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(pose, fx, fy, cx, cy, W, pixel_ids, n, rays_o, rays_d);
    return finish_launch("rn_get_rays");
}
