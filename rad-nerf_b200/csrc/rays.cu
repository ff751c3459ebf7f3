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

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" int rn_get_rays(const float* pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                           const int32_t* pixel_ids, uint32_t n, float* rays_o, float* rays_d, void* stream) {
    if (n == 0) return RN_OK;
    RN_REQUIRE(pose && rays_o && rays_d, "null pointer");
    RN_REQUIRE(pixel_ids || n == H * W, "n must be H*W when no pixel list is given");
    get_rays_kernel<<<wave_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(pose, fx, fy, cx, cy, W, pixel_ids, n, rays_o, rays_d);
    return finish_launch("rn_get_rays");
}
