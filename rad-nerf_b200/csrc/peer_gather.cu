// peer_gather.cu -- the image all-gather of the ray-sharded frame as direct stores into peer memory over NVLink/NVSwitch.
//
// Each rank renders the image rows of its tiles (interleaved 8-row tiles, radnerf_b200/sharding.py).  Instead of an NCCL
// all-gather into a rank-major buffer followed by an un-permute gather kernel, every rank writes its finished rows straight
// to their final position in EVERY rank's full-frame buffer (buffers come from torch symmetric memory: the same
// allocation mapped into all processes; `peers` is the device array of their base addresses).  A symmetric-memory barrier
// afterwards makes the stores visible.  3 MB per rank per 512x512 frame in total, fully coalesced 16-byte stores
// (a tile is 8 rows x W pixels x 3 floats of contiguous memory).
#include "common.cuh"

namespace rn {
namespace {

// row_f4 = float4s per pixel-row chunk handled as a unit: ids index ROWS of `row_floats` floats (a pixel = 3 floats is not
// 16-byte sized, so the unit of the scatter is a run of `run` consecutive pixels whose first id is ids[k * run]).
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float4* __restrict__ local, const int32_t* __restrict__ ids, uint32_t n_runs, uint32_t run_f4,
                    uint32_t run_pixels, const uint64_t* __restrict__ peers, uint32_t world) {
    const uint32_t total = n_runs * run_f4;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint32_t r = e / run_f4, c = e - r * run_f4;
        const uint32_t first_pixel = (uint32_t)__ldg(ids + (size_t)r * run_pixels);
        const float4 v = __ldg(local + e);
        const size_t dst = ((size_t)first_pixel * 3) / 4 + c;   // run starts are 16-byte aligned (checked on the host)
        for (uint32_t p = 0; p < world; ++p) reinterpret_cast<float4*>(peers[p])[dst] = v;
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

// local [n_local, 3] fp32 rows of this rank, ids [n_local] their pixel indices in the full frame, made of runs of
// `run_pixels` consecutive pixels (run_pixels * 3 floats must be a multiple of 4 and every run start * 3 too);
// peers: DEVICE array of `world` base addresses of the full-frame [H*W, 3] fp32 buffers (own rank included).
extern "C" int rn_scatter_rows_to_peers(const float* local, const int32_t* ids, uint32_t n_local, uint32_t run_pixels,
                                        const uint64_t* peers, uint32_t world, void* stream) {
    if (n_local == 0) return RN_OK;
    RN_REQUIRE(local && ids && peers && world >= 1, "null pointer");
    RN_REQUIRE(run_pixels >= 1 && n_local % run_pixels == 0 && (run_pixels * 3) % 4 == 0, "run_pixels must divide n_local and span whole float4s");
    RN_REQUIRE(((uintptr_t)local & 15) == 0, "local must be 16-byte aligned");
    const uint32_t run_f4 = run_pixels * 3 / 4, n_runs = n_local / run_pixels;
    scatter_rows_kernel<<<wave_grid(n_runs * run_f4, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(local), ids, n_runs, run_f4, run_pixels, peers, world);
    return finish_launch("rn_scatter_rows_to_peers");
}
