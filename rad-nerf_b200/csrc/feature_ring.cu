// feature_ring.cu -- streaming audio hand-off: the renderer's [8, dim, 16] attention window of one
// video frame gathered from the ASR feature ring in ONE launch, written where the frame graph reads it.
//
// Reference: ASR.get_next_feat (nerf/asr.py:160-183) keeps the 8 latest 16-row windows of `feat_queue [size, dim]` as a
// Python list of slice+permute views (or torch.cat copies when a window wraps around the ring end) and stacks them per frame:
// 8 slices, 8 permutes, a stack and, downstream, a device copy into the model input.  Here the host keeps 8 window
// descriptors; window w is rows (start_w + j) mod size, j = 0..15, read LIVE from the ring (the reference's views see later
// overwrites of the ring, so must we), or from a snapshot slot taken when the window was created (the reference's torch.cat
// copies), or all zeros (the four start-up windows, asr.py:109).
#include "common.cuh"

namespace rn {
namespace {

constexpr int kWin = RN_RING_WINDOW;     // 16 rows per window
constexpr int kDepth = RN_RING_DEPTH;    // 8 windows per frame

__global__ void __launch_bounds__(256)
feature_window_kernel(const float* __restrict__ ring, uint32_t size, uint32_t dim, rn_ring_windows w,
                      float* __restrict__ snapshots, float* __restrict__ out) {
    // 1) windows created by this call that wrap: copy their 16 rows into their snapshot slot (row-major [16, dim])
    const uint32_t per_win = kWin * dim;
    for (int k = 0; k < kDepth; ++k) {
        if (w.snapshot[k] >= 0 && w.fresh[k]) {
            float* dst = snapshots + (size_t)w.snapshot[k] * per_win;
            for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < per_win; e += gridDim.x * blockDim.x) {
                const uint32_t j = e / dim, d = e - j * dim;
                dst[e] = __ldg(ring + (size_t)((w.start[k] + j) % size) * dim + d);
            }
        }
    }
    // 2) the frame's block, [8, dim, 16]: element (k, d, j) = row j of window k, column d.  A fresh snapshot is read from the
    //    ring directly (identical values; the slot is being written by other threads of this launch)
    const uint32_t total = kDepth * per_win;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint32_t k = e / per_win, r = e - k * per_win;
        const uint32_t d = r / kWin, j = r - d * kWin;
        float v = 0.0f;
        if (w.start[k] >= 0) {
            if (w.snapshot[k] >= 0 && !w.fresh[k]) v = snapshots[(size_t)w.snapshot[k] * per_win + j * dim + d];
            else v = __ldg(ring + (size_t)((w.start[k] + j) % size) * dim + d);
        }
        out[e] = v;
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" int rn_feature_window(const float* ring, uint32_t size, uint32_t dim, const rn_ring_windows* windows,
                                 float* snapshots, float* out, void* stream) {
    RN_REQUIRE(ring && windows && out, "null pointer");
    RN_REQUIRE(size >= RN_RING_WINDOW && dim >= 1, "the ring must hold at least one window");
    for (int k = 0; k < RN_RING_DEPTH; ++k) {
        RN_REQUIRE(windows->start[k] < (int32_t)size, "window start outside the ring");
        RN_REQUIRE(windows->snapshot[k] < RN_RING_DEPTH, "snapshot slot outside [0, RN_RING_DEPTH)");
        RN_REQUIRE(windows->snapshot[k] < 0 || snapshots, "a snapshot window needs the snapshot buffer");
    }
    const uint32_t total = RN_RING_DEPTH * RN_RING_WINDOW * dim;
    feature_window_kernel<<<div_up(total, 256u), 256, 0, (cudaStream_t)stream>>>(ring, size, dim, *windows, snapshots, out);
    return finish_launch("rn_feature_window");
}
