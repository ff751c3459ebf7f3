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


// ---- gather-to-root with arrival / consumed flags: no cross-rank barrier, no host call between render and delivery ------------------
// Only ONE rank (the root) hands frames to the host, so only its frame buffer needs the other ranks' rows.  Frame buffers are used
// round-robin (one per frame lane); `seq` = 1, 2, ... counts the frames that went through a buffer, identically on every rank.
//   non-root rank: wait until the root has CONSUMED frame seq-1 of this buffer (`consumed` lives in this rank's memory, the root writes
//                  it), store its rows into the root's buffer, then add 1 per CTA to the root's `arrived` counter (release, system scope);
//   root:          (its own rows go into its buffer by the same kernel, without signalling -- stream order covers them) the staging
//                  kernel waits until `arrived` has reached seq * (world-1) * scatter CTAs, copies / converts the assembled frame to
//                  the staging buffer, and its last CTA writes seq into every other rank's `consumed` flag.
// Control words are 64-bit, in a symmetric allocation: ctrl[2k] = arrived, ctrl[2k+1] = consumed of frame buffer k.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_release_sys_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// bounded spin (a protocol bug or a dead peer traps instead of hanging the GPU): ~2 s at 2 GHz
__device__ __forceinline__ void spin_until(const unsigned long long* p, unsigned long long want) {
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(64);
        if (clock64() - t0 > (4ll << 30)) __trap();
    }
}

__global__ void __launch_bounds__(256)
scatter_rows_signal_kernel(const float4* __restrict__ local, const int32_t* __restrict__ ids, uint32_t n_runs, uint32_t run_f4,
                           uint32_t run_pixels, const uint64_t* __restrict__ frame_peers, const uint64_t* __restrict__ ctrl_peers, uint32_t rank,
                           uint32_t root, uint32_t slot, unsigned long long seq) {
    float4* __restrict__ dst_frame = reinterpret_cast<float4*>(frame_peers[root]);
    if (rank != root) {   // the root must have staged the previous frame of this buffer before it is overwritten
        if (threadIdx.x == 0) spin_until(reinterpret_cast<const unsigned long long*>(ctrl_peers[rank]) + 2 * slot + 1, seq - 1);
        __syncthreads();
    }
    const uint32_t total = n_runs * run_f4;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint32_t r = e / run_f4, c = e - r * run_f4;
        const uint32_t first_pixel = (uint32_t)__ldg(ids + (size_t)r * run_pixels);
        dst_frame[((size_t)first_pixel * 3) / 4 + c] = __ldg(local + e);
    }
    if (rank != root) {
        __threadfence_system();      // every thread: its stores are ordered before the signal below
        __syncthreads();
        if (threadIdx.x == 0) red_release_sys_add(reinterpret_cast<unsigned long long*>(ctrl_peers[root]) + 2 * slot, 1ull);
    }
}

__global__ void __launch_bounds__(256)
stage_frame_kernel(const float4* __restrict__ frame, void* __restrict__ dst, uint32_t n_f4, uint32_t to_uint8,
                   unsigned long long want_arrived, const uint64_t* __restrict__ ctrl_peers, uint32_t world, uint32_t root, uint32_t slot,
                   unsigned long long seq, uint32_t* ticket) {
    if (ctrl_peers) {
        if (threadIdx.x == 0) spin_until(reinterpret_cast<const unsigned long long*>(ctrl_peers[root]) + 2 * slot, want_arrived);
        __syncthreads();
    }
    if (dst) {
        // __ldcg: the frame was (partly) written by other GPUs through NVLink into this GPU's L2 -- never read it through L1
        if (to_uint8) {
            const uint32_t n16 = n_f4 / 4;
            uint4* out = reinterpret_cast<uint4*>(dst);
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n16; k += gridDim.x * blockDim.x) {
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = __ldcg(frame + 4 * k + q);
                    w[q] = (uint32_t)(uint8_t)__fmul_rn(v.x, 255.0f) | ((uint32_t)(uint8_t)__fmul_rn(v.y, 255.0f) << 8) |
                           ((uint32_t)(uint8_t)__fmul_rn(v.z, 255.0f) << 16) | ((uint32_t)(uint8_t)__fmul_rn(v.w, 255.0f) << 24);
                }
                out[k] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
            float4* out = reinterpret_cast<float4*>(dst);
            for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_f4; k += gridDim.x * blockDim.x) out[k] = __ldcg(frame + k);
        }
    }
    if (ctrl_peers) {   // the last CTA to finish tells every other rank that this buffer may be overwritten
        __shared__ uint32_t last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
        __syncthreads();
        if (last) {
            if (threadIdx.x == 0) *ticket = 0u;
            for (uint32_t p = threadIdx.x; p < world; p += blockDim.x)
                if (p != root) st_release_sys(reinterpret_cast<unsigned long long*>(ctrl_peers[p]) + 2 * slot + 1, seq);
        }
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

// grid of the scatter kernel: a pure function of the row count, so that every rank (and the root's expected arrival count) agrees
static uint32_t scatter_grid(uint32_t n_local, uint32_t run_pixels) { return wave_grid((uint64_t)(n_local / run_pixels) * (run_pixels * 3 / 4), 256, 2); }

extern "C" uint32_t rn_scatter_signal_ctas(uint32_t n_local, uint32_t run_pixels) {
    if (run_pixels == 0 || n_local % run_pixels || (run_pixels * 3) % 4) return 0;
    return scatter_grid(n_local, run_pixels);
}

extern "C" int rn_scatter_rows_to_root(const float* local, const int32_t* ids, uint32_t n_local, uint32_t run_pixels, const uint64_t* frame_peers,
                                       const uint64_t* ctrl_peers, uint32_t world, uint32_t rank, uint32_t root, uint32_t slot, uint64_t seq,
                                       void* stream) {
    if (n_local == 0) return RN_OK;
    RN_REQUIRE(local && ids && frame_peers && ctrl_peers && world >= 1 && rank < world && root < world && seq >= 1, "bad arguments");
    RN_REQUIRE(run_pixels >= 1 && n_local % run_pixels == 0 && (run_pixels * 3) % 4 == 0, "run_pixels must divide n_local and span whole float4s");
    RN_REQUIRE(((uintptr_t)local & 15) == 0, "local must be 16-byte aligned");
    const uint32_t run_f4 = run_pixels * 3 / 4, n_runs = n_local / run_pixels;
    scatter_rows_signal_kernel<<<scatter_grid(n_local, run_pixels), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(local), ids, n_runs, run_f4, run_pixels, frame_peers, ctrl_peers, rank, root, slot, (unsigned long long)seq);
    return finish_launch("rn_scatter_rows_to_root");
}

extern "C" int rn_stage_frame_at_root(const float* frame, void* dst, uint64_t n_values, uint32_t to_uint8, const uint64_t* ctrl_peers, uint32_t world,
                                      uint32_t root, uint32_t slot, uint64_t seq, uint32_t scatter_ctas, uint32_t* ticket, void* stream) {
    RN_REQUIRE(frame, "null frame");
    RN_REQUIRE(n_values % 16 == 0 && n_values < (1ull << 34) && ((uintptr_t)frame & 15) == 0 && ((uintptr_t)dst & 15) == 0,
               "n_values must be a multiple of 16 and the buffers 16-byte aligned");
    RN_REQUIRE(ctrl_peers == nullptr || (world >= 2 && root < world && seq >= 1 && scatter_ctas >= 1 && ticket), "bad exchange arguments");
    if (!dst && !ctrl_peers) return RN_OK;
    const uint32_t n_f4 = (uint32_t)(n_values / 4);
    const unsigned long long want = ctrl_peers ? (unsigned long long)seq * (world - 1) * scatter_ctas : 0ull;
    // a small grid: its CTAs may spin on the arrival counter while other frame lanes need the SMs
    stage_frame_kernel<<<dst ? wave_grid(n_f4, 256 * 8, 1) : 1, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(frame), dst, n_f4, to_uint8, want, ctrl_peers, world, root, slot, (unsigned long long)seq, ticket);
    return finish_launch("rn_stage_frame_at_root");
}
