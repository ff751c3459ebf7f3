// common.cuh -- shared host/device helpers for libradnerf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <atomic>

#include "../../include/radnerf_b200.h"

#ifndef RN_NUM_SMS
#define RN_NUM_SMS 148  // B200: 2 dies x 74 SMs
#endif

namespace rn {

// ---- error plumbing (thread-local message, see rn_last_error_string) -------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

inline int finish_launch(const char* what) {
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear sticky launch-config errors
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return RN_OK;
}

#define RN_REQUIRE(cond, msg)                              \
    do {                                                   \
        if (!(cond)) {                                     \
            rn::set_error("%s: %s", __func__, msg);        \
            return RN_E_BADARG;                            \
        }                                                  \
    } while (0)

template <typename T>
__host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

// Grid sizing: enough CTAs for `work` items at `per_cta` each, but when a grid-stride loop is used we cap
// at a whole number of waves over the 148 SMs.
inline uint32_t wave_grid(uint64_t work, uint32_t per_cta, uint32_t ctas_per_sm) {
    uint64_t need = (work + per_cta - 1) / per_cta;
    uint64_t cap = (uint64_t)RN_NUM_SMS * ctas_per_sm;
    if (need < 1) need = 1;
    return (uint32_t)(need < cap ? need : cap);
}

// ---- bit tricks shared by the occupancy-grid kernels (reference: raymarching.cu:56-81) -------------
// 10-bit -> 30-bit spread, two zero bits between consecutive input bits.
__host__ __device__ __forceinline__ uint32_t spread3(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__host__ __device__ __forceinline__ uint32_t morton_encode(uint32_t x, uint32_t y, uint32_t z) {
    return spread3(x) | (spread3(y) << 1) | (spread3(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t compact3(uint32_t x) {
    x &= 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return x;
}

}  // namespace rn
