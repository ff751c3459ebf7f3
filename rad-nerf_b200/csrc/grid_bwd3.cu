// grid_bwd3.cu -- table-gradient scatter of the 3-D, 2-feature grid (the head's spatial encoder), restructured.
//
// Replaces kernel_grid_backward (gridencoder/src/gridencoder.cu:247-339) for D = 3, C = 2, linear interpolation,
// align_corners = false, fp32 accumulation target -- the only 3-D configuration RAD-NeRF trains (nerf/network.py:134).
// The generic kernel (gridencoder_impl.cuh) issues 16 levels x 8 corners = 128 `red.global.add.v2.f32` per sample; ncu shows
// it bound by the L2 atomic units (profiles/r02_grid_backward3_ncu_full.json).  Three structural facts cut that traffic:
//
//   1. tiled levels whose running stride exceeds the level size never index z (the reference's early exit,
//      gridencoder.cu:72): their two z-corners are the SAME row, and w(.., z0) + w(.., z1) = w(..): 4 rows instead of 8 on
//      7 of the 16 levels;
//   2. the two x-corners of a pair are adjacent rows (stride 1): when the first is even they form ONE 16-byte
//      `red.global.add.v4.f32` (one L2 transaction instead of two);
//   3. consecutive samples of a training batch are consecutive steps of one ray (march_rays_train packs rays), 0.027 apart:
//      on the coarse levels several of them sit in the same cell.  A segmented warp reduction (runs of equal cells in
//      consecutive lanes, shuffles) sums them before ONE lane issues the atomics.
//
// VARIANT selects the pieces (experiments: tools/bwd3_experiment.py); the product entry point uses the winner.
#include "gridencoder_impl.cuh"

namespace rn {
namespace grid {
namespace {

constexpr uint32_t FULL = 0xffffffffu;

__device__ __forceinline__ void emit_pair(float* __restrict__ gt, uint32_t r0, uint32_t r1, float a0, float a1, float b0, float b1) {
    if (r1 == r0 + 1u && (r0 & 1u) == 0u) {
        red_add4(gt + 2 * (size_t)r0, a0, a1, b0, b1);
    } else {
        red_add2(gt + 2 * (size_t)r0, a0, a1);
        red_add2(gt + 2 * (size_t)r1, b0, b1);
    }
}

struct Geom {   // one level, D = 3
    float scale;
    uint32_t s1, s2, mask_mode, size, offset, res;   // mask_mode: 0 dense, 1 pow2 mask, 2 generic modulo, |4 hashed
};

__device__ __forceinline__ uint32_t row_of(const Geom& g, uint32_t x, uint32_t y, uint32_t z) {
    uint32_t idx;
    if (g.mask_mode & 4u) {
        const uint32_t p[3] = {x, y, z};
        idx = lattice_hash<3>(p);
    } else {
        idx = x + y * g.s1 + z * g.s2;
    }
    const uint32_t w = g.mask_mode & 3u;
    if (w == 1u) idx &= g.size - 1u;
    else if (w == 2u) idx %= g.size;
    return idx;
}

// VARIANT bit 0: z-merge + x-pair v4;  bit 1: segmented warp aggregation on levels < agg_levels;
// bit 2: dense leading levels privatised in shared memory (priv_levels of them, priv_rows rows in total)
template <typename T, int VARIANT>
__global__ void __launch_bounds__(VARIANT & 4 ? 1024 : 256)
grid_backward3_kernel(const T* __restrict__ grad, const float* __restrict__ inputs, const int32_t* __restrict__ offsets,
                      float* __restrict__ grad_table, uint32_t B, uint32_t L, float S, uint32_t H, uint32_t gridtype,
                      uint32_t level_mask, uint32_t agg_levels, uint32_t priv_levels, uint32_t priv_rows) {
    __shared__ Geom geom[MAX_LEVELS];
    extern __shared__ __align__(16) float s_priv[];   // [priv_rows][2]
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x) {
        LevelMeta m;
        make_level_meta(m, l, offsets, S, H, 3, gridtype, false);
        Geom g;
        g.scale = m.scale; g.s1 = m.stride[1]; g.s2 = m.stride[2]; g.size = m.size; g.offset = m.offset; g.res = m.resolution;
        g.mask_mode = (m.mode >> 1) | ((m.mode & 1u) << 2);
        geom[l] = g;
    }
    if constexpr (VARIANT & 4) {
        for (uint32_t i = threadIdx.x; i < priv_rows * 2; i += blockDim.x) s_priv[i] = 0.0f;
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;

    for (uint32_t base = blockIdx.x * blockDim.x; base < B; base += gridDim.x * blockDim.x) {
        const uint32_t b = base + threadIdx.x;
        bool valid = b < B;
        float x[3] = {0.f, 0.f, 0.f};
        if (valid) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                x[d] = __ldg(inputs + (size_t)b * 3 + d);
                if (x[d] < 0 || x[d] > 1) valid = false;
            }
        }
        uint32_t priv_base = 0;
#pragma unroll 1
        for (uint32_t l = 0; l < L; ++l) {
            const Geom g = geom[l];
            if (!((level_mask >> l) & 1u)) { if ((VARIANT & 4) && l < priv_levels) priv_base += g.size; continue; }
            float g0 = 0.f, g1 = 0.f;
            if (b < B) {
                const Row<T, 2> gr = load_row<T, 2>(grad + (size_t)b * L * 2 + (size_t)l * 2);
                g0 = to_f(gr.v[0]); g1 = to_f(gr.v[1]);
            }
            uint32_t pg[3];
            float fr[3];
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float pos = __fmaf_rn(x[d], g.scale, 0.5f);
                const float fl = floorf(pos);
                pg[d] = (uint32_t)fl;
                fr[d] = pos - (float)pg[d];
            }
            float* __restrict__ gt = grad_table + (size_t)g.offset * 2;
            const float x0 = 1.0f - fr[0], x1 = fr[0], y0 = 1.0f - fr[1], y1 = fr[1], z0 = 1.0f - fr[2], z1 = fr[2];
            const float w00 = x0 * y0, w10 = x1 * y0, w01 = x0 * y1, w11 = x1 * y1;

            if constexpr ((VARIANT & 1) == 0) {   // plain: 8 vector atomics, the generic kernel's traffic
                if (valid) {
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) {
                        const float w = ((k & 1u) ? x1 : x0) * ((k & 2u) ? y1 : y0) * ((k & 4u) ? z1 : z0);
                        red_add2(gt + 2 * (size_t)row_of(g, pg[0] + (k & 1u), pg[1] + ((k >> 1) & 1u), pg[2] + (k >> 2)), w * g0, w * g1);
                    }
                }
                continue;
            }
            const bool zdrop = g.s2 == 0u && !(g.mask_mode & 4u);
            if (zdrop) {   // warp-uniform: the level ignores z, 4 distinct rows
                if (valid) {
                    const uint32_t r00 = row_of(g, pg[0], pg[1], 0), r10 = row_of(g, pg[0] + 1, pg[1], 0);
                    const uint32_t r01 = row_of(g, pg[0], pg[1] + 1, 0), r11 = row_of(g, pg[0] + 1, pg[1] + 1, 0);
                    emit_pair(gt, r00, r10, w00 * g0, w00 * g1, w10 * g0, w10 * g1);
                    emit_pair(gt, r01, r11, w01 * g0, w01 * g1, w11 * g0, w11 * g1);
                }
                continue;
            }
            // 8 corners: v[2*k + c], k = x + 2y + 4z
            float v[16];
            {
                const float w[8] = {w00 * z0, w10 * z0, w01 * z0, w11 * z0, w00 * z1, w10 * z1, w01 * z1, w11 * z1};
#pragma unroll
                for (int k = 0; k < 8; ++k) { v[2 * k] = valid ? w[k] * g0 : 0.f; v[2 * k + 1] = valid ? w[k] * g1 : 0.f; }
            }
            bool emit = valid;
            if constexpr (VARIANT & 2) {
                if (l < agg_levels && g.res < 1023u) {   // warp-uniform
                    const uint32_t key = valid ? (pg[0] | (pg[1] << 10) | (pg[2] << 20)) : FULL;
                    const uint32_t prev = __shfl_up_sync(FULL, key, 1);
                    const bool head = lane == 0u || key != prev;
                    const uint32_t heads = __ballot_sync(FULL, head);
                    if (heads != FULL) {   // at least one run of two lanes in the same cell
                        const uint32_t above = heads & ~((2u << lane) - 1u);   // heads strictly above this lane (lane 31: shift wraps to 0 -> ~(-1) = 0)
                        const uint32_t run_end = (lane == 31u || above == 0u) ? 31u : (uint32_t)(__ffs(above) - 2);
#pragma unroll
                        for (uint32_t off = 1; off < 32; off <<= 1) {
                            const bool take = lane + off <= run_end;
                            if (!__any_sync(FULL, take)) break;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                const float t = __shfl_down_sync(FULL, v[j], off);
                                if (take) v[j] += t;
                            }
                        }
                        emit = valid && head;
                    }
                }
            }
            if (emit) {
                const uint32_t r[8] = {row_of(g, pg[0], pg[1], pg[2]),         row_of(g, pg[0] + 1, pg[1], pg[2]),
                                       row_of(g, pg[0], pg[1] + 1, pg[2]),     row_of(g, pg[0] + 1, pg[1] + 1, pg[2]),
                                       row_of(g, pg[0], pg[1], pg[2] + 1),     row_of(g, pg[0] + 1, pg[1], pg[2] + 1),
                                       row_of(g, pg[0], pg[1] + 1, pg[2] + 1), row_of(g, pg[0] + 1, pg[1] + 1, pg[2] + 1)};
                if ((VARIANT & 4) && l < priv_levels) {
                    float* sp = s_priv + 2 * (size_t)priv_base;
#pragma unroll
                    for (int k = 0; k < 8; ++k) { atomicAdd(sp + 2 * r[k], v[2 * k]); atomicAdd(sp + 2 * r[k] + 1, v[2 * k + 1]); }
                } else {
#pragma unroll
                    for (int k = 0; k < 8; k += 2) emit_pair(gt, r[k], r[k + 1], v[2 * k], v[2 * k + 1], v[2 * k + 2], v[2 * k + 3]);
                }
            }
            if ((VARIANT & 4) && l < priv_levels) priv_base += g.size;
        }
    }
    if constexpr (VARIANT & 4) {
        __syncthreads();
        // privatised levels are the leading ones: their rows are contiguous in the table from offset 0
        for (uint32_t i = threadIdx.x; i < priv_rows; i += blockDim.x) {
            const float a = s_priv[2 * i], c = s_priv[2 * i + 1];
            if (a != 0.0f || c != 0.0f) red_add2(grad_table + 2 * (size_t)i, a, c);
        }
    }
}

template <typename T, int VARIANT>
int launch_bwd3(const void* grad, const float* inputs, const int32_t* offsets, float* gt, uint32_t B, uint32_t L, float S, uint32_t H,
                uint32_t gridtype, uint32_t level_mask, uint32_t agg_levels, uint32_t priv_levels, uint32_t priv_rows, cudaStream_t st) {
    if constexpr (VARIANT & 4) {
        const size_t smem = (size_t)priv_rows * 8;
        cudaError_t e = cudaFuncSetAttribute(grid_backward3_kernel<T, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("grid_backward3: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
        grid_backward3_kernel<T, VARIANT><<<RN_NUM_SMS, 1024, smem, st>>>((const T*)grad, inputs, offsets, gt, B, L, S, H, gridtype, level_mask,
                                                                         agg_levels, priv_levels, priv_rows);
    } else {
        grid_backward3_kernel<T, VARIANT><<<wave_grid(B, 256, 8), 256, 0, st>>>((const T*)grad, inputs, offsets, gt, B, L, S, H, gridtype,
                                                                               level_mask, agg_levels, 0, 0);
    }
    return finish_launch("rn_grid_backward3");
}

}  // namespace
}  // namespace grid
}  // namespace rn

using namespace rn;
using namespace rn::grid;

// Experimental entry point (tools/bwd3_experiment.py, tests): table gradient of a 3-D, 2-feature, linearly interpolated grid.
// grad [B, L*2] (RN_LAYOUT_BLC) of `dtype`, inputs [B,3] in [0,1], grad_table fp32 [rows,2] accumulated into.
// variant: bit0 z-merge + x-pair, bit1 segmented warp aggregation below `agg_levels`, bit2 shared-memory privatisation of the
// first `priv_levels` levels (`priv_rows` = their total row count; they must be the dense leading levels).
// level_mask: bit l = process level l.
extern "C" int rn_grid_backward3(const void* grad, const float* inputs, const int32_t* offsets, float* grad_table, uint32_t B, uint32_t L,
                                 float S, uint32_t H, uint32_t gridtype, uint32_t dtype, uint32_t variant, uint32_t level_mask,
                                 uint32_t agg_levels, uint32_t priv_levels, uint32_t priv_rows, void* stream) {
    RN_REQUIRE(L >= 1 && L <= 32, "num_levels must be in [1, 32]");
    RN_REQUIRE(dtype <= 1 && gridtype <= 1 && variant < 8, "bad enum argument");
    RN_REQUIRE(!(variant & 4) || (priv_rows > 0 && priv_rows * 8 <= 220 * 1024), "privatised rows must fit in shared memory");
    if (B == 0) return RN_OK;
    RN_REQUIRE(grad && inputs && offsets && grad_table, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define RN_BWD3(V)                                                                                                                       \
    case V:                                                                                                                              \
        return dtype == RN_F16 ? launch_bwd3<__half, V>(grad, inputs, offsets, grad_table, B, L, S, H, gridtype, level_mask, agg_levels, \
                                                        priv_levels, priv_rows, st)                                                      \
                               : launch_bwd3<float, V>(grad, inputs, offsets, grad_table, B, L, S, H, gridtype, level_mask, agg_levels,  \
                                                       priv_levels, priv_rows, st);
    switch (variant) {
        RN_BWD3(0) RN_BWD3(1) RN_BWD3(2) RN_BWD3(3) RN_BWD3(4) RN_BWD3(5) RN_BWD3(6) RN_BWD3(7)
    }
#undef RN_BWD3
    return RN_E_UNSUPPORTED;
}
