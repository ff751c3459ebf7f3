// gridencoder.cu -- multi-resolution hash / tiled grid encoding for sm_100a.
//
// Replaces the reference's gridencoder extension (gridencoder/src/gridencoder.cu: kernel_grid :87-244,
// kernel_grid_backward :247-339, kernel_input_backward :342-368, kernel_grad_tv :505-609) behind the C ABI in
// include/radnerf_b200.h.  Not a port: the work decomposition is different.
//
//   reference : one thread per (sample, level), grid.y = level, scalar 2-byte loads, [L,B,C] output that the
//               Python wrapper permutes/copies to [B,L*C]; backward = one atomic per (sample, level, corner, pair).
//   here      : one thread per SAMPLE walking all levels in groups of 4 (32 independent gathers in flight per
//               thread, coordinates read once instead of L times).  All lanes of a warp sit on the same level at
//               the same time ("level-major"), so the lanes' gathers -- neighbouring samples of neighbouring
//               rays -- fall into the same few 128-byte lines.  One vector load per corner row (half2 / float2 /
//               float4 ...), per-level geometry computed once per CTA into shared memory, [B,L*C] rows written
//               directly with 16-byte stores (no permute pass).  Backward accumulates in fp32 with vector
//               red.global.add (v2/v4.f32) and fuses the input-gradient reduction.
//
// Numerics follow the reference exactly where it is deterministic: fp32 coordinates, FMA placement as nvcc
// contracts the reference expressions, and in fp16 mode the per-corner rounding of c10::Half arithmetic
// (product rounded to half, then a half add; gridencoder.cu:163,186) so forward results are bit-identical.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "umma.cuh"

namespace rn {
namespace grid {

constexpr int MAX_LEVELS = 64;
constexpr int LG = 4;  // levels per unrolled group

struct LevelMeta {
    float scale;          // exp2f(l*S)*H - 1                      (gridencoder.cu:138)
    uint32_t resolution;  // ceil(scale) + 1                       (gridencoder.cu:139)
    uint32_t offset;      // first row of the level
    uint32_t size;        // rows in the level ("hashmap_size", gridencoder.cu:137)
    uint32_t stride[5];   // per-dimension stride; 0 once the running stride exceeded `size` (the :72 early exit)
    uint32_t mode;        // bit0: index needs hashing; bits 1-2: 0 = no wrap needed, 1 = pow2 mask, 2 = generic %
};

__device__ __forceinline__ void make_level_meta(LevelMeta& m, uint32_t level, const int32_t* offsets, float S,
                                                uint32_t H, uint32_t D, uint32_t gridtype, bool align_corners) {
    const uint32_t off = (uint32_t)offsets[level];
    const uint32_t size = (uint32_t)offsets[level + 1] - off;
    const float scale = exp2f(level * S) * H - 1.0f;
    const uint32_t res = (uint32_t)ceilf(scale) + 1;
    m.scale = scale;
    m.resolution = res;
    m.offset = off;
    m.size = size;
    const uint32_t fac = align_corners ? res : res + 1;
    uint32_t stride = 1;
    bool stopped = false;
    uint64_t span = 1;  // number of distinct lattice points if every dimension were indexed
    for (uint32_t d = 0; d < 5; ++d) {
        if (d < D && !stopped && stride <= size) {
            m.stride[d] = stride;
            stride *= fac;  // uint32 wrap exactly as the reference
        } else {
            if (d < D) stopped = true;
            m.stride[d] = 0;
        }
        if (d < D) span = (span > (1ull << 40)) ? span : span * (uint64_t)(res + 1);
    }
    const bool overflow = stride > size;  // value of `stride > hashmap_size` after the loop (gridencoder.cu:79)
    uint32_t mode = 0;
    if (gridtype == 0 && overflow) mode |= 1u;
    uint32_t wrap;
    if (!(mode & 1u) && !align_corners && !stopped && span <= (uint64_t)size) wrap = 0;  // dense: index < size
    else if ((size & (size - 1)) == 0) wrap = 1;
    else wrap = 2;
    m.mode = mode | (wrap << 1);
}

// spatial hash of the reference (gridencoder.cu:50-63): xor of coordinate * prime.
template <int D>
__device__ __forceinline__ uint32_t lattice_hash(const uint32_t (&p)[D]) {
    constexpr uint32_t primes[7] = {1u, 2654435761u, 805459861u, 3674653429u, 2097192037u, 1434869437u, 2165219737u};
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < D; ++i) r ^= p[i] * primes[i];
    return r;
}

__device__ __forceinline__ uint32_t wrap_index(uint32_t idx, const LevelMeta& m) {
    const uint32_t w = m.mode >> 1;
    if (w == 0) return idx;
    if (w == 1) return idx & (m.size - 1);
    return idx % m.size;
}

template <int D>
__device__ __forceinline__ uint32_t corner_row(const LevelMeta& m, const uint32_t (&pg)[D], uint32_t corner) {
    uint32_t p[D];
#pragma unroll
    for (int d = 0; d < D; ++d) p[d] = pg[d] + ((corner >> d) & 1u);
    uint32_t idx;
    if (m.mode & 1u) {
        idx = lattice_hash<D>(p);
    } else {
        idx = 0;
#pragma unroll
        for (int d = 0; d < D; ++d) idx += p[d] * m.stride[d];
    }
    return wrap_index(idx, m);
}

// ---- feature-row access: one vector load per corner ---------------------------------------------------
template <typename T, int C>
struct Row {
    T v[C];
};

template <typename T, int C>
__device__ __forceinline__ Row<T, C> load_row(const T* __restrict__ p) {
    Row<T, C> r;
    constexpr int BYTES = C * sizeof(T);
    if constexpr (BYTES == 2) {
        *reinterpret_cast<unsigned short*>(r.v) = __ldg(reinterpret_cast<const unsigned short*>(p));
    } else if constexpr (BYTES == 4) {
        *reinterpret_cast<unsigned int*>(r.v) = __ldg(reinterpret_cast<const unsigned int*>(p));
    } else if constexpr (BYTES == 8) {
        *reinterpret_cast<uint2*>(r.v) = __ldg(reinterpret_cast<const uint2*>(p));
    } else if constexpr (BYTES == 16) {
        *reinterpret_cast<uint4*>(r.v) = __ldg(reinterpret_cast<const uint4*>(p));
    } else {  // 32 bytes
        reinterpret_cast<uint4*>(r.v)[0] = __ldg(reinterpret_cast<const uint4*>(p));
        reinterpret_cast<uint4*>(r.v)[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
    }
    return r;
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__half x) { return __half2float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __half from_f<__half>(float x) { return __float2half_rn(x); }

// acc += w * g with the reference's rounding: fp32 -> one FMA; fp16 -> round(w*g) to half, then half add.
__device__ __forceinline__ void accum(float& acc, float w, float g) { acc = __fmaf_rn(w, g, acc); }
__device__ __forceinline__ void accum(__half& acc, float w, __half g) {
    acc = __hadd(acc, __float2half_rn(__fmul_rn(w, __half2float(g))));
}
// dy_dx accumulation: acc += w * (gr - gl) * pd           (gridencoder.cu:234)
__device__ __forceinline__ void accum_diff(float& acc, float w, float gr, float gl, float pd) {
    acc = __fmaf_rn(__fmul_rn(w, __fsub_rn(gr, gl)), pd, acc);
}
__device__ __forceinline__ void accum_diff(__half& acc, float w, __half gr, __half gl, float pd) {
    const __half diff = __hsub(gr, gl);
    acc = __hadd(acc, __float2half_rn(__fmul_rn(__fmul_rn(w, __half2float(diff)), pd)));
}

__device__ __forceinline__ float smoothstep_f(float v) { return v * v * (3.0f - 2.0f * v); }
__device__ __forceinline__ float smoothstep_df(float v) { return 6 * v * (1.0f - v); }

template <int D>
struct Cell {
    float frac[D];   // interpolation weight along d (after smoothstep if enabled)
    float deriv[D];  // d(weight)/d(pos)
    uint32_t pg[D];  // integer lattice coordinate
};

template <int D>
__device__ __forceinline__ Cell<D> locate(const float (&x)[D], const LevelMeta& m, bool align_corners, uint32_t interp) {
    Cell<D> c;
    const float shift = align_corners ? 0.0f : 0.5f;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float pos = __fmaf_rn(x[d], m.scale, shift);
        const float fl = floorf(pos);
        c.pg[d] = (uint32_t)fl;
        pos -= (float)c.pg[d];
        if (interp == 1) {
            c.deriv[d] = smoothstep_df(pos);
            c.frac[d] = smoothstep_f(pos);
        } else {
            c.deriv[d] = 1.0f;
            c.frac[d] = pos;
        }
    }
    return c;
}

template <int D>
__device__ __forceinline__ float corner_weight(const Cell<D>& c, uint32_t corner) {
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < D; ++d) w *= ((corner >> d) & 1u) ? c.frac[d] : (1.0f - c.frac[d]);
    return w;
}

// ======================================================================================================
// forward
// ======================================================================================================
// STAGE: the north_star's "coarse levels in shared memory by TMA".  A persistent CTA copies the leading DENSE levels of the table
// (index < size without hashing or wrapping, so the level is one contiguous run of rows) into dynamic shared memory with one bulk copy
// per level, as many levels as `stage_bytes` holds, and gathers those levels with shared-memory loads; the other levels keep the
// read-only global path.  Same cells, same weights, same rounding: bit-identical outputs (tests/test_gpu_round2.py).
template <typename T, int C>
__device__ __forceinline__ Row<T, C> load_row_staged(const T* p) {
    Row<T, C> r;
    constexpr int BYTES = C * sizeof(T);
    if constexpr (BYTES == 2) *reinterpret_cast<unsigned short*>(r.v) = *reinterpret_cast<const unsigned short*>(p);
    else if constexpr (BYTES == 4) *reinterpret_cast<unsigned int*>(r.v) = *reinterpret_cast<const unsigned int*>(p);
    else if constexpr (BYTES == 8) *reinterpret_cast<uint2*>(r.v) = *reinterpret_cast<const uint2*>(p);
    else if constexpr (BYTES == 16) *reinterpret_cast<uint4*>(r.v) = *reinterpret_cast<const uint4*>(p);
    else {
        reinterpret_cast<uint4*>(r.v)[0] = reinterpret_cast<const uint4*>(p)[0];
        reinterpret_cast<uint4*>(r.v)[1] = reinterpret_cast<const uint4*>(p)[1];
    }
    return r;
}

constexpr uint32_t NOT_STAGED = 0xffffffffu;

template <typename T, int D, int C, bool VEC_OUT, bool STAGE = false>
__global__ void __launch_bounds__(STAGE ? 512 : 256)
grid_forward_kernel(const float* __restrict__ inputs, const T* __restrict__ table, const int32_t* __restrict__ offsets,
                    T* __restrict__ outputs, T* __restrict__ dy_dx, uint32_t B, uint32_t L, float S, uint32_t H,
                    uint32_t gridtype, uint32_t align_corners, uint32_t interp, uint32_t layout, uint32_t stage_bytes = 0) {
    __shared__ LevelMeta meta[MAX_LEVELS];
    __shared__ uint32_t staged_at[STAGE ? MAX_LEVELS : 1];   // byte offset of a staged level in the shared copy, or NOT_STAGED
    __shared__ alignas(8) uint64_t stage_bar;
    extern __shared__ __align__(128) unsigned char staged_rows[];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x)
        make_level_meta(meta[l], l, offsets, S, H, D, gridtype, align_corners != 0);
    __syncthreads();
    if constexpr (STAGE) {
        if (threadIdx.x == 0) {
            umma::mbar_init(&stage_bar, 1);
            umma::fence_mbar_init();
            uint32_t used = 0;
            for (uint32_t l = 0; l < L; ++l) {
                const uint32_t bytes = meta[l].size * (uint32_t)(C * sizeof(T));   // sizes are multiples of 8 rows: 16-byte granular
                const bool fits = meta[l].mode == 0 && (bytes & 15u) == 0 && used + bytes <= stage_bytes;
                staged_at[l] = fits ? used : NOT_STAGED;
                if (fits) used += bytes;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&stage_bar)), "r"(used) : "memory");
            for (uint32_t l = 0; l < L; ++l) {
                if (staged_at[l] == NOT_STAGED) continue;
                const uint32_t bytes = meta[l].size * (uint32_t)(C * sizeof(T));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 umma::smem_u32(staged_rows + staged_at[l])),
                             "l"(table + (size_t)meta[l].offset * C), "r"(bytes), "r"(umma::smem_u32(&stage_bar))
                             : "memory");
            }
        }
        __syncthreads();
        umma::mbar_wait(&stage_bar, 0);
    }

    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            x[d] = __ldg(inputs + (size_t)b * D + d);
            if (x[d] < 0 || x[d] > 1) oob = true;
        }

        for (uint32_t l0 = 0; l0 < L; l0 += LG) {
            union alignas(16) Group {
                T v[LG][C];
                uint4 q[(LG * C * sizeof(T) + 15) / 16];
                uint2 d[(LG * C * sizeof(T) + 7) / 8];
            } grp;
            T (&res)[LG][C] = grp.v;
#pragma unroll
            for (int j = 0; j < LG; ++j) {
#pragma unroll
                for (int c = 0; c < C; ++c) res[j][c] = from_f<T>(0.0f);
            }
#pragma unroll
            for (int j = 0; j < LG; ++j) {
                const uint32_t l = l0 + j;
                if (l >= L) break;
                if (oob) {
                    if (dy_dx) {
                        T* g = dy_dx + ((size_t)b * L + l) * D * C;
#pragma unroll
                        for (int k = 0; k < D * C; ++k) g[k] = from_f<T>(0.0f);
                    }
                    continue;
                }
                const LevelMeta m = meta[l];
                const T* __restrict__ tbl = table + (size_t)m.offset * C;
                const Cell<D> cell = locate<D>(x, m, align_corners != 0, interp);

                Row<T, C> rows[1 << D];
                bool from_shared = false;
                if constexpr (STAGE) from_shared = staged_at[l] != NOT_STAGED;     // uniform over the CTA (every thread is on level l)
                if (from_shared) {
                    const T* rows_l = reinterpret_cast<const T*>(staged_rows + staged_at[l]);
#pragma unroll
                    for (uint32_t k = 0; k < (1u << D); ++k)
                        rows[k] = load_row_staged<T, C>(rows_l + (size_t)corner_row<D>(m, cell.pg, k) * C);
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < (1u << D); ++k)
                        rows[k] = load_row<T, C>(tbl + (size_t)corner_row<D>(m, cell.pg, k) * C);
                }
#pragma unroll
                for (uint32_t k = 0; k < (1u << D); ++k) {
                    const float w = corner_weight<D>(cell, k);
#pragma unroll
                    for (int c = 0; c < C; ++c) accum(res[j][c], w, rows[k].v[c]);
                }

                if (dy_dx) {
                    // analytic d(out)/d(x): for each axis, difference of the two faces (gridencoder.cu:200-243).
                    // The 2^D rows are already in registers: corner k with bit gd set / cleared are the faces.
                    T* g = dy_dx + ((size_t)b * L + l) * D * C;
#pragma unroll
                    for (int gd = 0; gd < D; ++gd) {
                        T acc[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) acc[c] = from_f<T>(0.0f);
#pragma unroll
                        for (uint32_t idx = 0; idx < (1u << (D - 1)); ++idx) {
                            float w = m.scale;
                            uint32_t corner = 0;
#pragma unroll
                            for (int nd = 0; nd < D - 1; ++nd) {
                                const int d = (nd >= gd) ? nd + 1 : nd;
                                if ((idx >> nd) & 1u) {
                                    w *= cell.frac[d];
                                    corner |= 1u << d;
                                } else {
                                    w *= 1.0f - cell.frac[d];
                                }
                            }
#pragma unroll
                            for (int c = 0; c < C; ++c)
                                accum_diff(acc[c], w, rows[corner | (1u << gd)].v[c], rows[corner].v[c], cell.deriv[gd]);
                        }
#pragma unroll
                        for (int c = 0; c < C; ++c) g[gd * C + c] = acc[c];
                    }
                }
            }

            // ---- write the group ----
            if (layout == RN_LAYOUT_BLC) {
                T* o = outputs + (size_t)b * L * C + (size_t)l0 * C;
                if (VEC_OUT && l0 + LG <= L) {
                    constexpr int BYTES = LG * C * sizeof(T);
                    if constexpr (BYTES % 16 == 0) {
#pragma unroll
                        for (int q = 0; q < BYTES / 16; ++q) reinterpret_cast<uint4*>(o)[q] = grp.q[q];
                    } else {
#pragma unroll
                        for (int q = 0; q < BYTES / 8; ++q) reinterpret_cast<uint2*>(o)[q] = grp.d[q];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < LG; ++j) {
                        if (l0 + j < L) {
#pragma unroll
                            for (int c = 0; c < C; ++c) o[j * C + c] = res[j][c];
                        }
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < LG; ++j) {
                    if (l0 + j < L) {
                        T* o = outputs + ((size_t)(l0 + j) * B + b) * C;
#pragma unroll
                        for (int c = 0; c < C; ++c) o[c] = res[j][c];
                    }
                }
            }
        }
    }
}

// ======================================================================================================
// backward: table gradient (fp32 or fp16 accumulate target) + fused input gradient
// ======================================================================================================
__device__ __forceinline__ void red_add(float* p, float a) { atomicAdd(p, a); }
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add2(__half* p, float a, float b) {
    const __half2 v = __floats2half2_rn(a, b);
    atomicAdd(reinterpret_cast<__half2*>(p), v);
}

template <typename G, int C>
__device__ __forceinline__ void scatter_row(G* __restrict__ p, const float (&v)[C]) {
    if constexpr (sizeof(G) == 4) {
        if constexpr (C == 1) red_add((float*)p, v[0]);
        else if constexpr (C == 2) red_add2((float*)p, v[0], v[1]);
        else {
#pragma unroll
            for (int c = 0; c < C; c += 4) red_add4((float*)p + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
        }
    } else {
        if constexpr (C == 1) atomicAdd((__half*)p, __float2half_rn(v[0]));
        else {
#pragma unroll
            for (int c = 0; c < C; c += 2) red_add2((__half*)p + c, v[c], v[c + 1]);
        }
    }
}

template <typename T, typename G, int D, int C>
__global__ void __launch_bounds__(256)
grid_backward_kernel(const T* __restrict__ grad, const float* __restrict__ inputs, const int32_t* __restrict__ offsets,
                     G* __restrict__ grad_table, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs, uint32_t B,
                     uint32_t L, float S, uint32_t H, uint32_t gridtype, uint32_t align_corners, uint32_t interp,
                     uint32_t layout) {
    __shared__ LevelMeta meta[MAX_LEVELS];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x)
        make_level_meta(meta[l], l, offsets, S, H, D, gridtype, align_corners != 0);
    __syncthreads();

    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            x[d] = __ldg(inputs + (size_t)b * D + d);
            if (x[d] < 0 || x[d] > 1) oob = true;
        }
        float gin[D];
#pragma unroll
        for (int d = 0; d < D; ++d) gin[d] = 0.0f;

        for (uint32_t l = 0; l < L; ++l) {
            const T* gp = (layout == RN_LAYOUT_BLC) ? grad + (size_t)b * L * C + (size_t)l * C
                                                    : grad + ((size_t)l * B + b) * C;
            const Row<T, C> g = load_row<T, C>(gp);
            if (dy_dx) {
                // grad_inputs[b,d] = sum_{l,c} grad[l,b,c] * dy_dx[b,l,d,c]     (gridencoder.cu:342-368)
                const T* dp = dy_dx + ((size_t)b * L + l) * D * C;
#pragma unroll
                for (int d = 0; d < D; ++d) {
#pragma unroll
                    for (int c = 0; c < C; ++c) gin[d] = __fmaf_rn(to_f(g.v[c]), to_f(__ldg(dp + d * C + c)), gin[d]);
                }
            }
            if (oob) continue;  // reference returns before touching the table (gridencoder.cu:275-280)
            const LevelMeta m = meta[l];
            G* __restrict__ gt = grad_table + (size_t)m.offset * C;
            const Cell<D> cell = locate<D>(x, m, align_corners != 0, interp);
            float gf[C];
#pragma unroll
            for (int c = 0; c < C; ++c) gf[c] = to_f(g.v[c]);
#pragma unroll
            for (uint32_t k = 0; k < (1u << D); ++k) {
                const float w = corner_weight<D>(cell, k);
                float v[C];
#pragma unroll
                for (int c = 0; c < C; ++c) v[c] = w * gf[c];
                scatter_row<G, C>(gt + (size_t)corner_row<D>(m, cell.pg, k) * C, v);
            }
        }
        if (dy_dx && grad_inputs) {
#pragma unroll
            for (int d = 0; d < D; ++d) grad_inputs[(size_t)b * D + d] = from_f<T>(gin[d]);
        }
    }
}

// Warp-aggregated variant for D == 2 -- the ambient grid.  Its input is the network's own 2-D output, tanh(ambient_net(...)),
// a smooth function of position and audio that the ambient regulariser (nerf/utils.py:783-806) pulls towards one point: the
// 32 samples of a warp fall into the SAME cell on the coarse levels, and so do most of the batch's other warps.  One vector
// atomic per (lane, corner) then serialises on a handful of L2 addresses (measured: 1.77 ms for 206 k samples, 5x the 3-D
// table's backward and a third of the whole training step).  Here the lanes of a warp that share a cell (match.any on the
// packed cell coordinate) are summed with shuffles first and ONE lane issues the four atomics; a warp whose lanes spread over
// more than kMaxGroups cells (fine levels, uniform inputs) keeps the per-lane atomics.  All 32 lanes stay in the loop (the
// sample loop is uniform per CTA) so every shuffle runs with the full mask.
constexpr int kMaxGroups = 8;

// Second stage for the clustered case (HASH = true; fp32 table gradient, C = 2): what a warp has aggregated goes into a small hash
// table in SHARED memory keyed by the table row (all levels share it), and the CTA -- persistent, a few thousand samples each --
// flushes every touched row to global memory once.  As training proceeds the ambient coordinates of the whole batch collapse onto a
// few cells per level; after the warp stage that still left ~60 atomics per sample-warp aimed at a few hundred L2 addresses, and
// the kernel grew from 0.14 to 0.79 ms per step over the first 70 steps of a run (profiles/r02_train_timeline_graphed_after_70_steps.txt).
// A full table (linear probing, 4 probes) falls back to the global atomic, so uniform inputs cost what they did.
constexpr uint32_t kHashLog = 11, kHashSlots = 1u << kHashLog, kHashEmpty = 0xffffffffu;

__device__ __forceinline__ bool hash_accumulate(uint32_t* keys, float2* vals, uint32_t key, float a, float b) {
    uint32_t slot = (key * 2654435761u) >> (32 - kHashLog);
#pragma unroll
    for (int probe = 0; probe < 4; ++probe) {
        const uint32_t prev = atomicCAS(&keys[slot], kHashEmpty, key);
        if (prev == kHashEmpty || prev == key) {
            atomicAdd(&vals[slot].x, a);
            atomicAdd(&vals[slot].y, b);
            return true;
        }
        slot = (slot + 1) & (kHashSlots - 1);
    }
    return false;
}

template <typename T, typename G, int C, bool HASH = false>
__global__ void __launch_bounds__(256)
grid_backward_shared_cell_kernel(const T* __restrict__ grad, const float* __restrict__ inputs, const int32_t* __restrict__ offsets,
                                 G* __restrict__ grad_table, const T* __restrict__ dy_dx, T* __restrict__ grad_inputs, uint32_t B,
                                 uint32_t L, float S, uint32_t H, uint32_t gridtype, uint32_t align_corners, uint32_t interp,
                                 uint32_t layout) {
    constexpr int D = 2;
    constexpr uint32_t FULL = 0xffffffffu;
    __shared__ LevelMeta meta[MAX_LEVELS];
    __shared__ uint32_t h_keys[HASH ? kHashSlots : 1];
    __shared__ float2 h_vals[HASH ? kHashSlots : 1];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x)
        make_level_meta(meta[l], l, offsets, S, H, D, gridtype, align_corners != 0);
    if constexpr (HASH) {
        for (uint32_t i = threadIdx.x; i < kHashSlots; i += blockDim.x) { h_keys[i] = kHashEmpty; h_vals[i] = make_float2(0.f, 0.f); }
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u;

    for (uint32_t base = blockIdx.x * blockDim.x; base < B; base += gridDim.x * blockDim.x) {
        const uint32_t b = base + threadIdx.x;
        const bool in_range = b < B;
        float x[D] = {0.0f, 0.0f};
        bool valid = in_range;
        if (in_range) {
#pragma unroll
            for (int d = 0; d < D; ++d) {
                x[d] = __ldg(inputs + (size_t)b * D + d);
                if (x[d] < 0 || x[d] > 1) valid = false;   // reference returns before touching the table (gridencoder.cu:275-280)
            }
        }
        float gin[D] = {0.0f, 0.0f};

        // fp16 rows of 16 levels x 2 features in [B, L*C] order (the fused training step, the autocast GridEncoder): the sample's whole
        // gradient row (64 B) and d(features)/d(x) row (128 B) are fetched with 12 wide loads BEFORE the level loop -- one memory round
        // trip per sample instead of two per level (the kernel sat at 27 % issue utilisation waiting on 32 dependent 4- and 8-byte loads)
        constexpr bool kRowPrefetch = sizeof(T) == 2 && C == 2;
        const bool prefetch = kRowPrefetch && layout == RN_LAYOUT_BLC && L == 16 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0 &&
                              (reinterpret_cast<uintptr_t>(dy_dx) & 15) == 0;     // CTA-uniform
        uint32_t gw[16], dw[32];    // indexed by the level below: lives in (L1-resident) local memory, ~30 cycles instead of an L2 round trip
        if (prefetch && in_range) {
            const uint4* gp4 = reinterpret_cast<const uint4*>(grad + (size_t)b * 32);
            uint4 q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = __ldg(gp4 + i);
            uint4 r[8] = {};
            if (dy_dx) {
                const uint4* dp4 = reinterpret_cast<const uint4*>(dy_dx + (size_t)b * 64);
#pragma unroll
                for (int i = 0; i < 8; ++i) r[i] = __ldg(dp4 + i);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { gw[4 * i] = q[i].x; gw[4 * i + 1] = q[i].y; gw[4 * i + 2] = q[i].z; gw[4 * i + 3] = q[i].w; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { dw[4 * i] = r[i].x; dw[4 * i + 1] = r[i].y; dw[4 * i + 2] = r[i].z; dw[4 * i + 3] = r[i].w; }
        }

#pragma unroll 1
        for (uint32_t l = 0; l < L; ++l) {
            float gf[C];
#pragma unroll
            for (int c = 0; c < C; ++c) gf[c] = 0.0f;
            if (prefetch) {
                if constexpr (kRowPrefetch) {
                    if (in_range) {
                        // level l: features = 32-bit word l of the 16-word gradient row, derivatives = words 2l, 2l+1 of the 32-word row
                        const uint32_t g32 = gw[l & 15u];
                        const __half2 gh = *reinterpret_cast<const __half2*>(&g32);
                        gf[0] = __low2float(gh); gf[1] = __high2float(gh);
                        if (dy_dx) {
                            const uint32_t d0 = dw[(2 * l) & 31u], d1 = dw[(2 * l + 1) & 31u];
                            const __half2 dx = *reinterpret_cast<const __half2*>(&d0), dy = *reinterpret_cast<const __half2*>(&d1);
                            gin[0] = __fmaf_rn(gf[0], __low2float(dx), gin[0]); gin[0] = __fmaf_rn(gf[1], __high2float(dx), gin[0]);
                            gin[1] = __fmaf_rn(gf[0], __low2float(dy), gin[1]); gin[1] = __fmaf_rn(gf[1], __high2float(dy), gin[1]);
                        }
                    }
                }
            } else if (in_range) {
                const T* gp = (layout == RN_LAYOUT_BLC) ? grad + (size_t)b * L * C + (size_t)l * C
                                                        : grad + ((size_t)l * B + b) * C;
                const Row<T, C> g = load_row<T, C>(gp);
#pragma unroll
                for (int c = 0; c < C; ++c) gf[c] = to_f(g.v[c]);
                if (dy_dx) {
                    const T* dp = dy_dx + ((size_t)b * L + l) * D * C;
#pragma unroll
                    for (int d = 0; d < D; ++d) {
#pragma unroll
                        for (int c = 0; c < C; ++c) gin[d] = __fmaf_rn(gf[c], to_f(__ldg(dp + d * C + c)), gin[d]);
                    }
                }
            }
            const LevelMeta m = meta[l];
            G* __restrict__ gt = grad_table + (size_t)m.offset * C;
            Cell<D> cell = {};
            uint32_t key = FULL;                            // never a real cell: packed coordinates stay below 2^16 each
            const bool packable = m.resolution < 65535u;    // warp-uniform; a wider level (not a RAD-NeRF geometry) keeps per-lane atomics
            if (valid) {
                cell = locate<D>(x, m, align_corners != 0, interp);
                if (packable) key = cell.pg[0] | (cell.pg[1] << 16);
            }
            const bool groupable = valid && packable;
            const uint32_t peers = __match_any_sync(FULL, key);
            const bool leader = groupable && lane == (uint32_t)(__ffs(peers) - 1);
            uint32_t leaders = __ballot_sync(FULL, leader);
            const uint32_t n_valid = __popc(__ballot_sync(FULL, groupable));
            const uint32_t n_groups = __popc(leaders);
            if (n_groups < n_valid) {     // warp-uniform: at least one cell is shared by several lanes
                // Segmented sum over the lanes of each cell in ONE pass for all cells of the warp: `redux.sync` (__reduce_add_sync) with
                // the match.any peer mask as member mask reduces every group at the same time.  It is an integer instruction, so the
                // eight products (4 corners x 2 features) are summed in fixed point: scaled by a power of two such that the group's
                // largest magnitude sits at 2^24 (a 32-term sum stays below 2^30).  The quantisation step is 2^-24 of the largest
                // term -- the resolution an fp32 accumulation of the same terms has -- and the result does not depend on lane order.
                // 1 + 8 redux per level instead of 40 shuffles per distinct cell (the kernel was shuffle-bound: 603 us at 0.8 M samples).
                float v[4][C];
                float amax = 0.0f;
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    const float w = groupable ? corner_weight<D>(cell, k) : 0.0f;
#pragma unroll
                    for (int c = 0; c < C; ++c) { v[k][c] = w * gf[c]; amax = fmaxf(amax, fabsf(v[k][c])); }
                }
                const uint32_t mbits = __reduce_max_sync(peers, __float_as_uint(amax));      // non-negative floats order like their bit patterns
                const int e = (int)((mbits >> 23) & 0xffu);                                  // biased exponent of the group's largest term
                const bool live = e >= 32 && e < 255;                                        // all-zero / denormal-tiny / non-finite groups: per-lane path below
                const float scale = __uint_as_float((uint32_t)(278 - max(e, 32)) << 23);     // 2^(24 - (e - 127))
                const float unscale = __uint_as_float((uint32_t)(max(e, 32) - 24) << 23);    // 2^((e - 127) - 24)
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int q = live ? __float2int_rn(v[k][c] * scale) : 0;
                        const int sum = __reduce_add_sync(peers, q);
                        if (live) v[k][c] = (float)sum * unscale;
                    }
                }
                if (groupable && (leader || !live)) {      // the group's leader carries the sums; a non-live group keeps per-lane atomics
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) {
                        const uint32_t row = corner_row<D>(m, cell.pg, k);
                        if constexpr (HASH) {
                            if (hash_accumulate(h_keys, h_vals, m.offset + row, v[k][0], v[k][C > 1 ? 1 : 0])) continue;
                        }
                        scatter_row<G, C>(gt + (size_t)row * C, v[k]);
                    }
                }
            } else if (valid) {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k) {
                    const float w = corner_weight<D>(cell, k);
                    float v[C];
#pragma unroll
                    for (int c = 0; c < C; ++c) v[c] = w * gf[c];
                    scatter_row<G, C>(gt + (size_t)corner_row<D>(m, cell.pg, k) * C, v);
                }
            }
        }
        if (dy_dx && grad_inputs && in_range) {
#pragma unroll
            for (int d = 0; d < D; ++d) grad_inputs[(size_t)b * D + d] = from_f<T>(gin[d]);
        }
    }
    if constexpr (HASH) {   // one flush per CTA: every touched row once
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < kHashSlots; i += blockDim.x) {
            const uint32_t key = h_keys[i];
            if (key == kHashEmpty) continue;
            const float2 a = h_vals[i];
            float v[C];
            v[0] = a.x;
            if constexpr (C > 1) v[1] = a.y;
            scatter_row<G, C>(grad_table + (size_t)key * C, v);
        }
    }
}

// ======================================================================================================
// total-variation gradient (API parity; RAD-NeRF's trainer never calls it)   (gridencoder.cu:505-609)
// ======================================================================================================
template <typename T, int D, int C>
__global__ void __launch_bounds__(256)
grid_tv_kernel(const T* __restrict__ inputs, const T* __restrict__ table, T* __restrict__ grad,
               const int32_t* __restrict__ offsets, float weight, uint32_t B, uint32_t L, float S, uint32_t H,
               uint32_t gridtype, uint32_t align_corners) {
    __shared__ LevelMeta meta[MAX_LEVELS];
    for (uint32_t l = threadIdx.x; l < L; l += blockDim.x)
        make_level_meta(meta[l], l, offsets, S, H, D, gridtype, align_corners != 0);
    __syncthreads();
    const uint32_t level = blockIdx.y;
    const LevelMeta m = meta[level];
    const T* __restrict__ tbl = table + (size_t)m.offset * C;
    T* __restrict__ gt = grad + (size_t)m.offset * C;
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        float x[D];
        bool oob = false;
#pragma unroll
        for (int d = 0; d < D; ++d) {
            x[d] = to_f(inputs[(size_t)b * D + d]);
            if (x[d] < 0 || x[d] > 1) oob = true;
        }
        if (oob) continue;
        uint32_t pg[D];
#pragma unroll
        for (int d = 0; d < D; ++d) pg[d] = (uint32_t)floorf(__fmaf_rn(x[d], m.scale, align_corners ? 0.0f : 0.5f));
        const uint32_t centre = corner_row<D>(m, pg, 0);
        const Row<T, C> c0 = load_row<T, C>(tbl + (size_t)centre * C);
        T sum[C], sq[C];
#pragma unroll
        for (int c = 0; c < C; ++c) sum[c] = sq[c] = from_f<T>(0.0f);
        const T w = from_f<T>(weight / (2 * D));
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const uint32_t cur = pg[d];
            if (cur < m.resolution) {
                pg[d] = cur + 1;
                const Row<T, C> n = load_row<T, C>(tbl + (size_t)corner_row<D>(m, pg, 0) * C);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const T dlt = from_f<T>(to_f(c0.v[c]) - to_f(n.v[c]));
                    sum[c] = from_f<T>(to_f(sum[c]) + to_f(dlt));
                    sq[c] = from_f<T>(to_f(sq[c]) + to_f(from_f<T>(to_f(dlt) * to_f(dlt))));
                }
            }
            if (cur > 0) {
                pg[d] = cur - 1;
                const Row<T, C> n = load_row<T, C>(tbl + (size_t)corner_row<D>(m, pg, 0) * C);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const T dlt = from_f<T>(to_f(c0.v[c]) - to_f(n.v[c]));
                    sum[c] = from_f<T>(to_f(sum[c]) + to_f(dlt));
                    sq[c] = from_f<T>(to_f(sq[c]) + to_f(from_f<T>(to_f(dlt) * to_f(dlt))));
                }
            }
            pg[d] = cur;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float ws = to_f(from_f<T>(to_f(w) * to_f(sum[c])));
            const float v = ws * rsqrtf(to_f(sq[c]) + 1e-9f);
            if constexpr (sizeof(T) == 4) atomicAdd((float*)gt + (size_t)centre * C + c, v);
            else atomicAdd((__half*)gt + (size_t)centre * C + c, __float2half_rn(v));
        }
    }
}

static __global__ void level_geometry_kernel(float S, uint32_t H, uint32_t L, float* scales, uint32_t* res) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const float scale = exp2f(l * S) * H - 1.0f;
    if (scales) scales[l] = scale;
    if (res) res[l] = (uint32_t)ceilf(scale) + 1;
}


// ---- per-dimension dispatch (one translation unit per D keeps compile times short) -------------------
struct FwdArgs {
    const float* inputs; const void* emb; const int32_t* offsets; void* out; void* dy_dx;
    uint32_t B, C, L; float S; uint32_t H, gridtype, ac, interp, dtype, layout; cudaStream_t st;
};
struct BwdArgs {
    const void* grad; const float* inputs; const int32_t* offsets; void* gt; const void* dy_dx; void* gin;
    uint32_t B, C, L; float S; uint32_t H, gridtype, ac, interp, dtype, layout, gdtype; cudaStream_t st;
};
struct TvArgs {
    const void* inputs; const void* emb; void* grad; const int32_t* offsets; float weight;
    uint32_t B, C, L; float S; uint32_t H, gridtype, ac, dtype; cudaStream_t st;
};
template <int D> int forward_d(const FwdArgs& a);
template <int D> int backward_d(const BwdArgs& a);
template <int D> int tv_d(const TvArgs& a);

// shared memory per CTA for the staged forward, from the environment on every call (a getenv is nanoseconds next to a launch; tests
// and tools switch it between calls); capped at what one CTA can opt into
inline uint32_t forward_stage_bytes() {
    const char* v = std::getenv("RADNERF_GRID_STAGE_KB");
    if (!v || !*v) return 0;
    long kb = std::strtol(v, nullptr, 10);
    if (kb <= 0) return 0;
    if (kb > 220) kb = 220;
    return (uint32_t)kb * 1024u;
}

template <typename T, int D, int C>
int launch_forward(const FwdArgs& a) {
    const uint32_t threads = 256;
    const uint32_t grid = wave_grid(a.B, threads, 32);
    const bool vec = a.layout == RN_LAYOUT_BLC && ((uintptr_t)a.out % 16 == 0) && ((a.L * C * sizeof(T)) % 16 == 0);
    if constexpr (D <= 3 && C == 2) {
        // coarse levels staged in shared memory (see STAGE above).  Opt-in: RADNERF_GRID_STAGE_KB=<KiB of shared memory per CTA>;
        // measured against the plain kernel in profiles/r02_grid_forward_staged.json (DESIGN.md section 4 says which way it went)
        const uint32_t stage_bytes = forward_stage_bytes();
        if (stage_bytes && vec && ((uintptr_t)a.emb % 16 == 0) && a.B >= 4u * RN_NUM_SMS * 512u) {
            auto kernel = grid_forward_kernel<T, D, C, true, true>;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes);
            if (e != cudaSuccess) { set_error("grid_forward (staged): cannot reserve %u bytes of shared memory: %s", stage_bytes, cudaGetErrorString(e)); return (int)e; }
            uint32_t per_sm = (227u * 1024u) / (stage_bytes + 5u * 1024u);      // + static shared memory and the per-CTA reservation
            per_sm = per_sm < 1 ? 1 : (per_sm > 2 ? 2 : per_sm);                 // 64 registers x 512 threads: two CTAs per SM at most
            kernel<<<wave_grid(a.B, 512, per_sm), 512, stage_bytes, a.st>>>(a.inputs, (const T*)a.emb, a.offsets, (T*)a.out, (T*)a.dy_dx, a.B,
                                                                          a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout, stage_bytes);
            return finish_launch("rn_grid_encode_forward (staged)");
        }
    }
    if (vec)
        grid_forward_kernel<T, D, C, true><<<grid, threads, 0, a.st>>>(a.inputs, (const T*)a.emb, a.offsets, (T*)a.out,
            (T*)a.dy_dx, a.B, a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout);
    else
        grid_forward_kernel<T, D, C, false><<<grid, threads, 0, a.st>>>(a.inputs, (const T*)a.emb, a.offsets, (T*)a.out,
            (T*)a.dy_dx, a.B, a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout);
    return finish_launch("rn_grid_encode_forward");
}

template <typename T, typename G, int D, int C>
int launch_backward(const BwdArgs& a) {
    const uint32_t threads = 256;
    const uint32_t grid = wave_grid(a.B, threads, 32);
// measured at 0.82 M clustered samples of a real training step (tools/train_timeline.py, WARM=70): 816 us before, 732 us with the row
// prefetch alone, 603 us with prefetch + hash stage; uniform inputs (tests/test_gpu_kernel_rows.py): 0.65 ms without, 0.70 ms with the
// hash stage (smaller persistent grid).  The training case is the clustered one.
#ifndef RN_BWD2_HASH
#define RN_BWD2_HASH 1
#endif
    if constexpr (RN_BWD2_HASH && D == 2 && C == 2 && sizeof(G) == 4) {
        // the ambient / torso grids: warp stage + per-CTA hash stage; a persistent grid (4 CTAs per SM) so that a CTA sees a few thousand
        // samples before it flushes
        const uint32_t pgrid = std::min(grid, (uint32_t)RN_NUM_SMS * 4u);
        grid_backward_shared_cell_kernel<T, G, C, true><<<pgrid, threads, 0, a.st>>>((const T*)a.grad, a.inputs, a.offsets, (G*)a.gt,
            (const T*)a.dy_dx, (T*)a.gin, a.B, a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout);
    } else if constexpr (D == 2) {   // lanes of a warp that share a cell are summed before the atomics
        grid_backward_shared_cell_kernel<T, G, C><<<grid, threads, 0, a.st>>>((const T*)a.grad, a.inputs, a.offsets, (G*)a.gt,
            (const T*)a.dy_dx, (T*)a.gin, a.B, a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout);
    } else {
        grid_backward_kernel<T, G, D, C><<<grid, threads, 0, a.st>>>((const T*)a.grad, a.inputs, a.offsets, (G*)a.gt,
            (const T*)a.dy_dx, (T*)a.gin, a.B, a.L, a.S, a.H, a.gridtype, a.ac, a.interp, a.layout);
    }
    return finish_launch("rn_grid_encode_backward");
}

template <typename T, int D, int C>
int launch_tv(const TvArgs& a) {
    const uint32_t threads = 256;
    dim3 grid(wave_grid(a.B, threads, 8), a.L, 1);
    grid_tv_kernel<T, D, C><<<grid, threads, 0, a.st>>>((const T*)a.inputs, (const T*)a.emb, (T*)a.grad, a.offsets,
        a.weight, a.B, a.L, a.S, a.H, a.gridtype, a.ac);
    return finish_launch("rn_grad_total_variation");
}

#define RN_SWITCH_C(C, EXPR)                               \
    switch (C) {                                           \
        case 1: { constexpr int kC = 1; return EXPR; }     \
        case 2: { constexpr int kC = 2; return EXPR; }     \
        case 4: { constexpr int kC = 4; return EXPR; }     \
        case 8: { constexpr int kC = 8; return EXPR; }     \
        default: return RN_E_UNSUPPORTED;                  \
    }

#define RN_GRID_DEFINE_D(DD)                                                                              \
    template <> int forward_d<DD>(const FwdArgs& a) {                                                     \
        if (a.dtype == RN_F16) { RN_SWITCH_C(a.C, (launch_forward<__half, DD, kC>(a))) }                  \
        RN_SWITCH_C(a.C, (launch_forward<float, DD, kC>(a)))                                              \
    }                                                                                                     \
    template <> int backward_d<DD>(const BwdArgs& a) {                                                    \
        if (a.dtype == RN_F16 && a.gdtype == RN_F32) { RN_SWITCH_C(a.C, (launch_backward<__half, float, DD, kC>(a))) } \
        if (a.dtype == RN_F16) { RN_SWITCH_C(a.C, (launch_backward<__half, __half, DD, kC>(a))) }         \
        RN_SWITCH_C(a.C, (launch_backward<float, float, DD, kC>(a)))                                      \
    }                                                                                                     \
    template <> int tv_d<DD>(const TvArgs& a) {                                                           \
        if (a.dtype == RN_F16) { RN_SWITCH_C(a.C, (launch_tv<__half, DD, kC>(a))) }                       \
        RN_SWITCH_C(a.C, (launch_tv<float, DD, kC>(a)))                                                   \
    }

}  // namespace grid
}  // namespace rn
