// raymarching.cu -- occupancy-grid utilities, ray marching and volume-rendering compositing for sm_100a.
//
// Replaces the reference's raymarching extension (raymarching/src/raymarching.cu) behind the C ABI of
// include/radnerf_b200.h; one entry point per reference function, same buffers, same in-place semantics.
//
// Bit-exact contracts (north_star): Morton codes, packed bitfields, per-ray sample counts and every float the
// marchers emit.  The marchers therefore spell each floating-point operation explicitly (__fmaf_rn / __fmul_rn /
// __fadd_rn) in the form nvcc contracts the reference expressions to (checked against the sm_100a SASS of the
// reference build: FFMA for o + t*d, x*rb + 1, (.)*2 - 1, (.)*mb - x, level*H3 + morton; the double-precision
// cell quantisation 0.5*(..)*H equals a single fp32 multiply by 0.5f*H for H < 2^24).
//
// What is different from the reference:
//   * march_rays_train reserves output space with ONE pair of global atomics per CTA (block-wide exclusive scan
//     of the per-ray counts) instead of two atomics per ray; the ray list comes out ordered inside a CTA.
//   * packbits reads 32 cells per thread with two 128-bit loads... (see kernels) and writes one 32-bit word.
//   * elementwise utilities use grid-stride loops sized to whole waves of the 148 SMs.
#include "common.cuh"
#include "march.cuh"
#include <float.h>

namespace rn {
namespace {

// ------------------------------------------------------------------------------------------------------
// small elementwise utilities
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void swapf(float& a, float& b) { float c = a; a = b; b = c; }

// slab test against an axis-aligned box                                   (raymarching.cu:91-145)
__device__ __forceinline__ void ray_aabb(float ox, float oy, float oz, float dx, float dy, float dz,
                                         const float* __restrict__ aabb, float min_near, float& near_out,
                                         float& far_out) {
    const float rdx = 1 / dx, rdy = 1 / dy, rdz = 1 / dz;
    float near = __fmul_rn(aabb[0] - ox, rdx);
    float far = __fmul_rn(aabb[3] - ox, rdx);
    if (near > far) swapf(near, far);
    float near_y = __fmul_rn(aabb[1] - oy, rdy);
    float far_y = __fmul_rn(aabb[4] - oy, rdy);
    if (near_y > far_y) swapf(near_y, far_y);
    if (near > far_y || near_y > far) { near_out = far_out = FLT_MAX; return; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = __fmul_rn(aabb[2] - oz, rdz);
    float far_z = __fmul_rn(aabb[5] - oz, rdz);
    if (near_z > far_z) swapf(near_z, far_z);
    if (near > far_z || near_z > far) { near_out = far_out = FLT_MAX; return; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    near_out = near;
    far_out = far;
}

__global__ void __launch_bounds__(256)
near_far_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
                uint32_t N, float min_near, float* __restrict__ nears, float* __restrict__ fars) {
    __shared__ float box[6];
    if (threadIdx.x < 6) box[threadIdx.x] = aabb[threadIdx.x];
    __syncthreads();
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        float nr, fr;
        ray_aabb(__ldg(o), __ldg(o + 1), __ldg(o + 2), __ldg(d), __ldg(d + 1), __ldg(d + 2), box, min_near, nr, fr);
        nears[n] = nr;
        fars[n] = fr;
    }
}

// ray / sphere(radius) intersection -> (theta, phi) in [-1,1]^2          (raymarching.cu:162-198)
__global__ void __launch_bounds__(256)
sph_from_ray_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, float radius, uint32_t N,
                    float* __restrict__ coords) {
    constexpr float RPI = 0.3183098861837907f;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
        const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
        const float A = dx * dx + dy * dy + dz * dz;
        const float Bh = ox * dx + oy * dy + oz * dz;
        const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
        const float t = (-Bh + sqrtf(Bh * Bh - A * Cc)) / A;
        const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
        const float theta = atan2f(sqrtf(x * x + z * z), y);
        const float phi = atan2f(z, x);
        coords[n * 2] = 2 * theta * RPI - 1;
        coords[n * 2 + 1] = phi * RPI;
    }
}

__global__ void __launch_bounds__(256)
morton3D_kernel(const int32_t* __restrict__ coords, uint32_t N, int32_t* __restrict__ indices) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int32_t* c = coords + (size_t)n * 3;
        indices[n] = (int32_t)morton_encode((uint32_t)__ldg(c), (uint32_t)__ldg(c + 1), (uint32_t)__ldg(c + 2));
    }
}

__global__ void __launch_bounds__(256)
morton3D_invert_kernel(const int32_t* __restrict__ indices, uint32_t N, int32_t* __restrict__ coords) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const int32_t ind = __ldg(indices + n);  // arithmetic shift on the signed value, as the reference
        coords[(size_t)n * 3 + 0] = (int32_t)compact3((uint32_t)(ind >> 0));
        coords[(size_t)n * 3 + 1] = (int32_t)compact3((uint32_t)(ind >> 1));
        coords[(size_t)n * 3 + 2] = (int32_t)compact3((uint32_t)(ind >> 2));
    }
}

// threshold -> LSB-first bitfield                                       (raymarching.cu:267-289)
// One thread packs 32 cells (8 x float4, issued back to back) into one 32-bit word: 128 B in, 4 B out per thread,
// every load a full 16-byte vector.  Tail bytes (N % 4) are handled by a scalar epilogue.
__device__ __forceinline__ uint32_t pack8(const float4 a, const float4 b, float th) {
    uint32_t r = 0;
    r |= (a.x > th) ? 1u : 0u;
    r |= (a.y > th) ? 2u : 0u;
    r |= (a.z > th) ? 4u : 0u;
    r |= (a.w > th) ? 8u : 0u;
    r |= (b.x > th) ? 16u : 0u;
    r |= (b.y > th) ? 32u : 0u;
    r |= (b.z > th) ? 64u : 0u;
    r |= (b.w > th) ? 128u : 0u;
    return r;
}

__global__ void __launch_bounds__(256)
packbits_kernel(const float* __restrict__ grid, uint32_t N, float thresh, const float* __restrict__ thresh_dev, uint8_t* __restrict__ bitfield,
                uint32_t vec_ok) {
    if (thresh_dev) thresh = fminf(__ldg(thresh_dev), thresh);   // min(mean density on the device, density_thresh): renderer.py:471
    const uint32_t nwords = vec_ok ? N / 4 : 0;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += gridDim.x * blockDim.x) {
        const float4* g = reinterpret_cast<const float4*>(grid) + (size_t)w * 8;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldcs(g + i);
        uint32_t word = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) word |= pack8(v[2 * i], v[2 * i + 1], thresh) << (8 * i);
        reinterpret_cast<uint32_t*>(bitfield)[w] = word;
    }
    // scalar tail / unaligned fallback, one byte per thread
    for (uint32_t n = nwords * 4 + blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float* g = grid + (size_t)n * 8;
        uint32_t bits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) bits |= (g[i] > thresh) ? (1u << i) : 0u;
        bitfield[n] = (uint8_t)bits;
    }
}

// Occupancy merge of update_extra_state (nerf/renderer.py:463-468): where BOTH the running grid and the freshly queried (dilated) grid
// are valid (>= 0), grid = max(grid * decay, fresh); then mean(clamp(grid, 0)) over all cells.  One pass, in place, and the mean
// stays on the device (per-CTA partial sums in double, added in CTA order by the last CTA: deterministic) -- the reference's boolean
// mask indexing and `.item()` cost three host synchronisations per update.
__global__ void __launch_bounds__(256)
occupancy_merge_kernel(float* __restrict__ grid, const float* __restrict__ fresh, uint32_t n, float decay, double* __restrict__ partials,
                       uint32_t* __restrict__ ticket, float* __restrict__ mean_out) {
    __shared__ double s_sum[8];
    __shared__ uint32_t s_last;
    double acc = 0.0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float g = grid[i];
        const float f = __ldg(fresh + i);
        if (g >= 0.f && f >= 0.f) {
            g = fmaxf(__fmul_rn(g, decay), f);
            grid[i] = g;
        }
        acc += (double)fmaxf(g, 0.f);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_sum[w];
        partials[blockIdx.x] = t;
        __threadfence();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double t = 0.0;
        for (uint32_t b = 0; b < gridDim.x; ++b) t += partials[b];
        *mean_out = (float)(t / (double)n);
        *ticket = 0u;
    }
}

// 6-neighbour max in a Morton-indexed grid                              (raymarching.cu:304-335)
// A neighbour's Morton code is the voxel's own code with ONE coordinate stepped by +-1; on the interleaved bits that is a
// "dilated" add -- fill the other coordinates' bit positions with ones (increment) or zeros (decrement), add / subtract 1,
// mask -- 4 operations instead of a 3-coordinate re-encode (~35).  The coordinates are decoded once for the border tests.
constexpr uint32_t MORTON_X = 0x09249249u;   // bit positions 0, 3, 6, ... of a 10-bit coordinate
__device__ __forceinline__ uint32_t morton_step_up(uint32_t m, uint32_t mask) { return (((m | ~mask) + 1u) & mask) | (m & ~mask); }
__device__ __forceinline__ uint32_t morton_step_down(uint32_t m, uint32_t mask) { return (((m & mask) - 1u) & mask) | (m & ~mask); }

__global__ void __launch_bounds__(256)
dilation_kernel(const float* __restrict__ grid, uint32_t C, uint32_t H, float* __restrict__ out) {
    const uint32_t H3 = H * H * H;
    const uint32_t total = C * H3;
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < total; n += gridDim.x * blockDim.x) {
        const uint32_t c = n / H3;
        const uint32_t ind = n - c * H3;
        const uint32_t x = compact3(ind), y = compact3(ind >> 1), z = compact3(ind >> 2);
        const float* g = grid + (size_t)c * H3;
        float res = __ldg(grid + n);
        if (x + 1 < H) res = fmaxf(res, __ldg(g + morton_step_up(ind, MORTON_X)));
        if (x > 0) res = fmaxf(res, __ldg(g + morton_step_down(ind, MORTON_X)));
        if (y + 1 < H) res = fmaxf(res, __ldg(g + morton_step_up(ind, MORTON_X << 1)));
        if (y > 0) res = fmaxf(res, __ldg(g + morton_step_down(ind, MORTON_X << 1)));
        if (z + 1 < H) res = fmaxf(res, __ldg(g + morton_step_up(ind, MORTON_X << 2)));
        if (z > 0) res = fmaxf(res, __ldg(g + morton_step_down(ind, MORTON_X << 2)));
        out[n] = res;
    }
}

}  // namespace

namespace {

// ------------------------------------------------------------------------------------------------------
// training march: count, reserve (one atomic pair per CTA), write        (raymarching.cu:352-518)
// ------------------------------------------------------------------------------------------------------
constexpr int TRAIN_THREADS = 256;

__global__ void __launch_bounds__(TRAIN_THREADS)
march_rays_train_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, MarchParams p,
                        uint32_t max_steps, uint32_t N, uint32_t M, const float* __restrict__ nears,
                        const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                        float* __restrict__ deltas, int32_t* __restrict__ rays, int32_t* __restrict__ counter,
                        const float* __restrict__ noises, const int32_t* __restrict__ budget) {
    __shared__ uint32_t warp_sums[TRAIN_THREADS / 32];
    __shared__ uint32_t base_point, base_ray;

    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = n < N;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    Ray r;
    float t0 = 0.f, far = 0.f;
    uint32_t num_steps = 0;
    if (valid) {
        r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
        const float near = __ldg(nears + n);
        far = __ldg(fars + n);
        t0 = __fmaf_rn(step_size(p, near), __ldg(noises + n), near);
        float t = t0, x, y, z, dt;
        while (t < far && num_steps < max_steps) {
            if (march_probe(p, r, t, x, y, z, dt)) {
                ++num_steps;
                t = __fadd_rn(t, dt);
            }
        }
    }

    // block-wide exclusive scan of num_steps
    uint32_t incl = num_steps;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < TRAIN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int off = 1; off < TRAIN_THREADS / 32; off <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= (uint32_t)off) w += v;
        }
        if (lane < TRAIN_THREADS / 32) warp_sums[lane] = w;  // inclusive warp totals
        if (lane == TRAIN_THREADS / 32 - 1) {
            const uint32_t rays_here = min((uint32_t)TRAIN_THREADS, N - blockIdx.x * blockDim.x);
            base_point = (uint32_t)atomicAdd(counter, (int)w);
            base_ray = (uint32_t)atomicAdd(counter + 1, (int)rays_here);
        }
    }
    __syncthreads();
    if (!valid) return;

    const uint32_t point_index = base_point + (warp ? warp_sums[warp - 1] : 0) + incl - num_steps;
    const uint32_t ray_index = base_ray + threadIdx.x;
    rays[(size_t)ray_index * 3 + 0] = (int32_t)n;
    rays[(size_t)ray_index * 3 + 1] = (int32_t)point_index;
    rays[(size_t)ray_index * 3 + 2] = (int32_t)num_steps;

    if (num_steps == 0) return;
    // over budget: the ray is dropped (raymarching.cu:457).  The budget is M, or -- for a step replayed from a CUDA graph,
    // whose buffers keep a fixed capacity M while the reference's running estimate `mean_count` moves -- min(M, *budget)
    const uint32_t limit = budget ? min(M, (uint32_t)max(__ldg(budget), 0)) : M;
    if (point_index + num_steps > limit) return;

    float* px = xyzs + (size_t)point_index * 3;
    float* pd = dirs + (size_t)point_index * 3;
    float* pt = deltas + (size_t)point_index * 2;
    float t = t0, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
        if (march_probe(p, r, t, x, y, z, dt)) {
            t = __fadd_rn(t, dt);
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            *reinterpret_cast<float2*>(pt) = make_float2(dt, t);
            px += 3; pd += 3; pt += 2;
            ++step;
        }
    }
}

__global__ void __launch_bounds__(256)
march_rays_train_backward_kernel(const float* __restrict__ grad_xyzs, const float* __restrict__ grad_dirs,
                                 const int32_t* __restrict__ rays, const float* __restrict__ deltas, uint32_t N,
                                 uint32_t M, float* __restrict__ grad_rays_o, float* __restrict__ grad_rays_d) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= N) return;
    // The reference adds into row `slot` of grad_rays_* (raymarching.cu:550-555) although slot k of `rays` describes
    // ray rays[k,0]; the two agree only while the ray list is in identity order.  We address by the ray id, which
    // is what the gradient means ("deliberate deviations").
    const uint32_t n = (uint32_t)rays[(size_t)slot * 3];
    const uint32_t offset = (uint32_t)rays[(size_t)slot * 3 + 1];
    const uint32_t num_steps = (uint32_t)rays[(size_t)slot * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) return;
    const float* gx = grad_xyzs + (size_t)offset * 3;
    const float* gd = grad_dirs + (size_t)offset * 3;
    const float* dl = deltas + (size_t)offset * 2;
    float o0 = grad_rays_o[(size_t)n * 3], o1 = grad_rays_o[(size_t)n * 3 + 1], o2 = grad_rays_o[(size_t)n * 3 + 2];
    float d0 = grad_rays_d[(size_t)n * 3], d1 = grad_rays_d[(size_t)n * 3 + 1], d2 = grad_rays_d[(size_t)n * 3 + 2];
    for (uint32_t s = 0; s < num_steps; ++s) {
        const float t = dl[1];
        o0 += gx[0]; o1 += gx[1]; o2 += gx[2];
        d0 += __fmaf_rn(gx[0], t, gd[0]);
        d1 += __fmaf_rn(gx[1], t, gd[1]);
        d2 += __fmaf_rn(gx[2], t, gd[2]);
        gx += 3; gd += 3; dl += 2;
    }
    grad_rays_o[(size_t)n * 3] = o0; grad_rays_o[(size_t)n * 3 + 1] = o1; grad_rays_o[(size_t)n * 3 + 2] = o2;
    grad_rays_d[(size_t)n * 3] = d0; grad_rays_d[(size_t)n * 3 + 1] = d1; grad_rays_d[(size_t)n * 3 + 2] = d2;
}

// ------------------------------------------------------------------------------------------------------
// training compositing                                                  (raymarching.cu:603-687, :711-809)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
composite_train_fwd_kernel(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                           const float* __restrict__ ambient, const float* __restrict__ deltas,
                           const int32_t* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh,
                           float* __restrict__ weights_sum, float* __restrict__ ambient_sum, float* __restrict__ depth,
                           float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3);
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1);
    const uint32_t num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);

    float T = 1.0f, r = 0, g = 0, b = 0, ws = 0, d = 0, amb = 0;
    if (!(num_steps == 0 || offset + num_steps > M)) {
        const float* sg = sigmas + offset;
        const float* rg = rgbs + (size_t)offset * 3;
        const float* am = ambient + offset;
        const float2* dl = reinterpret_cast<const float2*>(deltas) + offset;
        // The early exit makes every step's loads wait for the previous step's test: a chain of up to 16 memory
        // latencies.  The next step's operands are therefore fetched before the current step is evaluated (the range
        // [offset, offset + num_steps) was checked above, so the extra loads are in bounds).
        float2 dd = __ldg(dl);
        float sgm = __ldg(sg), c0 = __ldg(rg), c1 = __ldg(rg + 1), c2 = __ldg(rg + 2), a = __ldg(am);
        for (uint32_t s = 0; s < num_steps; ++s) {
            const uint32_t sn = min(s + 1, num_steps - 1);
            const float2 dd_n = __ldg(dl + sn);
            const float sgm_n = __ldg(sg + sn), c0_n = __ldg(rg + 3 * sn), c1_n = __ldg(rg + 3 * sn + 1), c2_n = __ldg(rg + 3 * sn + 2),
                        a_n = __ldg(am + sn);
            const float alpha = 1.0f - __expf(-sgm * dd.x);
            const float weight = __fmul_rn(alpha, T);
            r = __fmaf_rn(weight, c0, r);
            g = __fmaf_rn(weight, c1, g);
            b = __fmaf_rn(weight, c2, b);
            d = __fmaf_rn(weight, dd.y, d);
            ws = __fadd_rn(ws, weight);
            amb = __fadd_rn(amb, a);
            T = __fmul_rn(T, 1.0f - alpha);
            if (T < T_thresh) break;
            dd = dd_n; sgm = sgm_n; c0 = c0_n; c1 = c1_n; c2 = c2_n; a = a_n;
        }
    }
    weights_sum[index] = ws;
    ambient_sum[index] = amb;
    depth[index] = d;
    image[(size_t)index * 3] = r;
    image[(size_t)index * 3 + 1] = g;
    image[(size_t)index * 3 + 2] = b;
}

__global__ void __launch_bounds__(256)
composite_train_bwd_kernel(const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_ambient_sum,
                           const float* __restrict__ grad_image, const float* __restrict__ sigmas,
                           const float* __restrict__ rgbs, const float* __restrict__ deltas,
                           const int32_t* __restrict__ rays, const float* __restrict__ weights_sum,
                           const float* __restrict__ image, uint32_t M, uint32_t N, float T_thresh,
                           float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs,
                           float* __restrict__ grad_ambient) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const uint32_t index = (uint32_t)__ldg(rays + (size_t)n * 3);
    const uint32_t offset = (uint32_t)__ldg(rays + (size_t)n * 3 + 1);
    const uint32_t num_steps = (uint32_t)__ldg(rays + (size_t)n * 3 + 2);
    if (num_steps == 0 || offset + num_steps > M) return;

    const float gws = __ldg(grad_weights_sum + index);
    const float gam = __ldg(grad_ambient_sum + index);
    const float gi0 = __ldg(grad_image + (size_t)index * 3), gi1 = __ldg(grad_image + (size_t)index * 3 + 1),
                gi2 = __ldg(grad_image + (size_t)index * 3 + 2);
    const float r_final = __ldg(image + (size_t)index * 3), g_final = __ldg(image + (size_t)index * 3 + 1),
                b_final = __ldg(image + (size_t)index * 3 + 2);
    const float ws_final = __ldg(weights_sum + index);
    const float tail = __fmul_rn(gws, 1 - ws_final);

    const float* sg = sigmas + offset;
    const float* rg = rgbs + (size_t)offset * 3;
    const float2* dl = reinterpret_cast<const float2*>(deltas) + offset;
    float* gs = grad_sigmas + offset;
    float* gr = grad_rgbs + (size_t)offset * 3;
    float* ga = grad_ambient + offset;

    float T = 1.0f, r = 0, g = 0, b = 0;
    // operands of step s + 1 are fetched before step s is evaluated (see composite_train_fwd_kernel)
    float dt = __ldg(dl).x, sgm = __ldg(sg), c0 = __ldg(rg), c1 = __ldg(rg + 1), c2 = __ldg(rg + 2);
    for (uint32_t s = 0; s < num_steps; ++s) {
        const uint32_t sn = min(s + 1, num_steps - 1);
        const float dt_n = __ldg(dl + sn).x, sgm_n = __ldg(sg + sn);
        const float c0_n = __ldg(rg + 3 * sn), c1_n = __ldg(rg + 3 * sn + 1), c2_n = __ldg(rg + 3 * sn + 2);
        const float alpha = 1.0f - __expf(-sgm * dt);
        const float weight = __fmul_rn(alpha, T);
        r = __fmaf_rn(weight, c0, r);
        g = __fmaf_rn(weight, c1, g);
        b = __fmaf_rn(weight, c2, b);
        T = __fmul_rn(T, 1.0f - alpha);
        gr[3 * s] = __fmul_rn(gi0, weight);
        gr[3 * s + 1] = __fmul_rn(gi1, weight);
        gr[3 * s + 2] = __fmul_rn(gi2, weight);
        ga[s] = gam;
        // d(image)/d(sigma_s) = dt * ( T_{s+1} * c_s - (C_final - C_s) ), plus the weights_sum term
        float acc = __fmul_rn(gi0, __fmaf_rn(T, c0, -(r_final - r)));
        acc = __fmaf_rn(gi1, __fmaf_rn(T, c1, -(g_final - g)), acc);
        acc = __fmaf_rn(gi2, __fmaf_rn(T, c2, -(b_final - b)), acc);
        acc = __fadd_rn(acc, tail);
        gs[s] = __fmul_rn(dt, acc);
        if (T < T_thresh) break;
        dt = dt_n; sgm = sgm_n; c0 = c0_n; c1 = c1_n; c2 = c2_n;
    }
}

// ------------------------------------------------------------------------------------------------------
// inference march / composite with the reference's slot layout          (raymarching.cu:827-929, :942-1029)
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
march_rays_kernel(uint32_t n_alive, uint32_t n_step, const int32_t* __restrict__ rays_alive,
                  const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  MarchParams p, const float* __restrict__ fars, float* __restrict__ xyzs, float* __restrict__ dirs,
                  float* __restrict__ deltas, const float* __restrict__ noises) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = __ldg(rays_alive + n);
    Ray r;
    r.load(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
    float t = __ldg(rays_t + index);
    const float far = __ldg(fars + index);
    t = __fmaf_rn(__ldg(noises + n), step_size(p, t), t);

    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float* pt = deltas + (size_t)n * n_step * 2;
    uint32_t step = 0;
    float x, y, z, dt;
    while (t < far && step < n_step) {
        if (march_probe(p, r, t, x, y, z, dt)) {
            t = __fadd_rn(t, dt);
            px[0] = x; px[1] = y; px[2] = z;
            pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
            *reinterpret_cast<float2*>(pt) = make_float2(dt, t);
            px += 3; pd += 3; pt += 2;
            ++step;
        }
    }
}

__global__ void __launch_bounds__(128)
composite_rays_kernel(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* __restrict__ rays_alive,
                      float* __restrict__ rays_t, const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                      const float* __restrict__ deltas, float* __restrict__ weights_sum, float* __restrict__ depth,
                      float* __restrict__ image) {
    const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_alive) return;
    const int32_t index = rays_alive[n];
    const float* sg = sigmas + (size_t)n * n_step;
    const float* rg = rgbs + (size_t)n * n_step * 3;
    const float2* dl = reinterpret_cast<const float2*>(deltas) + (size_t)n * n_step;

    float t = rays_t[index];
    float ws = weights_sum[index], d = depth[index];
    float r = image[(size_t)index * 3], g = image[(size_t)index * 3 + 1], b = image[(size_t)index * 3 + 2];

    uint32_t step = 0;
    while (step < n_step) {
        const float2 dd = __ldg(dl + step);
        if (dd.x == 0) break;  // zero-filled slot: the marcher ran out of samples
        const float alpha = 1.0f - __expf(-__ldg(sg + step) * dd.x);
        const float T = 1 - ws;
        const float weight = __fmul_rn(alpha, T);
        ws = __fadd_rn(ws, weight);
        t = dd.y;
        d = __fmaf_rn(weight, t, d);
        r = __fmaf_rn(weight, __ldg(rg + 3 * step), r);
        g = __fmaf_rn(weight, __ldg(rg + 3 * step + 1), g);
        b = __fmaf_rn(weight, __ldg(rg + 3 * step + 2), b);
        if (T < T_thresh) break;
        ++step;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = ws;
    depth[index] = d;
    image[(size_t)index * 3] = r;
    image[(size_t)index * 3 + 1] = g;
    image[(size_t)index * 3 + 2] = b;
}

}  // namespace
}  // namespace rn

using namespace rn;

#define RN_STREAM ((cudaStream_t)stream)

extern "C" int rn_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                                     float min_near, float* nears, float* fars, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(rays_o && rays_d && aabb && nears && fars, "null pointer");
    near_far_kernel<<<wave_grid(N, 256, 8), 256, 0, RN_STREAM>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
    return finish_launch("rn_near_far_from_aabb");
}

extern "C" int rn_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                               void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(rays_o && rays_d && coords, "null pointer");
    sph_from_ray_kernel<<<wave_grid(N, 256, 8), 256, 0, RN_STREAM>>>(rays_o, rays_d, radius, N, coords);
    return finish_launch("rn_sph_from_ray");
}

extern "C" int rn_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(coords && indices, "null pointer");
    morton3D_kernel<<<wave_grid(N, 256, 8), 256, 0, RN_STREAM>>>(coords, N, indices);
    return finish_launch("rn_morton3D");
}

extern "C" int rn_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(coords && indices, "null pointer");
    morton3D_invert_kernel<<<wave_grid(N, 256, 8), 256, 0, RN_STREAM>>>(indices, N, coords);
    return finish_launch("rn_morton3D_invert");
}

extern "C" int rn_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream) {
    return rn_packbits_min(grid, N, density_thresh, nullptr, bitfield, stream);
}

extern "C" int rn_packbits_min(const float* grid, uint32_t N, float density_thresh, const float* mean_density, uint8_t* bitfield, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(grid && bitfield, "null pointer");
    const uint32_t vec_ok = ((uintptr_t)grid % 16 == 0) && ((uintptr_t)bitfield % 4 == 0);
    const uint32_t work = vec_ok ? (N + 3) / 4 : N;
    packbits_kernel<<<wave_grid(work, 256, 8), 256, 0, RN_STREAM>>>(grid, N, density_thresh, mean_density, bitfield, vec_ok);
    return finish_launch("rn_packbits");
}

extern "C" uint32_t rn_occupancy_merge_workspace_bytes(void) { return RN_NUM_SMS * 8 * (uint32_t)sizeof(double) + 16; }

extern "C" int rn_occupancy_merge(float* grid, const float* fresh, uint32_t n, float decay, void* workspace, float* mean_out, void* stream) {
    RN_REQUIRE(grid && fresh && workspace && mean_out && n >= 1, "null pointer");
    RN_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
    double* partials = (double*)workspace;
    uint32_t* ticket = (uint32_t*)((uint8_t*)workspace + RN_NUM_SMS * 8 * sizeof(double));   // zero before the first call; the kernel re-arms it
    occupancy_merge_kernel<<<wave_grid(n, 256 * 4, 8), 256, 0, RN_STREAM>>>(grid, fresh, n, decay, partials, ticket, mean_out);
    return finish_launch("rn_occupancy_merge");
}

extern "C" int rn_morton3D_dilation(const float* grid, uint32_t C, uint32_t H, float* grid_dilation, void* stream) {
    if (C == 0 || H == 0) return RN_OK;
    RN_REQUIRE(grid && grid_dilation, "null pointer");
    RN_REQUIRE(H <= 1024, "H must be <= 1024 (10-bit Morton coordinates)");
    const uint64_t total = (uint64_t)C * H * H * H;
    RN_REQUIRE(total < (1ull << 32), "C*H^3 must fit 32 bits");
    dilation_kernel<<<wave_grid(total, 256, 8), 256, 0, RN_STREAM>>>(grid, C, H, grid_dilation);
    return finish_launch("rn_morton3D_dilation");
}

extern "C" int rn_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                   float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                   const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                   int32_t* rays, int32_t* counter, const float* noises, void* stream) {
    return rn_march_rays_train_budget(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nullptr, nears, fars, xyzs, dirs,
                                      deltas, rays, counter, noises, stream);
}

extern "C" int rn_march_rays_train_budget(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                          float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                          const int32_t* budget, const float* nears, const float* fars, float* xyzs, float* dirs,
                                          float* deltas, int32_t* rays, int32_t* counter, const float* noises, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(rays_o && rays_d && grid && nears && fars && rays && counter && noises, "null pointer");
    RN_REQUIRE(M == 0 || (xyzs && dirs && deltas), "null sample buffers");
    RN_REQUIRE(C >= 1 && C <= 16 && H >= 1 && max_steps >= 1, "bad C/H/max_steps");
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    march_rays_train_kernel<<<div_up(N, (uint32_t)TRAIN_THREADS), TRAIN_THREADS, 0, RN_STREAM>>>(
        rays_o, rays_d, p, max_steps, N, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises, budget);
    return finish_launch("rn_march_rays_train");
}

extern "C" int rn_march_rays_train_backward(const float* grad_xyzs, const float* grad_dirs, const int32_t* rays,
                                            const float* deltas, uint32_t N, uint32_t M, float* grad_rays_o,
                                            float* grad_rays_d, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(grad_xyzs && grad_dirs && rays && deltas && grad_rays_o && grad_rays_d, "null pointer");
    march_rays_train_backward_kernel<<<div_up(N, 256u), 256, 0, RN_STREAM>>>(grad_xyzs, grad_dirs, rays, deltas, N, M,
                                                                             grad_rays_o, grad_rays_d);
    return finish_launch("rn_march_rays_train_backward");
}

extern "C" int rn_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ambient,
                                               const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                               float T_thresh, float* weights_sum, float* ambient_sum, float* depth,
                                               float* image, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(rays && weights_sum && ambient_sum && depth && image, "null pointer");
    RN_REQUIRE(M == 0 || (sigmas && rgbs && ambient && deltas), "null sample buffers");
    composite_train_fwd_kernel<<<div_up(N, 256u), 256, 0, RN_STREAM>>>(sigmas, rgbs, ambient, deltas, rays, M, N,
                                                                       T_thresh, weights_sum, ambient_sum, depth, image);
    return finish_launch("rn_composite_rays_train_forward");
}

extern "C" int rn_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_ambient_sum,
                                                const float* grad_image, const float* sigmas, const float* rgbs,
                                                const float* ambient, const float* deltas, const int32_t* rays,
                                                const float* weights_sum, const float* ambient_sum, const float* image,
                                                uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                                float* grad_rgbs, float* grad_ambient, void* stream) {
    (void)ambient; (void)ambient_sum;
    if (N == 0) return RN_OK;
    RN_REQUIRE(grad_weights_sum && grad_ambient_sum && grad_image && rays && weights_sum && image, "null pointer");
    RN_REQUIRE(M == 0 || (sigmas && rgbs && deltas && grad_sigmas && grad_rgbs && grad_ambient), "null sample buffers");
    composite_train_bwd_kernel<<<div_up(N, 256u), 256, 0, RN_STREAM>>>(grad_weights_sum, grad_ambient_sum, grad_image,
                                                                       sigmas, rgbs, deltas, rays, weights_sum, image, M,
                                                                       N, T_thresh, grad_sigmas, grad_rgbs, grad_ambient);
    return finish_launch("rn_composite_rays_train_backward");
}

extern "C" int rn_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                             const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                             uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                             float* xyzs, float* dirs, float* deltas, const float* noises, void* stream) {
    (void)nears;
    if (n_alive == 0) return RN_OK;
    RN_REQUIRE(rays_alive && rays_t && rays_o && rays_d && grid && fars && xyzs && dirs && deltas && noises,
               "null pointer");
    RN_REQUIRE(C >= 1 && C <= 16 && H >= 1 && max_steps >= 1, "bad C/H/max_steps");
    const MarchParams p = make_march_params(bound, dt_gamma, max_steps, C, H, grid);
    march_rays_kernel<<<div_up(n_alive, 128u), 128, 0, RN_STREAM>>>(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d,
                                                                    p, fars, xyzs, dirs, deltas, noises);
    return finish_launch("rn_march_rays");
}

extern "C" int rn_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                                 const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                                 float* depth, float* image, void* stream) {
    if (n_alive == 0) return RN_OK;
    RN_REQUIRE(rays_alive && rays_t && sigmas && rgbs && deltas && weights_sum && depth && image, "null pointer");
    composite_rays_kernel<<<div_up(n_alive, 128u), 128, 0, RN_STREAM>>>(n_alive, n_step, T_thresh, rays_alive, rays_t,
                                                                        sigmas, rgbs, deltas, weights_sum, depth, image);
    return finish_launch("rn_composite_rays");
}
