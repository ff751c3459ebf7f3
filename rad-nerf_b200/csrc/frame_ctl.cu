// frame_ctl.cu -- the non-tensor kernels of the fused inference frame: per-ray init, compacting march, compositing with
// survivor compaction and the device-side loop controller, torso mask, final blend.  See frame.cuh for the pipeline.
//
// Replaces, for inference, the host loop of NeRFRenderer.run_cuda (nerf/renderer.py:229-262) and its per-iteration
// kernels + torch glue: march_rays / composite_rays (raymarching.cu:827-1029), `rays_alive[rays_alive >= 0]` (a host
// sync per iteration, renderer.py:258), three zero-fills per march call (raymarching.py:385-387).
#include "frame.cuh"
#include "march.cuh"
#include <float.h>

namespace rn {

constexpr int CTL_THREADS = 128;


__host__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(FrameWorkspace& w, uint8_t* base, uint32_t N) {
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += align256(bytes); return p; };
    w.ctl = (FrameCtl*)take(sizeof(FrameCtl) * (FRAME_MAX_ITERS + 1));
    w.cur = (FrameCur*)take(sizeof(FrameCur));
    w.alive[0] = (int32_t*)take(4ull * N);
    w.alive[1] = (int32_t*)take(4ull * N);
    w.rays_t = (float*)take(4ull * N);
    w.ray_cnt = (uint32_t*)take(4ull * N);
    w.sample_idx = (uint32_t*)take(4ull * 8 * N);
    w.samples = (float4*)take(16ull * (N + EVAL_TILE));
    w.deltas = (float2*)take(8ull * (N + EVAL_TILE));
    w.evals = (float4*)take(16ull * (N + EVAL_TILE));
    w.torso_pix = (int32_t*)take(4ull * N);
    w.torso_out = (float4*)take(16ull * (N + EVAL_TILE));
    w.tmisc = (uint32_t*)take(32);
    w.stats = (uint32_t*)take(32);   // keep last: radnerf_b200/frame.py reads it at workspace_bytes - 256
    return off;
}

namespace {

// block-wide exclusive scan over CTL_THREADS values; returns the exclusive prefix, total in `total`
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t& total, uint32_t* warp_sums) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < CTL_THREADS / 32; ++w) {
        const uint32_t s = warp_sums[w];
        if ((uint32_t)w < warp) base += s;
        tot += s;
    }
    total = tot;
    __syncthreads();  // warp_sums may be reused
    return base + incl - v;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
frame_init_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ aabb,
                  const float* __restrict__ occ_aabb, uint32_t N, float min_near, uint32_t max_steps, float* __restrict__ nears, float* __restrict__ fars,
                  float* __restrict__ rays_t, int32_t* __restrict__ alive0, float* __restrict__ weights_sum,
                  float* __restrict__ depth, float* __restrict__ image, FrameCur* __restrict__ cur) {
    __shared__ float box[6], occ[6];
    if (threadIdx.x < 6) {
        box[threadIdx.x] = aabb[threadIdx.x];
        occ[threadIdx.x] = occ_aabb ? occ_aabb[threadIdx.x] : (threadIdx.x < 3 ? -FLT_MAX : FLT_MAX);
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        // the history ctl[] and *cur were zeroed by the memset nodes that precede this kernel
        if (threadIdx.x == 0) {
            cur->c.n_alive = N;
            cur->c.n_step = 1;   // clamp(N // N, 1, 8)
            cur->c.step = 0;
            cur->c.done = (max_steps == 0 || N == 0) ? 1u : 0u;
        }
    }
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float* o = rays_o + (size_t)n * 3;
        const float* d = rays_d + (size_t)n * 3;
        // slab test, identical arithmetic to rn_near_far_from_aabb (raymarching.cu:91-145)
        const float ox = __ldg(o), oy = __ldg(o + 1), oz = __ldg(o + 2);
        const float rdx = 1 / __ldg(d), rdy = 1 / __ldg(d + 1), rdz = 1 / __ldg(d + 2);
        float near = __fmul_rn(box[0] - ox, rdx), far = __fmul_rn(box[3] - ox, rdx);
        if (near > far) { float c = near; near = far; far = c; }
        float ny = __fmul_rn(box[1] - oy, rdy), fy = __fmul_rn(box[4] - oy, rdy);
        if (ny > fy) { float c = ny; ny = fy; fy = c; }
        bool miss = near > fy || ny > far;
        if (!miss) {
            if (ny > near) near = ny;
            if (fy < far) far = fy;
            float nz = __fmul_rn(box[2] - oz, rdz), fz = __fmul_rn(box[5] - oz, rdz);
            if (nz > fz) { float c = nz; nz = fz; fz = c; }
            miss = near > fz || nz > far;
            if (!miss) {
                if (nz > near) near = nz;
                if (fz < far) far = fz;
                if (near < min_near) near = min_near;
            }
        }
        if (miss) near = far = FLT_MAX;
        // Rays that cannot meet an occupied cell are not marched: start them at `far`.  Every position the marcher probes
        // lies on the segment [near, far] of the ray, so a segment that misses the (inflated) bounding box of the occupied
        // cells emits no sample -- exactly what marching it through empty space would find.  The test only prunes on a
        // proven miss: any NaN (0 * inf on a slab plane) compares false and the ray is marched as usual.
        float t0 = near, t1 = far;
        {
            const float ax = (occ[0] - ox) * rdx, bx = (occ[3] - ox) * rdx;
            const float ay = (occ[1] - oy) * rdy, by = (occ[4] - oy) * rdy;
            const float az = (occ[2] - oz) * rdz, bz = (occ[5] - oz) * rdz;
            t0 = fmaxf(t0, fmaxf(fminf(ax, bx), fmaxf(fminf(ay, by), fminf(az, bz))));
            t1 = fminf(t1, fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz))));
        }
        const bool prune = t0 > t1;
        nears[n] = near;
        fars[n] = far;
        rays_t[n] = prune ? far : near;
        alive0[n] = (int32_t)n;
        weights_sum[n] = 0.f;
        depth[n] = 0.f;
        image[(size_t)n * 3] = 0.f; image[(size_t)n * 3 + 1] = 0.f; image[(size_t)n * 3 + 2] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// march_compact: one thread per alive ray, <= 8 samples staged in shared memory, compacted write.
__global__ void __launch_bounds__(CTL_THREADS)
march_compact_kernel(FrameCur* cur, const int32_t* __restrict__ alive0,
                     const int32_t* __restrict__ alive1, const float* __restrict__ rays_t, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                     const float* __restrict__ fars, MarchParams p, const float* __restrict__ noises,
                     uint32_t* __restrict__ ray_cnt, uint32_t* __restrict__ sample_idx, uint32_t N, float4* __restrict__ samples,
                     float2* __restrict__ deltas, const OccPack* __restrict__ occ, uint32_t occ_words, const float* __restrict__ occ_aabb) {
    extern __shared__ __align__(16) uint32_t s_occ[];      // occ_words words: the box-packed occupancy bits (occ_pack.cu)
    __shared__ OccLevel s_lv[OCC_MAX_LEVELS];
    __shared__ float4 s_xyz[8][CTL_THREADS];
    __shared__ float2 s_dt[8][CTL_THREADS];
    __shared__ unsigned long long s_wsum[CTL_THREADS / 32];
    __shared__ uint32_t s_base;

    const uint4 head = *reinterpret_cast<const uint4*>(&cur->c);   // n_alive, n_step, step, done
    FrameCtl* ctl_rw = &cur->c;
    if (head.w) return;
    const uint32_t n_alive = head.x, n_step = head.y, it = cur->it;
    if (blockIdx.x * CTL_THREADS >= n_alive) return;
    const uint32_t j = blockIdx.x * CTL_THREADS + threadIdx.x;
    const int32_t* __restrict__ alive = (it & 1u) ? alive1 : alive0;
    if (it != 0) noises = nullptr;   // the per-ray start offset is applied once, by the first march of the frame

    // stage the occupancy bits: every DDA probe below is then a shared-memory read instead of a dependent L2 round trip.
    // (CTA-uniform: the launch was sized for `occ_words`, the pack says whether it was written and still has that size)
    if (occ && occ_words && occ->usable && (uint32_t)occ->total_words <= occ_words) {
        const uint32_t n16 = ((uint32_t)occ->total_words + 3u) / 4u;
        const uint4* __restrict__ src = reinterpret_cast<const uint4*>(occ->bits);
        for (uint32_t k = threadIdx.x; k < n16; k += CTL_THREADS) reinterpret_cast<uint4*>(s_occ)[k] = __ldg(src + k);
        if (threadIdx.x < OCC_MAX_LEVELS) s_lv[threadIdx.x] = occ->lv[threadIdx.x];
        __syncthreads();
        p.occ_bits = s_occ;
        p.occ_lv = s_lv;
    }

    uint32_t cnt = 0;
    int32_t ray = 0;
    if (j < n_alive) {
        ray = __ldg(alive + j);
        Ray r;
        r.load(rays_o + (size_t)ray * 3, rays_d + (size_t)ray * 3);
        float t = __ldg(rays_t + ray);
        float far = __ldg(fars + ray);
        if (occ_aabb) {
            // Past the exit of the (one-cell inflated) box of all occupied cells a ray can only probe empty cells: stop there instead of
            // stepping cell by cell to the far plane (up to ~60 dependent probes of ~100 instructions each, the bulk of this kernel's
            // latency for rays that leave the head).  Same samples: everything between the exit and `far` emits nothing.  A NaN slab
            // (0 * inf) compares false in fminf's favour of the other operand, i.e. falls back to `far`.
            const float ax = (__ldg(occ_aabb + 0) - r.ox) * r.rdx, bx = (__ldg(occ_aabb + 3) - r.ox) * r.rdx;
            const float ay = (__ldg(occ_aabb + 1) - r.oy) * r.rdy, by = (__ldg(occ_aabb + 4) - r.oy) * r.rdy;
            const float az = (__ldg(occ_aabb + 2) - r.oz) * r.rdz, bz = (__ldg(occ_aabb + 5) - r.oz) * r.rdz;
            const float t_exit = fminf(fmaxf(ax, bx), fminf(fmaxf(ay, by), fmaxf(az, bz)));
            far = fminf(far, t_exit);
        }
        const float noise = noises ? __ldg(noises + j) : 0.0f;
        t = __fmaf_rn(noise, step_size(p, t), t);  // raymarching.cu:873
        float x, y, z, dt;
        while (t < far && cnt < n_step) {
            if (march_probe(p, r, t, x, y, z, dt)) {
                t = __fadd_rn(t, dt);
                s_xyz[cnt][threadIdx.x] = make_float4(x, y, z, __int_as_float(ray));
                s_dt[cnt][threadIdx.x] = make_float2(dt, t);
                ++cnt;
            }
        }
    }
    // ---- k-major compaction inside the CTA's chunk: first every ray's sample 0, then every ray's sample 1, ...
    // Lanes of a warp in the network kernel then hold samples of NEIGHBOURING RAYS at the same depth index (a few grid
    // cells apart) instead of consecutive samples of one ray (~27 fine cells apart): far fewer cache lines per gather.
    // One 64-bit scan carries the eight per-k flags in 8-bit fields (a CTA has 128 rays, so a field never overflows).
    unsigned long long flags = 0ull;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) flags |= (unsigned long long)(cnt > k ? 1u : 0u) << (8 * k);
    unsigned long long incl = flags;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long n = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= (uint32_t)off) incl += n;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    unsigned long long wbase = 0ull, tot = 0ull;
#pragma unroll
    for (int w = 0; w < CTL_THREADS / 32; ++w) {
        const unsigned long long v = s_wsum[w];
        if ((uint32_t)w < warp) wbase += v;
        tot += v;
    }
    const unsigned long long excl = wbase + incl - flags;  // per-k exclusive rank of this ray among the CTA's rays
    uint32_t total = 0;
#pragma unroll
    for (uint32_t k = 0; k < 8; ++k) total += (uint32_t)(tot >> (8 * k)) & 0xffu;
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&ctl_rw->n_samples, total) : 0u;
    __syncthreads();
    if (j < n_alive) {
        ray_cnt[j] = cnt;
        uint32_t kbase = s_base;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            if (k < cnt) {
                const uint32_t pos = kbase + ((uint32_t)(excl >> (8 * k)) & 0xffu);
                sample_idx[(size_t)k * N + j] = pos;
                samples[pos] = s_xyz[k][threadIdx.x];
                deltas[pos] = s_dt[k][threadIdx.x];
            }
            kbase += (uint32_t)(tot >> (8 * k)) & 0xffu;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// composite_compact: kernel_composite_rays semantics (raymarching.cu:942-1029) on the compacted sample list, then
// survivor compaction; the last CTA plays the host loop (renderer.py:241-262) and publishes ctl[it + 1].
__global__ void __launch_bounds__(CTL_THREADS)
composite_compact_kernel(FrameCur* cur, FrameCtl* __restrict__ history, uint32_t* __restrict__ stats, unsigned long long cond_handle,
                         int32_t* __restrict__ alive0, int32_t* __restrict__ alive1, float* __restrict__ rays_t,
                         const uint32_t* __restrict__ ray_cnt, const uint32_t* __restrict__ sample_idx, const float2* __restrict__ deltas,
                         const float4* __restrict__ evals,
                         float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image, float T_thresh,
                         uint32_t N, uint32_t max_steps) {
    __shared__ uint32_t warp_sums[CTL_THREADS / 32];
    __shared__ uint32_t s_base;

    const uint4 head = *reinterpret_cast<const uint4*>(&cur->c);   // n_alive, n_step, step, done
    FrameCtl* ctl_rw = &cur->c;
    if (head.w) return;
    const uint32_t n_alive = head.x, n_step = head.y, it = cur->it;
    if (blockIdx.x * CTL_THREADS >= n_alive) return;
    const uint32_t j = blockIdx.x * CTL_THREADS + threadIdx.x;
    const int32_t* __restrict__ alive_in = (it & 1u) ? alive1 : alive0;
    int32_t* __restrict__ alive_out = (it & 1u) ? alive0 : alive1;

    uint32_t survive = 0;
    int32_t ray = 0;
    if (j < n_alive) {
        ray = __ldg(alive_in + j);
        const uint32_t cnt = ray_cnt[j];
        float ws = weights_sum[ray], d = depth[ray];
        float r = image[(size_t)ray * 3], g = image[(size_t)ray * 3 + 1], b = image[(size_t)ray * 3 + 2];
        float t = 0.f;
        uint32_t step = 0;
        // all of the ray's <= 8 samples are fetched before the serial compositing recurrence: two dependent memory round trips per
        // ray instead of two per sample (the kernel is a latency chain: one thread per ray, a handful of warps per SM)
        const uint32_t n_have = min(cnt, n_step);
        uint32_t pos_k[8];
        float2 dd_k[8];
        float4 ev_k[8];
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) pos_k[k] = k < n_have ? __ldg(sample_idx + (size_t)k * N + j) : 0u;
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            if (k < n_have) { dd_k[k] = __ldg(deltas + pos_k[k]); ev_k[k] = __ldg(evals + pos_k[k]); }
            else { dd_k[k] = make_float2(0.f, 0.f); ev_k[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
        }
#pragma unroll
        for (uint32_t k = 0; k < 8; ++k) {
            if (step != k || step >= n_step) continue;   // (the loop below, unrolled so that the prefetched registers are indexed statically)
            if (step >= cnt) break;  // the marcher ran out of samples (zero-filled slot in the reference)
            const float2 dd = dd_k[k];
            const float4 e = ev_k[k];
            const float alpha = 1.0f - __expf(-e.x * dd.x);
            const float T = 1 - ws;
            const float weight = __fmul_rn(alpha, T);
            ws = __fadd_rn(ws, weight);
            t = dd.y;
            d = __fmaf_rn(weight, t, d);
            r = __fmaf_rn(weight, e.y, r);
            g = __fmaf_rn(weight, e.z, g);
            b = __fmaf_rn(weight, e.w, b);
            if (T < T_thresh) break;
            ++step;
        }
        survive = (step == n_step) ? 1u : 0u;
        if (survive) rays_t[ray] = t;
        weights_sum[ray] = ws;
        depth[ray] = d;
        image[(size_t)ray * 3] = r; image[(size_t)ray * 3 + 1] = g; image[(size_t)ray * 3 + 2] = b;
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(survive, total, warp_sums);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(&ctl_rw->next_alive, total) : 0u;
    __syncthreads();
    if (survive) alive_out[s_base + excl] = ray;

    // ---- last CTA: the loop controller
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t active = (n_alive + CTL_THREADS - 1) / CTL_THREADS;
        const uint32_t ticket = atomicAdd(&ctl_rw->blocks_done, 1u);
        if (ticket == active - 1) {
            __threadfence();
            const uint32_t next = atomicAdd(&ctl_rw->next_alive, 0u);
            const uint32_t n_samples = atomicAdd(&ctl_rw->n_samples, 0u);
            const uint32_t step = head.z + n_step;  // `step += n_step`
            const uint32_t total = ctl_rw->total_samples + n_samples;
            FrameCtl h;   // history entry of the iteration that just finished
            h.n_alive = n_alive; h.n_step = n_step; h.step = head.z; h.done = 0;
            h.n_samples = n_samples; h.next_alive = next; h.blocks_done = active; h.total_samples = total;
            history[it] = h;
            FrameCtl c;
            c.n_alive = next;
            c.n_step = next ? max(min(N / next, 8u), 1u) : 1u;  // `max(min(N // n_alive, 8), 1)`
            c.step = step;
            c.done = (step >= max_steps || next == 0) ? 1u : 0u;
            c.n_samples = 0; c.next_alive = 0; c.blocks_done = 0;
            c.total_samples = total;
            // Every CTA that works on this iteration has taken its ticket, i.e. is past its reads of *cur.  A CTA of this
            // launch that starts late may see any mix of old and new fields; n_alive never grows and `done` only rises, so
            // it still takes its early exit.
            cur->c = c;
            cur->it = it + 1;
            stats[0] += 1;
            if (cond_handle) cudaGraphSetConditional(cond_handle, c.done ? 0u : 1u);   // WHILE node: run the body again?
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// torso mask: bilinear lookup of the 2-D occupancy grid at the pixel's background coordinate (F.grid_sample with
// align_corners=True, zero padding; nerf/renderer.py:281-283), thresholded, compacted into a pixel list.
__global__ void __launch_bounds__(CTL_THREADS)
torso_mask_kernel(const float* __restrict__ bg_coords, const float* __restrict__ grid, uint32_t G, float thresh, uint32_t N,
                  int32_t* __restrict__ pix, uint32_t* __restrict__ n_torso) {
    __shared__ uint32_t warp_sums[CTL_THREADS / 32];
    __shared__ uint32_t s_base;
    const uint32_t n = blockIdx.x * CTL_THREADS + threadIdx.x;
    uint32_t hit = 0;
    if (n < N) {
        const float gx = __ldg(bg_coords + (size_t)n * 2), gy = __ldg(bg_coords + (size_t)n * 2 + 1);
        const float ix = ((gx + 1.f) / 2) * (float)(G - 1), iy = ((gy + 1.f) / 2) * (float)(G - 1);
        const float x0f = floorf(ix), y0f = floorf(iy);
        const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
        const float nw = ((float)x1 - ix) * ((float)y1 - iy), ne = (ix - (float)x0) * ((float)y1 - iy);
        const float sw = ((float)x1 - ix) * (iy - (float)y0), se = (ix - (float)x0) * (iy - (float)y0);
        auto at = [&](int x, int y) { return (x >= 0 && y >= 0 && x < (int)G && y < (int)G) ? __ldg(grid + (size_t)y * G + x) : 0.f; };
        float v = 0.f;
        v = __fmaf_rn(at(x0, y0), nw, v);
        v = __fmaf_rn(at(x1, y0), ne, v);
        v = __fmaf_rn(at(x0, y1), sw, v);
        v = __fmaf_rn(at(x1, y1), se, v);
        hit = v > thresh ? 1u : 0u;
    }
    uint32_t total;
    const uint32_t excl = block_exclusive_scan(hit, total, warp_sums);
    if (threadIdx.x == 0) s_base = total ? atomicAdd(n_torso, total) : 0u;
    __syncthreads();
    if (hit) pix[s_base + excl] = (int32_t)n;
}

// scatter the compact torso results to per-pixel (alpha, rgb) -- zero elsewhere (buffer pre-zeroed)
__global__ void __launch_bounds__(256)
torso_scatter_kernel(const int32_t* __restrict__ pix, const float4* __restrict__ torso_out, const uint32_t* __restrict__ n_torso,
                     float* __restrict__ torso_alpha, float* __restrict__ torso_color) {
    const uint32_t n = *n_torso;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const int32_t p = __ldg(pix + k);
        const float4 o = __ldg(torso_out + k);
        torso_alpha[p] = o.x;
        torso_color[(size_t)p * 3] = o.y; torso_color[(size_t)p * 3 + 1] = o.z; torso_color[(size_t)p * 3 + 2] = o.w;
    }
}

// final blend (nerf/renderer.py:299-310), one thread per pixel; every op rounded separately as torch's eager kernels do
__global__ void __launch_bounds__(256)
finalize_kernel(uint32_t N, const float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
                const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ bg_color /*[N,3] or null*/,
                float bg_scalar, const float* __restrict__ torso_alpha /*nullable*/, const float* __restrict__ torso_color,
                float* __restrict__ torso_bg_out /*nullable [N,3]*/) {
    for (uint32_t n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        const float one_m_ws = 1 - weights_sum[n];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float bg = bg_color ? bg_color[(size_t)n * 3 + c] : bg_scalar;
            if (torso_alpha) {
                const float a = torso_alpha[n];
                bg = __fadd_rn(__fmul_rn(torso_color[(size_t)n * 3 + c], a), __fmul_rn(bg, 1 - a));
                if (torso_bg_out) torso_bg_out[(size_t)n * 3 + c] = bg;
            }
            const float v = __fadd_rn(image[(size_t)n * 3 + c], __fmul_rn(one_m_ws, bg));
            image[(size_t)n * 3 + c] = fminf(fmaxf(v, 0.f), 1.f);
        }
        const float near = nears[n], far = fars[n];
        depth[n] = fmaxf(depth[n] - near, 0.f) / (far - near);
    }
}

}  // namespace

// ---- launchers used by frame.cu --------------------------------------------------------------------------------
int launch_frame_init(const float* rays_o, const float* rays_d, const float* aabb, const float* occ_aabb, uint32_t N, float min_near,
                      uint32_t max_steps, float* nears, float* fars, const FrameWorkspace& w, float* weights_sum, float* depth,
                      float* image, cudaStream_t st) {
    cudaMemsetAsync(w.ctl, 0, sizeof(FrameCtl) * (FRAME_MAX_ITERS + 1), st);
    cudaMemsetAsync(w.cur, 0, sizeof(FrameCur), st);
    frame_init_kernel<<<wave_grid(N, 256, 8), 256, 0, st>>>(rays_o, rays_d, aabb, occ_aabb, N, min_near, max_steps, nears, fars, w.rays_t,
                                                            w.alive[0], weights_sum, depth, image, w.cur);
    return finish_launch("frame_init");
}

int launch_march_compact(uint32_t N, const FrameWorkspace& w, const float* rays_o, const float* rays_d, const float* fars,
                         const MarchParams& p, const float* noises, const void* occ_pack, uint32_t occ_words, const float* occ_aabb, cudaStream_t st) {
    // dynamic shared memory = the packed occupancy bits (rounded to 16 bytes); with the 24 KiB of sample staging a CTA stays under the
    // 48 KiB that need no opt-in as long as the pack is <= 20 KiB, larger packs opt in
    const uint32_t smem = occ_pack ? ((occ_words * 4u + 15u) & ~15u) : 0u;
    if (smem > 20u * 1024u) {
        cudaError_t e = cudaFuncSetAttribute(march_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RN_OCC_PACK_MAX_BYTES);
        if (e != cudaSuccess) { set_error("march_compact: cannot reserve %u bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
    }
    march_compact_kernel<<<div_up(N, (uint32_t)CTL_THREADS), CTL_THREADS, smem, st>>>(w.cur, w.alive[0], w.alive[1], w.rays_t,
                                                                                    rays_o, rays_d, fars, p, noises, w.ray_cnt,
                                                                                    w.sample_idx, N, w.samples, w.deltas,
                                                                                    (const OccPack*)occ_pack, occ_pack ? occ_words : 0u, occ_aabb);
    return finish_launch("march_compact");
}

int launch_composite_compact(uint32_t N, uint32_t max_steps, float T_thresh, const FrameWorkspace& w, float* weights_sum,
                             float* depth, float* image, unsigned long long cond_handle, cudaStream_t st) {
    composite_compact_kernel<<<div_up(N, (uint32_t)CTL_THREADS), CTL_THREADS, 0, st>>>(
        w.cur, w.ctl, w.stats, cond_handle, w.alive[0], w.alive[1], w.rays_t, w.ray_cnt, w.sample_idx, w.deltas, w.evals, weights_sum,
        depth, image, T_thresh, N, max_steps);
    return finish_launch("composite_compact");
}

int launch_torso_mask(const float* bg_coords, const float* grid, uint32_t G, float thresh, uint32_t N, const FrameWorkspace& w,
                      cudaStream_t st) {
    torso_mask_kernel<<<div_up(N, (uint32_t)CTL_THREADS), CTL_THREADS, 0, st>>>(bg_coords, grid, G, thresh, N, w.torso_pix, w.tmisc);
    return finish_launch("torso_mask");
}

int launch_torso_scatter(uint32_t N, const FrameWorkspace& w, float* torso_alpha, float* torso_color, cudaStream_t st) {
    cudaMemsetAsync(torso_alpha, 0, 4ull * N, st);
    cudaMemsetAsync(torso_color, 0, 12ull * N, st);
    torso_scatter_kernel<<<wave_grid(N, 256, 4), 256, 0, st>>>(w.torso_pix, w.torso_out, w.tmisc, torso_alpha, torso_color);
    return finish_launch("torso_scatter");
}

int launch_finalize(uint32_t N, const float* weights_sum, float* depth, float* image, const float* nears, const float* fars,
                    const float* bg_color, float bg_scalar, const float* torso_alpha, const float* torso_color, float* torso_bg_out,
                    cudaStream_t st) {
    finalize_kernel<<<wave_grid(N, 256, 8), 256, 0, st>>>(N, weights_sum, depth, image, nears, fars, bg_color, bg_scalar, torso_alpha,
                                                          torso_color, torso_bg_out);
    return finish_launch("finalize");
}

}  // namespace rn
