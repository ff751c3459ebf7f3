// march.cuh -- the occupancy-grid DDA shared by the training / inference marchers (raymarching.cu) and the fused frame
// renderer (frame.cu).  Every floating-point operation is spelled explicitly (__fmaf_rn / __fmul_rn / __fadd_rn) in the
// form nvcc contracts the reference expressions to (raymarching/src/raymarching.cu:386-441, checked against the sm_100a
// SASS of the reference build), so emitted samples are bit-identical to the reference's.
#pragma once
#include "common.cuh"
#include "occ_pack.cuh"

namespace rn {

struct MarchParams {
    float bound, dt_gamma, dt_min, dt_max;
    float Hf, halfH, Hm1f, rH, H3f, rbound;
    int Cm1;
    const uint8_t* __restrict__ grid;
    // optional: the occupied boxes of the bitfield as linear bit arrays staged in SHARED memory (occ_pack.cuh); null = probe `grid`
    const uint32_t* occ_bits;
    const OccLevel* occ_lv;
};

__host__ __device__ inline MarchParams make_march_params(float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                                                         uint32_t H, const uint8_t* grid) {
    MarchParams p;
    p.bound = bound;
    p.dt_gamma = dt_gamma;
    // dt_max = 2*sqrt3 * 2^(C-1) / H ; dt_min = min(dt_max, 2*sqrt3 / max_steps)      (raymarching.cu:386-387)
    p.dt_max = ((float)(1 << (C - 1)) * 3.4641015529632568359f) / (float)H;
    const float q = 3.4641015529632568359f / (float)max_steps;
    p.dt_min = q < p.dt_max ? q : p.dt_max;  // fminf
    p.Hf = (float)H;
    p.halfH = 0.5f * (float)H;
    p.Hm1f = (float)(H - 1);
    p.rH = 1.0f / (float)H;
    p.H3f = (float)(H * H * H);
    p.rbound = 1.0f / bound;  // IEEE division, same value the device computes for 1 / mip_bound when mip_bound == bound
    p.Cm1 = (int)C - 1;
    p.grid = grid;
    p.occ_bits = nullptr;
    p.occ_lv = nullptr;
    return p;
}

struct Ray {
    float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, hsx, hsy, hsz;
    __device__ __forceinline__ void load(const float* __restrict__ o, const float* __restrict__ d) {
        ox = __ldg(o); oy = __ldg(o + 1); oz = __ldg(o + 2);
        dx = __ldg(d); dy = __ldg(d + 1); dz = __ldg(d + 2);
        rdx = 1 / dx; rdy = 1 / dy; rdz = 1 / dz;
        hsx = copysignf(1.0f, dx); hsy = copysignf(1.0f, dy); hsz = copysignf(1.0f, dz);
    }
};

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

__device__ __forceinline__ float step_size(const MarchParams& p, float t) {
    return clampf(__fmul_rn(t, p.dt_gamma), p.dt_min, p.dt_max);
}

// clamp(frexp exponent of mx, 0, C-1): [0,0.5) -> <=-1 -> 0, [0.5,1) -> 0, [1,2) -> 1 ...        (raymarching.cu:42-54)
// mx is finite and >= 0 here, so the exponent comes straight from the bit pattern (zero / denormals have a biased exponent
// of 0 and land on level 0 exactly as frexpf's e <= 0 does); libdevice's frexpf costs ~30 instructions, twice per step.
__device__ __forceinline__ int cascade_of(float mx, int Cm1) {
    const int e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 126;
    return min(Cm1, max(0, e));
}

// Probe the occupancy grid at parameter t.  Returns true when the cell is occupied (x,y,z,dt are then the sample);
// otherwise advances t past the current voxel exactly as the reference's skip loop does.
__device__ __forceinline__ bool march_probe(const MarchParams& p, const Ray& r, float& t, float& x, float& y, float& z,
                                            float& dt) {
    x = clampf(__fmaf_rn(r.dx, t, r.ox), -p.bound, p.bound);
    y = clampf(__fmaf_rn(r.dy, t, r.oy), -p.bound, p.bound);
    z = clampf(__fmaf_rn(r.dz, t, r.oz), -p.bound, p.bound);
    dt = step_size(p, t);

    const int lvl_pos = cascade_of(fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z))), p.Cm1);
    const int lvl_dt = cascade_of(__fmul_rn(__fmul_rn(dt, p.Hf), 0.5f), p.Cm1);
    const int level = max(lvl_pos, lvl_dt);

    const float pow2 = __int_as_float((127 + level) << 23);
    const float mip_bound = fminf(pow2, p.bound);
    // 1 / mip_bound without the division sequence: exact for the power of two, precomputed (IEEE) for the scene bound
    const float mip_rbound = pow2 <= p.bound ? __int_as_float((127 - level) << 23) : p.rbound;

    const int nx = (int)clampf(__fmul_rn(__fmaf_rn(x, mip_rbound, 1.0f), p.halfH), 0.0f, p.Hm1f);
    const int ny = (int)clampf(__fmul_rn(__fmaf_rn(y, mip_rbound, 1.0f), p.halfH), 0.0f, p.Hm1f);
    const int nz = (int)clampf(__fmul_rn(__fmaf_rn(z, mip_rbound, 1.0f), p.halfH), 0.0f, p.Hm1f);

    bool occ;
    if (p.occ_bits) {
        // the same cell, looked up in the box-packed copy in shared memory: outside the box of the occupied cells every bit is zero
        const OccLevel& L = p.occ_lv[level];
        const uint32_t bx = (uint32_t)(nx - L.lo[0]), by = (uint32_t)(ny - L.lo[1]), bz = (uint32_t)(nz - L.lo[2]);
        occ = false;
        if (bx < (uint32_t)L.dim[0] && by < (uint32_t)L.dim[1] && bz < (uint32_t)L.dim[2]) {
            const uint32_t i = (bz * (uint32_t)L.dim[1] + by) * (uint32_t)L.dim[0] + bx;
            occ = (p.occ_bits[(uint32_t)L.word_off + (i >> 5)] >> (i & 31u)) & 1u;
        }
    } else {
        // level * H^3 + morton, evaluated in fp32 as the reference does (H3 is a float there, raymarching.cu:380,419)
        const uint32_t index = (uint32_t)__fmaf_rn((float)level, p.H3f, (float)morton_encode(nx, ny, nz));
        occ = (__ldg(p.grid + (index >> 3)) >> (index & 7u)) & 1u;
    }
    if (occ) return true;

    // distance to the next voxel boundary along each axis              (raymarching.cu:431-439)
    const float ax = __fmaf_rn(r.hsx, 0.5f, (float)nx + 0.5f);
    const float ay = __fmaf_rn(r.hsy, 0.5f, (float)ny + 0.5f);
    const float az = __fmaf_rn(r.hsz, 0.5f, (float)nz + 0.5f);
    const float tx = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(ax, p.rH), 2.0f, -1.0f), -x), r.rdx);
    const float ty = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(ay, p.rH), 2.0f, -1.0f), -y), r.rdy);
    const float tz = __fmul_rn(__fmaf_rn(mip_bound, __fmaf_rn(__fmul_rn(az, p.rH), 2.0f, -1.0f), -z), r.rdz);
    const float tt = __fadd_rn(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
    do {
        t = __fadd_rn(step_size(p, t), t);
    } while (t < tt);
    return false;
}


}  // namespace rn
