// mlp_tile.cuh -- building blocks shared by the fused-MLP kernels (head_eval.cu, torso_eval.cu):
//   mma_stage            publish the group's operand rows, one thread issues the tcgen05 MMAs, everybody waits on the mbarrier
//   epilogue_to_operand  TMEM row -> (ReLU) -> fp16 -> next layer's A operand row in shared memory.  There is no bias add:
//                        the per-frame constant terms enter the accumulator through the MMA itself (a K = 16 slab of a
//                        constant "ones" operand against [hi(b) | lo(b)] rows), which took ~15% of the kernel's
//                        instructions out of the epilogues
//   fast_encode<D>       one point through the 16 levels of a tiled, linearly interpolated 2-feature fp16 grid
// mma_stage / epilogue are deliberately __noinline__: the kernel body has to stay inside the instruction cache (the first,
// fully inlined version was 179 KB of SASS and spent 47% of its issue slots waiting for instructions).
#pragma once
#include "umma.cuh"
#include "gridencoder_impl.cuh"

namespace rn {
namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// fp32 pair -> packed fp16 with ReLU in the conversion itself (cvt.rn.relu.f16x2.f32): relu(round(x)) == round(relu(x)), one
// instruction instead of F2FP + HMNMX2 (the epilogues were 14 % of head_eval's instructions)
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));    // upper half <- first source
    return d;
}
__device__ __forceinline__ uint32_t relu2(uint32_t v) {
    const __half2 h = __hmax2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&h);
}

// 32 accumulator columns [col0 + 32c, +32) per step, c < nch32: all TMEM loads are issued before the single wait.
// relu(round(x)) == round(relu(x)), so ReLU runs on packed halves.
template <int NCH32>
static __device__ __noinline__ void epilogue_to_operand(uint32_t tmem_row, uint32_t col0, bool relu, uint8_t* dst, uint32_t row,
                                                        uint32_t Kdst, uint32_t dcol0) {
    uint32_t v[NCH32][32];
#pragma unroll
    for (int c = 0; c < NCH32; ++c) umma::tmem_ld32(tmem_row + col0 + 32 * c, v[c]);
    umma::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NCH32; ++c) {
        uint32_t h[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            h[j] = relu ? pack2_relu(__uint_as_float(v[c][2 * j]), __uint_as_float(v[c][2 * j + 1]))
                        : pack2(__uint_as_float(v[c][2 * j]), __uint_as_float(v[c][2 * j + 1]));
#pragma unroll
        for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(dst + umma::il_offset(row, dcol0 + 32 * c + 8 * q, Kdst)) =
                make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
    }
}

// publish this group's freshly written operand rows, let thread 0 issue D = A0*B0^T (+ A1*B1^T), wait for completion.
// ka0 = first A column of the K0-slab inside an operand whose rows are Ka0 halves long.
// (a2, b2): optional K = 16 constant slab (ones operand x bias rows), 0 = none.
static __device__ __noinline__ void mma_stage(uint32_t tmem_acc, uint32_t a0, uint32_t Ka0, uint32_t ka0, uint32_t b0, uint32_t K0,
                                              uint32_t a1, uint32_t b1, uint32_t K1, uint32_t a2, uint32_t b2, uint32_t N, uint64_t* mbar,
                                              uint32_t& phase, uint32_t bar_id, uint32_t t) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    umma::group_sync(bar_id, 128);
    if (t == 0) {
        umma::fence_after_sync();
        umma::gemm_issue(tmem_acc, a0, b0, Ka0, K0, ka0, K0, N, false);
        if (K1) umma::gemm_issue(tmem_acc, a1, b1, K1, K1, 0, K1, N, true);
        if (a2) umma::gemm_issue(tmem_acc, a2, b2, 16, 16, 0, 16, N, true);
        umma::commit(mbar);
    }
    umma::mbar_wait(mbar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}

// ---- specialised grid level: tiled indexing, linear interpolation, align_corners = false, 2 fp16 features -------------
// The packed fp16 copy of the table has 8-byte rows [features of the cell | features of its +x neighbour] (the neighbour of a
// wrapping level is the wrapped one), so the two x-corners of every pair come from ONE load.
// row(corner) = ((base + corner_delta) & mask) | or_off inside that copy:
//   dense level   (index < size, no wrap): base includes the level's first row, mask = ~0, or_off = 0
//   capped level  (power-of-two size):     mask = size - 1, or_off = first row, which must be a multiple of size
// -- one logic op per corner instead of an and plus a 64-bit add (6 -> 3-4 instructions per gathered corner).
struct FastLevel {
    float scale;        // exp2f(l*S)*H - 1
    uint32_t s1, s2;    // strides of dimensions 1, 2 (0 once the running stride exceeded the level size: gridencoder.cu:72)
    uint32_t mask, or_off, base_add;
};

// false if the level cannot be expressed (generic modulo, hashed, misaligned packing) -- the host checks supported() first
__device__ __forceinline__ bool make_fast_level(FastLevel& f, const grid::LevelMeta& m, uint32_t first_row) {
    f.scale = m.scale;
    f.s1 = m.stride[1];
    f.s2 = m.stride[2];
    const uint32_t wrap = m.mode >> 1;
    if (wrap == 0) { f.mask = 0xffffffffu; f.or_off = 0; f.base_add = first_row; }
    else { f.mask = m.size - 1; f.or_off = first_row; f.base_add = 0; }
    return (m.mode & 1u) == 0 && wrap != 2 && (wrap == 0 || (first_row & (m.size - 1)) == 0);
}

// acc += w * g on both features, accumulated in fp32 and rounded to fp16 once per level.  (grid_forward_kernel<half>
// reproduces c10::Half's per-corner rounding bit for bit; here the 8 intermediate roundings are dropped -- two
// instructions fewer per corner, and the result is closer to the exact interpolation than the reference's.)
__device__ __forceinline__ void accum2(float2& acc, float w, uint32_t g) {
    const __half2 gv = *reinterpret_cast<const __half2*>(&g);
    acc.x = __fmaf_rn(w, __low2float(gv), acc.x);
    acc.y = __fmaf_rn(w, __high2float(gv), acc.y);
}

// One point through all 16 levels; writes its 32 halfs as 4 x 16 B into row `row` of an interleaved [128 x Kdst] operand at
// column dcol0.  Same cell / weight arithmetic as grid_forward_kernel<half, D, 2> (identical cells and corner weights), fp32
// accumulation (see accum2); the index math is specialised: base = p0 + p1*s1 + p2*s2,
// corner = (base + dx + dy*s1 + dz*s2) & mask.
// LB = levels per batch: all gathers of a batch are issued before the first interpolation, so a point costs 16 / LB dependent memory
// round trips.  Measured (round 2): LB = 8 for the 2-D grids changes nothing (tile 31 605 -> 31 106 cycles, registers 107 -> 126): the
// encode phases are not waiting on gather round trips but on their own dependent instruction chains with 4 warps per scheduler.
// Also measured and not kept: a bilinear fast path for the 3-D levels whose z stride is 0 (levels 9-15: both z-corners are one row, so
// 2 loads + 8 FMAs instead of 4 + 16; taken per 4-level batch, registers unchanged): head_eval 52.3 -> 53.2 us per launch, i.e. nothing.
template <int D, int LB = 4>
static __device__ __noinline__ void fast_encode(const float (&x)[D], const uint2* __restrict__ table64, const FastLevel* __restrict__ lv,
                                                uint8_t* dst, uint32_t row, uint32_t Kdst, uint32_t dcol0) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = true;
#pragma unroll 1
    for (int l0 = 0; l0 < 16; l0 += LB) {
        uint32_t g[LB][1 << D];
        float fr[LB][D];
#pragma unroll
        for (int j = 0; j < LB; ++j) {  // all gathers of the batch first ...
            const FastLevel L = lv[l0 + j];
            uint32_t pg[D];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const float pos = __fmaf_rn(x[d], L.scale, 0.5f);
                const float fl = floorf(pos);
                pg[d] = (uint32_t)fl;
                fr[j][d] = pos - (float)pg[d];
            }
            const uint2* __restrict__ tb = table64;
            if (oob) {
#pragma unroll
                for (int k = 0; k < (1 << D); ++k) g[j][k] = 0u;
            } else if constexpr (D == 3) {
                const uint32_t b00 = pg[0] + pg[1] * L.s1 + pg[2] * L.s2 + L.base_add, b10 = b00 + L.s1, b01 = b00 + L.s2, b11 = b10 + L.s2;
                // a row of the packed table holds the cell's features AND those of its +x neighbour (8 bytes): one load per x-pair
                const uint2 q0 = __ldg(tb + ((b00 & L.mask) | L.or_off)), q1 = __ldg(tb + ((b10 & L.mask) | L.or_off));
                const uint2 q2 = __ldg(tb + ((b01 & L.mask) | L.or_off)), q3 = __ldg(tb + ((b11 & L.mask) | L.or_off));
                g[j][0] = q0.x; g[j][1] = q0.y; g[j][2] = q1.x; g[j][3] = q1.y;
                g[j][4] = q2.x; g[j][5] = q2.y; g[j][6] = q3.x; g[j][7] = q3.y;
            } else {
                const uint32_t b0 = pg[0] + pg[1] * L.s1 + L.base_add, b1 = b0 + L.s1;
                const uint2 q0 = __ldg(tb + ((b0 & L.mask) | L.or_off)), q1 = __ldg(tb + ((b1 & L.mask) | L.or_off));
                g[j][0] = q0.x; g[j][1] = q0.y; g[j][2] = q1.x; g[j][3] = q1.y;
            }
        }
        uint32_t packed[LB];
#pragma unroll
        for (int j = 0; j < LB; ++j) {  // ... then the interpolation, corner order = bit d of k selects dimension d
            float2 acc = make_float2(0.f, 0.f);
            if (!oob) {
                const float x0 = 1.0f - fr[j][0], x1 = fr[j][0], y0 = 1.0f - fr[j][1], y1 = fr[j][1];
                if constexpr (D == 3) {
                    const float z0 = 1.0f - fr[j][2], z1 = fr[j][2];
                    const float w00 = __fmul_rn(x0, y0), w10 = __fmul_rn(x1, y0), w01 = __fmul_rn(x0, y1), w11 = __fmul_rn(x1, y1);
                    accum2(acc, __fmul_rn(w00, z0), g[j][0]); accum2(acc, __fmul_rn(w10, z0), g[j][1]);
                    accum2(acc, __fmul_rn(w01, z0), g[j][2]); accum2(acc, __fmul_rn(w11, z0), g[j][3]);
                    accum2(acc, __fmul_rn(w00, z1), g[j][4]); accum2(acc, __fmul_rn(w10, z1), g[j][5]);
                    accum2(acc, __fmul_rn(w01, z1), g[j][6]); accum2(acc, __fmul_rn(w11, z1), g[j][7]);
                } else {
                    accum2(acc, __fmul_rn(x0, y0), g[j][0]); accum2(acc, __fmul_rn(x1, y0), g[j][1]);
                    accum2(acc, __fmul_rn(x0, y1), g[j][2]); accum2(acc, __fmul_rn(x1, y1), g[j][3]);
                }
            }
            packed[j] = pack2(acc.x, acc.y);
        }
#pragma unroll
        for (int q = 0; q < LB / 4; ++q)
            *reinterpret_cast<uint4*>(dst + umma::il_offset(row, dcol0 + 2 * l0 + 8 * q, Kdst)) =
                make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
}

}  // namespace
}  // namespace rn
