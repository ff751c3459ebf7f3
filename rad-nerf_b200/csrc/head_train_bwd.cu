// head_train_bwd.cu -- backward of the fused training step of the head network (see head_train.cuh for the plan).
//
//   bwd<1>: colour + sigma nets.  in: dL/d(rgb), dL/d(sigma) per sample; out: dL/d(enc_x) (part 1), dL/d(enc_w), dW of 5 layers
//   bwd<2>: ambient net.          in: dL/d(ambient) (compositor + 2-D grid path);  out: dL/d(enc_x) (total),  dW of 3 layers
//
// CTA = BWD_GROUPS tile groups of 128 worker threads + ONE issuer warp.  A worker owns sample row t of its group's tile (TMEM lane t):
// it turns the accumulator of the previous data-gradient GEMM into the next fp16 operand row (relu mask from the saved activation
// tile, activation derivatives in fp32), then arrives on its group's `ready` mbarrier.  The issuer thread sees the group ready, issues
//      dgrad   D_work[128 x N]  = G[128 x K] W^T-blob[N x K]^T                (K-major operands, as the forward)
//      wgrad   dW[rows x cols] += G^T A   or   A^T G                           (both operands MN-major views of the tiles, umma.cuh)
// commits to the group's `done` mbarrier and prefetches the activation tile of the group's NEXT stage with a bulk copy.  All MMAs of
// the CTA come from that one thread: the weight-gradient accumulators in TMEM are shared by the groups and must see their updates in
// program order.  dW stays in TMEM for the whole persistent CTA and is written out once (per-CTA partials, summed by dw_reduce_kernel).
#include "frame.cuh"
#include "umma.cuh"
#include "mlp_tile.cuh"
#include "head_train.cuh"

namespace rn {
namespace train {
namespace {

constexpr uint32_t ACT_BYTES = 128 * 80 * 2;                       // largest saved tile (CIN)
constexpr uint32_t GRP_BYTES = 3 * ACT_BYTES;                      // ACT[0], ACT[1], GRAD
constexpr uint32_t ONES_BYTES = 128 * 16 * 2;
constexpr uint32_t ISSUER_WARP = BWD_GROUPS * 4;
constexpr uint32_t BWD_THREADS = BWD_GROUPS * 128 + 32;
constexpr uint32_t TMEM_DW_BASE = BWD_GROUPS * BWD_WORK_COLS;

template <int WHICH> struct Prog;
template <> struct Prog<1> { static constexpr uint32_t STAGES = 5, BLOB = BW1_BYTES, DW_COLS = DW1_COLS; };
template <> struct Prog<2> { static constexpr uint32_t STAGES = 3, BLOB = BW2_BYTES, DW_COLS = DW2_COLS; };

struct BwdParams {
    uint32_t M;
    const uint8_t* acts;
    const uint8_t* blob;
    // bwd<1>
    const float* d_sigma; const float* d_rgb; const float* rgb; const float* sigma_pre;
    // bwd<2>
    const float* d_ambient; const __half* d_amb01; const float* xyzs; float bound, inv2bound;
    const float* ambient;
    __half* dEx;        // [M_pad, 32]  bwd<1> writes part 1, bwd<2> adds its part in place
    __half* dEw;        // [M_pad, 32]
    float* amb01;       // [M_pad, 2]   the 2-D encoder's input, recomputed with the forward's own expression
    float* x01;         // [M_pad, 3]   the 3-D encoder's input
    float* partials;    // [gridDim.x][DW_COLS][128]
    const int32_t* m_valid;   // optional device scalar: rows >= min(M, *m_valid) are padding
};

__device__ __forceinline__ void mbar_arrive(uint64_t* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(mbar)) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(umma::smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(mbar)), "r"(bytes) : "memory");
}

// activation tile of stage `stage` of tile `tile` -> dst, completion on `bar`
template <int WHICH>
__device__ __forceinline__ void issue_load(const uint8_t* acts, uint32_t tile, uint32_t stage, uint32_t dst, uint64_t* bar) {
    const uint8_t* rec = acts + (size_t)tile * TILE_RECORD_BYTES;
    if constexpr (WHICH == 1) {
        if (stage == 4) {
            expect_tx(bar, 2 * 128 * 32 * 2);
            bulk_load(dst, rec + T_A0, 128 * 32 * 2, bar);
            bulk_load(dst + 128 * 32 * 2, rec + T_EW, 128 * 32 * 2, bar);
        } else {
            const uint32_t off = stage == 0 ? T_HC1 : stage == 1 ? T_CIN : stage == 2 ? T_HS2 : T_HS1;
            const uint32_t bytes = stage == 1 ? 128 * 80 * 2 : 128 * 64 * 2;
            expect_tx(bar, bytes);
            bulk_load(dst, rec + off, bytes, bar);
        }
    } else {
        const uint32_t off = stage == 0 ? T_HA2 : stage == 1 ? T_HA1 : T_A0;
        const uint32_t bytes = stage == 2 ? 128 * 32 * 2 : 128 * 64 * 2;
        expect_tx(bar, bytes);
        bulk_load(dst, rec + off, bytes, bar);
    }
}

// the MMAs of one stage.  act = the stage's activation tile, grad = the group's gradient operand, dw_init: bit s set once stage s's
// accumulators hold data (TMEM is not zeroed: the first update of a region overwrites)
template <int WHICH>
__device__ __forceinline__ void issue_stage(uint32_t stage, uint32_t tmem_work, uint32_t tmem_dw, uint32_t act, uint32_t grad, uint32_t blob,
                                            uint32_t ones, bool acc) {
    using umma::gemm_issue;
    using umma::gemm_issue_mn;
    if constexpr (WHICH == 1) {
        switch (stage) {
            case 0:   // colour L2:  dH = dZc2 [128x16] Wc2 ;  dWc2^T [64 x 16] += Hc1^T dZc2
                gemm_issue(tmem_work, grad, blob + BW1_C2, 16, 16, 0, 16, 64, false);
                gemm_issue_mn(tmem_dw + DW1_C2T, act, 64, 0, grad, 16, 0, 16, 128, acc);
                break;
            case 1:   // colour L1:  d(geo) = dZc1 [128x64] Wc1[:,16:80] ;  dWc1 [64 x 80] += dZc1^T [sh | geo]
                gemm_issue(tmem_work, grad, blob + BW1_C1G, 64, 64, 0, 64, 64, false);
                gemm_issue_mn(tmem_dw + DW1_C1, grad, 64, 0, act, 80, 0, 80, 128, acc);
                break;
            case 2:   // sigma L3:   dH = dH3 [128x80] Ws3p ;  dWs3p [80 x 64] += dH3^T Hs2
                gemm_issue(tmem_work, grad, blob + BW1_S3, 80, 80, 0, 80, 64, false);
                gemm_issue_mn(tmem_dw + DW1_S3, grad, 80, 0, act, 64, 0, 64, 128, acc);
                break;
            case 3:   // sigma L2
                gemm_issue(tmem_work, grad, blob + BW1_S2, 64, 64, 0, 64, 64, false);
                gemm_issue_mn(tmem_dw + DW1_S2, grad, 64, 0, act, 64, 0, 64, 128, acc);
                break;
            default:  // sigma L1:   d[enc_x | enc_w] = dZs1 Ws1[:, 0:64] ;  dWs1 += dZs1^T [enc_x | enc_w] ; colsum(dZs1) for the eye column
                gemm_issue(tmem_work, grad, blob + BW1_S1, 64, 64, 0, 64, 64, false);
                gemm_issue_mn(tmem_dw + DW1_S1, grad, 64, 0, act, 32, 0, 32, 128, acc);
                gemm_issue_mn(tmem_dw + DW1_S1 + 32, grad, 64, 0, act + 128 * 32 * 2, 32, 0, 32, 128, acc);
                gemm_issue_mn(tmem_dw + DW1_CS, grad, 64, 0, ones, 16, 0, 16, 128, acc);
                break;
        }
    } else {
        switch (stage) {
            case 0:   // ambient L3: dH = dZa3 [128x16] Wa3 ;  dWa3^T [64 x 16] += Ha2^T dZa3
                gemm_issue(tmem_work, grad, blob + BW2_A3, 16, 16, 0, 16, 64, false);
                gemm_issue_mn(tmem_dw + DW2_A3T, act, 64, 0, grad, 16, 0, 16, 128, acc);
                break;
            case 1:   // ambient L2
                gemm_issue(tmem_work, grad, blob + BW2_A2, 64, 64, 0, 64, 64, false);
                gemm_issue_mn(tmem_dw + DW2_A2, grad, 64, 0, act, 64, 0, 64, 128, acc);
                break;
            default:  // ambient L1: d(enc_x) = dZa1 Wa1[:, 0:32] ;  dWa1[:, 0:32] += dZa1^T enc_x ; colsum(dZa1) for the audio columns
                gemm_issue(tmem_work, grad, blob + BW2_A1, 64, 64, 0, 64, 32, false);
                gemm_issue_mn(tmem_dw + DW2_A1, grad, 64, 0, act, 32, 0, 32, 128, acc);
                gemm_issue_mn(tmem_dw + DW2_CS, grad, 64, 0, ones, 16, 0, 16, 128, acc);
                break;
        }
    }
}

// accumulator row (64 fp32 columns at tmem_row) x relu'(saved activation row) -> fp16 row of the gradient operand [128 x 64]
static __device__ __noinline__ void mask_epilogue(uint32_t tmem_row, const uint8_t* act, uint8_t* dst, uint32_t row) {
    uint32_t v[2][32];
    umma::tmem_ld32(tmem_row, v[0]);
    umma::tmem_ld32(tmem_row + 32, v[1]);
    umma::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t off = umma::il_offset(row, 32 * c + 8 * q, 64);
            const uint4 a = *reinterpret_cast<const uint4*>(act + off);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 av = *reinterpret_cast<const __half2*>(&aw[j]);
                const float g0 = __low2float(av) > 0.f ? __uint_as_float(v[c][8 * q + 2 * j]) : 0.f;
                const float g1 = __high2float(av) > 0.f ? __uint_as_float(v[c][8 * q + 2 * j + 1]) : 0.f;
                h[j] = pack2(g0, g1);
            }
            *reinterpret_cast<uint4*>(dst + off) = make_uint4(h[0], h[1], h[2], h[3]);
        }
    }
}

template <int WHICH>
__global__ void __launch_bounds__(BWD_THREADS, 1)
head_train_bwd_kernel(BwdParams p) {
    using P = Prog<WHICH>;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_ready[BWD_GROUPS], bar_done[BWD_GROUPS], bar_ld[BWD_GROUPS][2], bar_w;
    __shared__ uint32_t tmem_slot;

    const uint32_t M_eff = p.m_valid ? min(p.M, (uint32_t)max(__ldg(p.m_valid), 0)) : p.M;
    const uint32_t n_tiles = (M_eff + 127) / 128, n_tiles_all = (p.M + 127) / 128;
    // padding rows (zero positions: all in the same grid cells) are marked out of range for the table scatter that follows this kernel;
    // whole padding tiles by a grid-stride sweep here, the tail of the last real tile by its own threads below
    {
        const uint32_t first = n_tiles * 128u, last = n_tiles_all * 128u;
        for (uint32_t s = first + blockIdx.x * blockDim.x + threadIdx.x; s < last; s += gridDim.x * blockDim.x) {
            if constexpr (WHICH == 1) { p.amb01[(size_t)s * 2] = -1.0f; p.amb01[(size_t)s * 2 + 1] = -1.0f; }
            else { p.x01[(size_t)s * 3] = -1.0f; p.x01[(size_t)s * 3 + 1] = -1.0f; p.x01[(size_t)s * 3 + 2] = -1.0f; }
        }
    }
    if (blockIdx.x * BWD_GROUPS >= n_tiles) {
        // no tile for this CTA (the launch was sized for M, the batch holds fewer samples): its partial sums are zeros
        float* out = p.partials + (size_t)blockIdx.x * P::DW_COLS * 128;
        for (uint32_t i = threadIdx.x; i < P::DW_COLS * 128; i += blockDim.x) out[i] = 0.0f;
        return;
    }
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    uint8_t* s_groups = smem;                                     // groups first: the MN-major reads of an 80-wide operand run 768 bytes
    uint8_t* s_blob = smem + BWD_GROUPS * GRP_BYTES;              // past their buffer -- into valid shared memory
    uint8_t* s_ones = s_blob + P::BLOB;

    if (tid == 0) {
        for (int g = 0; g < BWD_GROUPS; ++g) {
            umma::mbar_init(&bar_ready[g], 128);
            umma::mbar_init(&bar_done[g], 1);
            umma::mbar_init(&bar_ld[g][0], 1);
            umma::mbar_init(&bar_ld[g][1], 1);
        }
        umma::mbar_init(&bar_w, 1);
        umma::fence_mbar_init();
        expect_tx(&bar_w, P::BLOB);
        bulk_load(umma::smem_u32(s_blob), p.blob, P::BLOB, &bar_w);
    }
    if (warp == ISSUER_WARP) umma::tmem_alloc(&tmem_slot, 512);
    if (tid < 128) {
        *reinterpret_cast<uint4*>(s_ones + umma::il_offset(tid, 0, 16)) = make_uint4(pack2(1.0f, 1.0f), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(s_ones + umma::il_offset(tid, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    umma::mbar_wait(&bar_w, 0);
    const uint32_t tmem_base = tmem_slot;
    const uint32_t stride = gridDim.x * BWD_GROUPS;

    if (warp == ISSUER_WARP) {
        // ================================================== MMA / copy issuer (one thread) ==================================================
        if ((tid & 31u) == 0) {
            uint32_t n[BWD_GROUPS], tile[BWD_GROUPS], ph[BWD_GROUPS];
            uint32_t active = 0, dw_init = 0;
            const uint32_t a_blob = umma::smem_u32(s_blob), a_ones = umma::smem_u32(s_ones);
#pragma unroll
            for (int g = 0; g < BWD_GROUPS; ++g) {
                n[g] = 0; ph[g] = 0;
                tile[g] = blockIdx.x * BWD_GROUPS + g;
                if (tile[g] < n_tiles) {
                    ++active;
                    issue_load<WHICH>(p.acts, tile[g], 0, umma::smem_u32(s_groups + g * GRP_BYTES), &bar_ld[g][0]);
                }
            }
            while (active) {
#pragma unroll
                for (int g = 0; g < BWD_GROUPS; ++g) {
                    if (tile[g] >= n_tiles) continue;
                    if (!umma::mbar_try_wait(&bar_ready[g], ph[g])) continue;
                    ph[g] ^= 1u;
                    umma::fence_after_sync();
                    const uint32_t k = n[g], stage = k % P::STAGES;
                    const uint32_t a_grp = umma::smem_u32(s_groups + g * GRP_BYTES);
                    issue_stage<WHICH>(stage, tmem_base + g * BWD_WORK_COLS, tmem_base + TMEM_DW_BASE, a_grp + (k & 1u) * ACT_BYTES,
                                       a_grp + 2 * ACT_BYTES, a_blob, a_ones, (dw_init >> stage) & 1u);
                    dw_init |= 1u << stage;
                    umma::commit(&bar_done[g]);
                    // prefetch the next stage's activation tile into the buffer the previous stage has released
                    const uint32_t nstage = stage + 1 == P::STAGES ? 0 : stage + 1;
                    const uint32_t ntile = nstage == 0 ? tile[g] + stride : tile[g];
                    if (ntile < n_tiles) issue_load<WHICH>(p.acts, ntile, nstage, a_grp + ((k + 1) & 1u) * ACT_BYTES, &bar_ld[g][(k + 1) & 1u]);
                    n[g] = k + 1;
                    if (nstage == 0) {
                        tile[g] = ntile;
                        if (ntile >= n_tiles) --active;
                    }
                }
            }
        }
    } else {
        // ================================================== workers ==================================================
        const uint32_t g = tid >> 7, t = tid & 127;
        uint8_t* s_grp = s_groups + g * GRP_BYTES;
        uint8_t* sGRAD = s_grp + 2 * ACT_BYTES;
        const uint32_t tmem_row = tmem_base + g * BWD_WORK_COLS + (((warp & 3u) * 32u) << 16);
        uint32_t k = 0, ph_done = 0;
        auto act = [&](uint32_t kk) -> uint8_t* { return s_grp + (kk & 1u) * ACT_BYTES; };
        // every stage: (operand row written) -> its activation tile has landed -> tell the issuer -> wait for the stage's MMAs
        auto run_stage = [&]() {
            umma::mbar_wait(&bar_ld[g][k & 1u], (k >> 1) & 1u);
            umma::fence_async_smem();
            umma::fence_before_sync();
            mbar_arrive(&bar_ready[g]);
            umma::mbar_wait(&bar_done[g], ph_done);
            ph_done ^= 1u;
            umma::fence_after_sync();
            ++k;
        };
        for (uint32_t tile = blockIdx.x * BWD_GROUPS + g; tile < n_tiles; tile += stride) {
            const uint32_t s = tile * 128 + t;
            const bool valid = s < M_eff;
            const uint32_t k0 = k;
            if constexpr (WHICH == 1) {
                // ---- dL/d(colour pre-activation) = dL/d(rgb) * sigmoid'  -> dZc2 [128 x 16]
                {
                    float dz[3] = {0.f, 0.f, 0.f};
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            const float r = __ldg(p.rgb + (size_t)s * 3 + j);
                            dz[j] = __ldg(p.d_rgb + (size_t)s * 3 + j) * r * (1.0f - r);
                        }
                    }
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 0, 16)) = make_uint4(pack2(dz[0], dz[1]), pack2(dz[2], 0.f), 0u, 0u);
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
                }
                run_stage();                                   // colour L2
                mask_epilogue(tmem_row, act(k0), sGRAD, t);    // relu'(Hc1)
                run_stage();                                   // colour L1
                {   // dH3 [128 x 80] = [d(geo_feat) | dL/d(log-density) | 0]
                    uint32_t v[2][32];
                    umma::tmem_ld32(tmem_row, v[0]);
                    umma::tmem_ld32(tmem_row + 32, v[1]);
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 32 * c + 8 * q, 80)) =
                                make_uint4(pack2(__uint_as_float(v[c][8 * q]), __uint_as_float(v[c][8 * q + 1])),
                                           pack2(__uint_as_float(v[c][8 * q + 2]), __uint_as_float(v[c][8 * q + 3])),
                                           pack2(__uint_as_float(v[c][8 * q + 4]), __uint_as_float(v[c][8 * q + 5])),
                                           pack2(__uint_as_float(v[c][8 * q + 6]), __uint_as_float(v[c][8 * q + 7])));
                    }
                    float dsig = 0.f;   // trunc_exp backward (activation.py:12-15): g * exp(clamp(x, -15, 15))
                    if (valid) dsig = __ldg(p.d_sigma + s) * expf(fminf(fmaxf(__ldg(p.sigma_pre + s), -15.0f), 15.0f));
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 64, 80)) = make_uint4(pack2(dsig, 0.f), 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 72, 80)) = make_uint4(0u, 0u, 0u, 0u);
                }
                run_stage();                                       // sigma L3
                mask_epilogue(tmem_row, act(k0 + 2), sGRAD, t);    // relu'(Hs2)
                run_stage();                                       // sigma L2
                mask_epilogue(tmem_row, act(k0 + 3), sGRAD, t);    // relu'(Hs1)
                run_stage();                                       // sigma L1
                {
                    uint32_t v[2][32];
                    umma::tmem_ld32(tmem_row, v[0]);
                    umma::tmem_ld32(tmem_row + 32, v[1]);
                    umma::tmem_ld_wait();
                    uint4* ex = reinterpret_cast<uint4*>(p.dEx + (size_t)s * 32);
                    uint4* ew = reinterpret_cast<uint4*>(p.dEw + (size_t)s * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ex[q] = make_uint4(pack2(__uint_as_float(v[0][8 * q]), __uint_as_float(v[0][8 * q + 1])),
                                           pack2(__uint_as_float(v[0][8 * q + 2]), __uint_as_float(v[0][8 * q + 3])),
                                           pack2(__uint_as_float(v[0][8 * q + 4]), __uint_as_float(v[0][8 * q + 5])),
                                           pack2(__uint_as_float(v[0][8 * q + 6]), __uint_as_float(v[0][8 * q + 7])));
                        ew[q] = make_uint4(pack2(__uint_as_float(v[1][8 * q]), __uint_as_float(v[1][8 * q + 1])),
                                           pack2(__uint_as_float(v[1][8 * q + 2]), __uint_as_float(v[1][8 * q + 3])),
                                           pack2(__uint_as_float(v[1][8 * q + 4]), __uint_as_float(v[1][8 * q + 5])),
                                           pack2(__uint_as_float(v[1][8 * q + 6]), __uint_as_float(v[1][8 * q + 7])));
                    }
                    float a0 = 0.f, a1 = 0.f;
                    if (valid) { a0 = __ldg(p.ambient + (size_t)s * 2); a1 = __ldg(p.ambient + (size_t)s * 2 + 1); }
                    *reinterpret_cast<float2*>(p.amb01 + (size_t)s * 2) =
                        valid ? make_float2(__fmul_rn(__fadd_rn(a0, 1.0f), 0.5f), __fmul_rn(__fadd_rn(a1, 1.0f), 0.5f)) : make_float2(-1.0f, -1.0f);
                }
            } else {
                // ---- dL/d(ambient pre-activation) = (compositor path + 0.5 * 2-D grid path) * tanh'  -> dZa3 [128 x 16]
                {
                    float dz[2] = {0.f, 0.f};
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const float a = __ldg(p.ambient + (size_t)s * 2 + j);
                            const float da = __ldg(p.d_ambient + (size_t)s * 2 + j) + 0.5f * __half2float(p.d_amb01[(size_t)s * 2 + j]);
                            dz[j] = da * (1.0f - a * a);
                        }
                    }
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 0, 16)) = make_uint4(pack2(dz[0], dz[1]), 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(sGRAD + umma::il_offset(t, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
                }
                run_stage();                                   // ambient L3
                mask_epilogue(tmem_row, act(k0), sGRAD, t);    // relu'(Ha2)
                run_stage();                                   // ambient L2
                mask_epilogue(tmem_row, act(k0 + 1), sGRAD, t);  // relu'(Ha1)
                run_stage();                                   // ambient L1
                {
                    uint32_t v[32];
                    umma::tmem_ld32(tmem_row, v);
                    umma::tmem_ld_wait();
                    uint4* ex = reinterpret_cast<uint4*>(p.dEx + (size_t)s * 32);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 prev = ex[q];   // part 1, written by bwd<1>
                        const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
                        uint32_t o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&pw[j]));
                            o[j] = pack2(f.x + __uint_as_float(v[8 * q + 2 * j]), f.y + __uint_as_float(v[8 * q + 2 * j + 1]));
                        }
                        ex[q] = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                    float px = 0.f, py = 0.f, pz = 0.f;
                    if (valid) { px = __ldg(p.xyzs + (size_t)s * 3); py = __ldg(p.xyzs + (size_t)s * 3 + 1); pz = __ldg(p.xyzs + (size_t)s * 3 + 2); }
                    float* xo = p.x01 + (size_t)s * 3;
                    xo[0] = valid ? __fmul_rn(__fadd_rn(px, p.bound), p.inv2bound) : -1.0f;
                    xo[1] = valid ? __fmul_rn(__fadd_rn(py, p.bound), p.inv2bound) : -1.0f;
                    xo[2] = valid ? __fmul_rn(__fadd_rn(pz, p.bound), p.inv2bound) : -1.0f;
                }
            }
        }
    }

    // ---- write the CTA's weight-gradient accumulators out: partials[cta][column][lane]
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (warp != ISSUER_WARP) {
        const uint32_t g = tid >> 7, t = tid & 127;
        const uint32_t row = tmem_base + TMEM_DW_BASE + (((warp & 3u) * 32u) << 16);
        float* out = p.partials + (size_t)blockIdx.x * P::DW_COLS * 128;
        for (uint32_t c0 = g * 16; c0 < P::DW_COLS; c0 += BWD_GROUPS * 16) {
            uint32_t v[16];
            umma::tmem_ld16(row + c0, v);
            umma::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) out[(size_t)(c0 + j) * 128 + t] = __uint_as_float(v[j]);
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == ISSUER_WARP) umma::tmem_dealloc(tmem_base, 512);
}

// ---- sum the per-CTA partials into nn.Linear-shaped fp32 gradients (head_train.cuh G_*) ----------------------------------------------
struct ReduceParams { const float* p1; const float* p2; uint32_t n1, n2; float* out; };

__device__ __forceinline__ float sum_partials(const float* p, uint32_t n_cta, uint32_t cols, uint32_t col, uint32_t lane) {
    float acc = 0.f;
    for (uint32_t c = 0; c < n_cta; ++c) acc += __ldg(p + ((size_t)c * cols + col) * 128 + lane);
    return acc;
}

__global__ void __launch_bounds__(256)
dw_reduce_kernel(ReduceParams r) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= G_FLOATS) return;
    float v;
    if (e < G_WA2) { const uint32_t o = e / 32, i = e % 32; v = sum_partials(r.p2, r.n2, DW2_COLS, DW2_A1 + i, o); }
    else if (e < G_WA3) { const uint32_t x = e - G_WA2, o = x / 64, i = x % 64; v = sum_partials(r.p2, r.n2, DW2_COLS, DW2_A2 + i, o); }
    else if (e < G_WS1) { const uint32_t x = e - G_WA3, o = x / 64, i = x % 64; v = sum_partials(r.p2, r.n2, DW2_COLS, DW2_A3T + o, i); }
    else if (e < G_WS2) { const uint32_t x = e - G_WS1, o = x / 64, i = x % 64; v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_S1 + i, o); }
    else if (e < G_WS3) { const uint32_t x = e - G_WS2, o = x / 64, i = x % 64; v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_S2 + i, o); }
    else if (e < G_WC1) { const uint32_t x = e - G_WS3, o = x / 64, i = x % 64; v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_S3 + i, o == 0 ? 64u : o - 1); }
    else if (e < G_WC2) { const uint32_t x = e - G_WC1, o = x / 80, i = x % 80; v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_C1 + i, o); }
    else if (e < G_CS_A1) { const uint32_t x = e - G_WC2, o = x / 64, i = x % 64; v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_C2T + o, i); }
    else if (e < G_CS_S1) v = sum_partials(r.p2, r.n2, DW2_COLS, DW2_CS, e - G_CS_A1);
    else if (e < G_CS_C1) v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_CS, e - G_CS_S1);
    else   // SH band 0 is the constant 0.28209479 (stored as fp16 in the tile): that column of dWc1 is the column sum times it
        v = sum_partials(r.p1, r.n1, DW1_COLS, DW1_C1, e - G_CS_C1) / __half2float(__float2half_rn(0.28209479177387814f));
    r.out[e] = v;
}

constexpr uint32_t smem_bytes(uint32_t blob) { return BWD_GROUPS * GRP_BYTES + blob + ONES_BYTES; }

struct Workspace {
    __half* dEx; __half* dEw; __half* d_amb01; float* amb01; float* x01; float* part1; float* part2;
};
size_t carve_ws(Workspace& w, uint8_t* base, uint32_t M) {
    const size_t mp = ((size_t)M + 127) / 128 * 128;
    size_t off = 0;
    auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off += (bytes + 255) & ~(size_t)255; return p; };
    w.dEx = (__half*)take(mp * 64);
    w.dEw = (__half*)take(mp * 64);
    w.d_amb01 = (__half*)take(mp * 4);
    w.amb01 = (float*)take(mp * 8);
    w.x01 = (float*)take(mp * 12);
    w.part1 = (float*)take((size_t)RN_NUM_SMS * DW1_COLS * 128 * 4);
    w.part2 = (float*)take((size_t)RN_NUM_SMS * DW2_COLS * 128 * 4);
    return off;
}

}  // namespace

int launch_head_train_fwd(const rn_head_train_desc* d, cudaStream_t st);

}  // namespace train
}  // namespace rn

using namespace rn;
using namespace rn::train;

extern "C" uint64_t rn_head_train_acts_bytes(uint32_t M) { return (uint64_t)((M + 127) / 128) * TILE_RECORD_BYTES; }
extern "C" uint64_t rn_head_train_workspace_bytes(uint32_t M) { Workspace w; return (uint64_t)carve_ws(w, nullptr, M); }
extern "C" uint32_t rn_head_train_bwd_blob_bytes(void) { return BWD_BLOB_BYTES; }
extern "C" uint32_t rn_head_train_dw_floats(void) { return G_FLOATS; }

static int check_common(const rn_head_train_desc* d, bool training) {
    RN_REQUIRE(d, "null descriptor");
    RN_REQUIRE(d->xyzs && d->fwd_blob && d->consts && d->sigma, "null pointer");
    // forward only: acts / dy_dx2 / sigma_pre / ambient may be NULL (nothing saved), rgb == NULL makes it a density query
    RN_REQUIRE(!training || (d->dirs && d->acts && d->dy_dx2 && d->rgb && d->ambient && d->sigma_pre), "training needs every forward buffer");
    RN_REQUIRE(d->rgb == nullptr || d->dirs, "colours need view directions");
    RN_REQUIRE(d->grid3d.table_f16 && d->grid3d.offsets && d->grid3d.packed_offsets && d->grid2d.table_f16 && d->grid2d.offsets &&
                   d->grid2d.packed_offsets, "grid tables incomplete");
    RN_REQUIRE(d->bound > 0.f, "bound must be positive");
    RN_REQUIRE(((uintptr_t)d->fwd_blob & 15) == 0 && ((uintptr_t)d->acts & 127) == 0, "blob / activation buffers must be 16 / 128-byte aligned");
    return RN_OK;
}

extern "C" int rn_head_train_forward(const rn_head_train_desc* d, void* stream) {
    if (int rc = check_common(d, false)) return rc;
    if (d->M == 0) return RN_OK;
    return launch_head_train_fwd(d, (cudaStream_t)stream);
}

extern "C" int rn_head_train_backward(const rn_head_train_desc* d, void* stream) {
    if (int rc = check_common(d, true)) return rc;
    RN_REQUIRE(d->bwd_blob && d->d_sigma && d->d_rgb && d->d_ambient && d->d_table3 && d->d_table2 && d->d_weights, "null backward pointer");
    RN_REQUIRE(d->workspace && d->workspace_bytes >= rn_head_train_workspace_bytes(d->M), "workspace too small");
    RN_REQUIRE(((uintptr_t)d->bwd_blob & 15) == 0 && ((uintptr_t)d->workspace & 255) == 0, "bwd_blob / workspace alignment");
    cudaStream_t st = (cudaStream_t)stream;
    Workspace w;
    carve_ws(w, (uint8_t*)d->workspace, d->M);
    const uint32_t n_tiles = (d->M + 127) / 128;
    uint32_t grid = (n_tiles + BWD_GROUPS - 1) / BWD_GROUPS;
    if (grid > RN_NUM_SMS) grid = RN_NUM_SMS;
    if (d->M == 0) {
        cudaMemsetAsync(d->d_weights, 0, G_FLOATS * 4, st);
        return RN_OK;
    }
    BwdParams p{};
    p.M = d->M; p.acts = (const uint8_t*)d->acts;
    p.d_sigma = d->d_sigma; p.d_rgb = d->d_rgb; p.rgb = d->rgb; p.sigma_pre = d->sigma_pre;
    p.d_ambient = d->d_ambient; p.d_amb01 = w.d_amb01; p.xyzs = d->xyzs; p.bound = d->bound; p.inv2bound = 1.0f / (2.0f * d->bound);
    p.ambient = d->ambient; p.dEx = w.dEx; p.dEw = w.dEw; p.amb01 = w.amb01; p.x01 = w.x01;
    p.m_valid = d->m_valid;

    // ---- colour + sigma nets
    {
        constexpr uint32_t smem = smem_bytes(BW1_BYTES);
        cudaError_t e = cudaFuncSetAttribute(head_train_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("head_train_bwd<1>: cannot reserve %u bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
        p.blob = (const uint8_t*)d->bwd_blob; p.partials = w.part1;
        head_train_bwd_kernel<1><<<grid, BWD_THREADS, smem, st>>>(p);
        if (int rc = finish_launch("rn_head_train_backward (colour + sigma)")) return rc;
    }
    // ---- 2-D ambient grid: table scatter + dL/d(coordinate) through the saved d(enc)/d(coordinate)
    if (int rc = rn_grid_encode_backward(w.dEw, w.amb01, nullptr, d->grid2d.offsets, d->d_table2, d->M, 2, 2, 16, d->grid2d.S, d->grid2d.H, d->dy_dx2,
                                         w.d_amb01, 1, 0, 0, RN_F16, RN_LAYOUT_BLC, RN_F32, stream))
        return rc;
    // ---- ambient net
    {
        constexpr uint32_t smem = smem_bytes(BW2_BYTES);
        cudaError_t e = cudaFuncSetAttribute(head_train_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("head_train_bwd<2>: cannot reserve %u bytes of shared memory: %s", smem, cudaGetErrorString(e)); return (int)e; }
        p.blob = (const uint8_t*)d->bwd_blob + BW1_BYTES; p.partials = w.part2;
        head_train_bwd_kernel<2><<<grid, BWD_THREADS, smem, st>>>(p);
        if (int rc = finish_launch("rn_head_train_backward (ambient)")) return rc;
    }
    // ---- 3-D spatial grid: table scatter
    if (int rc = rn_grid_encode_backward(w.dEx, w.x01, nullptr, d->grid3d.offsets, d->d_table3, d->M, 3, 2, 16, d->grid3d.S, d->grid3d.H, nullptr, nullptr,
                                         1, 0, 0, RN_F16, RN_LAYOUT_BLC, RN_F32, stream))
        return rc;
    // ---- weight gradients
    ReduceParams r{w.part1, w.part2, grid, grid, d->d_weights};
    dw_reduce_kernel<<<div_up((uint32_t)G_FLOATS, 256u), 256, 0, st>>>(r);
    return finish_launch("rn_head_train_backward (dW reduce)");
}
