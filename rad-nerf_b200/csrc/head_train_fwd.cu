// head_train_fwd.cu -- forward of the fused training step: NeRFNetwork.forward (nerf/network.py:222-283) on a batch of
// march_rays_train samples, as ONE persistent tcgen05 kernel that also saves what the backward needs.
//
// Same pipeline, same rounding points as head_eval.cu (the inference kernel): 128-sample tiles, thread t of a tile group owns
// sample row t, every layer input is built as an fp16 row of an interleaved UMMA operand in shared memory, accumulators live in
// TMEM.  Training adds, per produced row, ONE extra copy of the row to global memory at the same interleaved offset inside the
// tile's record (head_train.cuh, 928 B/sample): the backward reloads whole tiles with bulk copies and uses them both as the
// relu masks and as the MN-major operands of the weight-gradient MMAs.  Also saved: the log-density before trunc_exp, and the
// 2-D grid's d(enc)/d(coordinate) in the layout rn_grid_encode_backward expects (gridencoder.cu:200-243).
//
// Also here: the two packing kernels that rebuild the fp16 operand copies (tables, weight blobs) after every optimiser step in
// one launch each (the inference path did this with ~60 torch ops once per checkpoint).
#include "frame.cuh"
#include "umma.cuh"
#include "sh.cuh"
#include "mlp_tile.cuh"
#include "head_train.cuh"

namespace rn {
namespace train {
namespace {

// forward blob sub-matrix byte offsets (identical to head_eval.cu)
constexpr uint32_t B_WA1 = 0;
constexpr uint32_t B_WA2 = B_WA1 + 64 * 32 * 2;
constexpr uint32_t B_WA3 = B_WA2 + 64 * 64 * 2;
constexpr uint32_t B_WS1A = B_WA3 + 16 * 64 * 2;
constexpr uint32_t B_WS1B = B_WS1A + 64 * 32 * 2;
constexpr uint32_t B_WS2 = B_WS1B + 64 * 32 * 2;
constexpr uint32_t B_WS3 = B_WS2 + 64 * 64 * 2;
constexpr uint32_t B_WC1 = B_WS3 + 80 * 64 * 2;
constexpr uint32_t B_WC2 = B_WC1 + 64 * 80 * 2;
static_assert(B_WC2 + 16 * 64 * 2 == HEAD_BLOB_BYTES, "blob layout");

constexpr uint32_t G_A0 = 0;
constexpr uint32_t G_H1 = G_A0 + 128 * 32 * 2;
constexpr uint32_t G_CIN = 0;
constexpr uint32_t G_H0 = G_H1 + 128 * 64 * 2;
constexpr uint32_t G_EW = G_H0;
constexpr uint32_t GROUP_BYTES = G_H0 + 128 * 64 * 2;
constexpr int FWD_GROUPS = 4;
constexpr uint32_t FWD_SMEM = HEAD_BLOB_BYTES + FWD_GROUPS * GROUP_BYTES;
constexpr uint32_t TMEM_COLS_PER_GROUP = 128;

struct FwdParams {
    const float* xyzs; const float* dirs; uint32_t M;
    const __half* table3; const int32_t* offs3; const int32_t* poffs3; float S3; uint32_t H3;
    const __half* table2; const int32_t* offs2; const int32_t* poffs2; float S2; uint32_t H2;
    const uint8_t* blob; const float* consts;
    float* sigma; float* rgb; float* ambient; float* sigma_pre;
    uint8_t* acts; __half* dy_dx2;
    float bound, inv2bound;
    const int32_t* m_valid;
};

// TMEM row -> (ReLU) -> fp16 -> the next layer's A operand row in shared memory AND the same row of the tile record in global memory
template <int NCH32>
static __device__ __noinline__ void epilogue_save(uint32_t tmem_row, uint32_t col0, bool relu, uint8_t* dst, uint8_t* gdst, uint32_t row,
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
        for (int q = 0; q < 4; ++q) {
            const uint32_t off = umma::il_offset(row, dcol0 + 32 * c + 8 * q, Kdst);
            const uint4 w = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
            *reinterpret_cast<uint4*>(dst + off) = w;
            if (gdst) *reinterpret_cast<uint4*>(gdst + off) = w;
        }
    }
}

// mma_stage (mlp_tile.cuh) that also SAVES the stage's freshly published A operand: the tile sits in shared memory in exactly the
// interleaved layout of its slot in the tile record, so thread 0 hands it to the TMA engine as one bulk store (8-20 KB) right after
// issuing the MMAs -- instead of every thread mirroring each of its 16-byte operand stores to global memory (58 per sample: the
// forward ran at half the speed of the inference kernel because of them).  The source buffer is only rewritten by an epilogue that
// follows a LATER stage's MMA, and thread 0 waits for the reads of all earlier bulk stores before it issues that later stage.
static __device__ __noinline__ void mma_stage_save(uint32_t tmem_acc, uint32_t a0, uint32_t Ka0, uint32_t ka0, uint32_t b0, uint32_t K0,
                                                   uint32_t a1, uint32_t b1, uint32_t K1, uint32_t a2, uint32_t b2, uint32_t N, uint64_t* mbar,
                                                   uint32_t& phase, uint32_t bar_id, uint32_t t, uint8_t* save_dst, uint32_t save_src, uint32_t save_bytes) {
    umma::fence_async_smem();
    umma::fence_before_sync();
    umma::group_sync(bar_id, 128);
    if (t == 0) {
        umma::fence_after_sync();
        if (save_dst) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // earlier saves have left shared memory
        umma::gemm_issue(tmem_acc, a0, b0, Ka0, K0, ka0, K0, N, false);
        if (K1) umma::gemm_issue(tmem_acc, a1, b1, K1, K1, 0, K1, N, true);
        if (a2) umma::gemm_issue(tmem_acc, a2, b2, 16, 16, 0, 16, N, true);
        umma::commit(mbar);
        if (save_dst) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(save_dst), "r"(save_src), "r"(save_bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    umma::mbar_wait(mbar, phase);
    phase ^= 1u;
    umma::fence_after_sync();
}

// fast_encode<D> (mlp_tile.cuh) + the saved copy; for D == 2 additionally d(features)/d(x) per level, dy_dx[l][d][c] fp16
// (the reference's kernel_grid with calc_grad_inputs, gridencoder.cu:200-243; fp32 accumulation, one rounding)
template <int D>
static __device__ __noinline__ void encode_save(const float (&x)[D], const uint2* __restrict__ table64, const FastLevel* __restrict__ lv,
                                                uint8_t* dst, uint8_t* gdst, uint32_t row, uint32_t Kdst, __half* dy_dx /* [16*D*2] or null */) {
    bool oob = false;
#pragma unroll
    for (int d = 0; d < D; ++d) if (x[d] < 0 || x[d] > 1) oob = true;
#pragma unroll 1
    for (int l0 = 0; l0 < 16; l0 += 4) {
        uint32_t g[4][1 << D];
        float fr[4][D];
        float sc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const FastLevel L = lv[l0 + j];
            sc[j] = L.scale;
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
        uint32_t packed[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
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
            if constexpr (D == 2) {
                if (dy_dx) {
                    // corners: g[0] = (0,0), g[1] = (1,0), g[2] = (0,1), g[3] = (1,1);  d/dx = scale * sum_y w_y (g(1,y) - g(0,y))
                    float2 c00, c10, c01, c11;
                    {
                        const __half2 h0 = *reinterpret_cast<const __half2*>(&g[j][0]), h1 = *reinterpret_cast<const __half2*>(&g[j][1]);
                        const __half2 h2 = *reinterpret_cast<const __half2*>(&g[j][2]), h3 = *reinterpret_cast<const __half2*>(&g[j][3]);
                        c00 = __half22float2(h0); c10 = __half22float2(h1); c01 = __half22float2(h2); c11 = __half22float2(h3);
                    }
                    const float y0 = 1.0f - fr[j][1], y1 = fr[j][1], x0 = 1.0f - fr[j][0], x1 = fr[j][0];
                    const float dxa = sc[j] * (y0 * (c10.x - c00.x) + y1 * (c11.x - c01.x));
                    const float dxb = sc[j] * (y0 * (c10.y - c00.y) + y1 * (c11.y - c01.y));
                    const float dya = sc[j] * (x0 * (c01.x - c00.x) + x1 * (c11.x - c10.x));
                    const float dyb = sc[j] * (x0 * (c01.y - c00.y) + x1 * (c11.y - c10.y));
                    // [l][d][c]: 4 halfs per level
                    *reinterpret_cast<uint2*>(dy_dx + (l0 + j) * 4) = make_uint2(oob ? 0u : pack2(dxa, dxb), oob ? 0u : pack2(dya, dyb));
                }
            }
        }
        const uint32_t off = umma::il_offset(row, 2 * l0, Kdst);
        const uint4 w = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        *reinterpret_cast<uint4*>(dst + off) = w;
        if (gdst) *reinterpret_cast<uint4*>(gdst + off) = w;
    }
}

__global__ void __launch_bounds__(FWD_GROUPS * 128, 1)
head_train_fwd_kernel(FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ FastLevel lv3[16], lv2[16];
    __shared__ __align__(128) uint8_t s_ones[128 * 16 * 2];
    __shared__ __align__(128) uint8_t s_biasop[3][64 * 16 * 2];
    __shared__ __align__(8) uint64_t mbar_group[FWD_GROUPS];
    __shared__ __align__(8) uint64_t mbar_w;
    __shared__ uint32_t tmem_slot;

    // rows past *m_valid are padding of the sample buffers: whole padding tiles are skipped (their outputs stay as the caller left them)
    const uint32_t M_eff = p.m_valid ? min(p.M, (uint32_t)max(__ldg(p.m_valid), 0)) : p.M;
    const uint32_t n_tiles = (M_eff + EVAL_TILE - 1) / EVAL_TILE;
    if (blockIdx.x * FWD_GROUPS >= n_tiles) return;

    const uint32_t tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = tid >> 5;
    uint8_t* s_blob = smem;
    uint8_t* s_grp = smem + HEAD_BLOB_BYTES + g * GROUP_BYTES;

    if (tid == 0) {
        for (int i = 0; i < FWD_GROUPS; ++i) umma::mbar_init(&mbar_group[i], 1);
        umma::mbar_init(&mbar_w, 1);
        umma::fence_mbar_init();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&mbar_w)), "r"(HEAD_BLOB_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(s_blob)),
                     "l"(p.blob), "r"(HEAD_BLOB_BYTES), "r"(umma::smem_u32(&mbar_w))
                     : "memory");
    }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 512);
    if (tid >= 64 && tid < 80) {
        grid::LevelMeta m;
        grid::make_level_meta(m, tid - 64, p.offs3, p.S3, p.H3, 3, 1, false);
        if (!make_fast_level(lv3[tid - 64], m, (uint32_t)__ldg(p.poffs3 + (tid - 64)))) __trap();
    }
    if (tid >= 96 && tid < 112) {
        grid::LevelMeta m;
        grid::make_level_meta(m, tid - 96, p.offs2, p.S2, p.H2, 2, 1, false);
        if (!make_fast_level(lv2[tid - 96], m, (uint32_t)__ldg(p.poffs2 + (tid - 96)))) __trap();
    }
    if (tid < 128) {
        *reinterpret_cast<uint4*>(s_ones + umma::il_offset(tid, 0, 16)) = make_uint4(pack2(1.0f, 1.0f), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(s_ones + umma::il_offset(tid, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
    } else if (tid < 128 + 192) {
        const uint32_t layer = (tid - 128) >> 6, n = (tid - 128) & 63;
        const float b = __ldg(p.consts + (tid - 128));
        const float hi = __half2float(__float2half_rn(b));
        *reinterpret_cast<uint4*>(s_biasop[layer] + umma::il_offset(n, 0, 16)) = make_uint4(pack2(hi, b - hi), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(s_biasop[layer] + umma::il_offset(n, 8, 16)) = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    bool weights_ready = false;     // awaited at the first MMA (see head_eval.cu)

    const uint32_t tmem_acc = tmem_slot + g * TMEM_COLS_PER_GROUP;
    const uint32_t tmem_row = tmem_acc + (((warp & 3u) * 32u) << 16);
    uint64_t* mbar = &mbar_group[g];
    uint32_t phase = 0;
    const uint32_t bar_id = 1 + g;
    uint8_t* sA0 = s_grp + G_A0;
    uint8_t* sEW = s_grp + G_EW;
    uint8_t* sH0 = s_grp + G_H0;
    uint8_t* sH1 = s_grp + G_H1;
    uint8_t* sCIN = s_grp + G_CIN;
    const uint32_t aA0 = umma::smem_u32(sA0), aEW = umma::smem_u32(sEW), aH0 = umma::smem_u32(sH0), aH1 = umma::smem_u32(sH1),
                   aCIN = umma::smem_u32(sCIN);
    const uint32_t aW = umma::smem_u32(s_blob);
    const uint32_t aOnes = umma::smem_u32(s_ones), aB0 = umma::smem_u32(s_biasop[0]), aB1 = umma::smem_u32(s_biasop[1]), aB2 = umma::smem_u32(s_biasop[2]);
    const uint2* table3 = reinterpret_cast<const uint2*>(p.table3);
    const uint2* table2 = reinterpret_cast<const uint2*>(p.table2);

    for (uint32_t tile = blockIdx.x * FWD_GROUPS + g; tile < n_tiles; tile += gridDim.x * FWD_GROUPS) {
        const uint32_t s = tile * EVAL_TILE + t;
        const bool valid = s < M_eff;
        // p.acts == nullptr: nothing is saved (the density query of the occupancy update); REC(x) = where layer input x is saved
        uint8_t* rec = p.acts ? p.acts + (size_t)tile * TILE_RECORD_BYTES : nullptr;
#define REC(off) (rec ? rec + (off) : nullptr)
        float px = 0.f, py = 0.f, pz = 0.f;
        if (valid) { px = __ldg(p.xyzs + (size_t)s * 3); py = __ldg(p.xyzs + (size_t)s * 3 + 1); pz = __ldg(p.xyzs + (size_t)s * 3 + 2); }

        // ---- 3-D encode -> A0
        {
            float x[3] = {__fmul_rn(__fadd_rn(px, p.bound), p.inv2bound), __fmul_rn(__fadd_rn(py, p.bound), p.inv2bound),
                          __fmul_rn(__fadd_rn(pz, p.bound), p.inv2bound)};
            encode_save<3>(x, table3, lv3, sA0, nullptr, t, 32, nullptr);
        }
        // ---- ambient L1 -> H0 (A0 saved)
        if (!weights_ready) { umma::mbar_wait(&mbar_w, 0); weights_ready = true; }
        mma_stage_save(tmem_acc, aA0, 32, 0, aW + B_WA1, 32, 0, 0, 0, aOnes, aB0, 64, mbar, phase, bar_id, t, REC(T_A0), aA0, 128 * 32 * 2);
        epilogue_save<2>(tmem_row, 0, true, sH0, nullptr, t, 64, 0);
        // ---- ambient L2 -> H1 (H0 saved as HA1)
        mma_stage_save(tmem_acc, aH0, 64, 0, aW + B_WA2, 64, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t, REC(T_HA1), aH0, 128 * 64 * 2);
        epilogue_save<2>(tmem_row, 0, true, sH1, nullptr, t, 64, 0);
        // ---- ambient L3 -> tanh -> 2-D encode -> EW (H1 saved as HA2)
        mma_stage_save(tmem_acc, aH1, 64, 0, aW + B_WA3, 64, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t, REC(T_HA2), aH1, 128 * 64 * 2);
        {
            uint32_t v[16];
            umma::tmem_ld16(tmem_row, v);
            umma::tmem_ld_wait();
            const float a0 = tanhf(__half2float(__float2half_rn(__uint_as_float(v[0]))));
            const float a1 = tanhf(__half2float(__float2half_rn(__uint_as_float(v[1]))));
            if (valid && p.ambient) *reinterpret_cast<float2*>(p.ambient + (size_t)s * 2) = make_float2(a0, a1);
            float x[2] = {__fmul_rn(__fadd_rn(a0, 1.0f), 0.5f), __fmul_rn(__fadd_rn(a1, 1.0f), 0.5f)};
            // rows past M still own 128 bytes of dy_dx2 (the buffer is sized for whole tiles)
            encode_save<2>(x, table2, lv2, sEW, nullptr, t, 32, p.dy_dx2 ? p.dy_dx2 + (size_t)s * 64 : nullptr);
        }
        // ---- sigma L1 -> H1 (EW saved)
        mma_stage_save(tmem_acc, aA0, 32, 0, aW + B_WS1A, 32, aEW, aW + B_WS1B, 32, aOnes, aB1, 64, mbar, phase, bar_id, t, REC(T_EW), aEW, 128 * 32 * 2);
        epilogue_save<2>(tmem_row, 0, true, sH1, nullptr, t, 64, 0);
        // ---- sigma L2 -> H0 (H1 saved as HS1)
        mma_stage_save(tmem_acc, aH1, 64, 0, aW + B_WS2, 64, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t, REC(T_HS1), aH1, 128 * 64 * 2);
        epilogue_save<2>(tmem_row, 0, true, sH0, nullptr, t, 64, 0);
        // ---- sigma L3 -> geo_feat (columns 0..63), log-density (column 64)  (H0 saved as HS2)
        mma_stage_save(tmem_acc, aH0, 64, 0, aW + B_WS3, 64, 0, 0, 0, 0, 0, 80, mbar, phase, bar_id, t, REC(T_HS2), aH0, 128 * 64 * 2);
        float sigma;
        {
            epilogue_save<2>(tmem_row, 0, false, sCIN, nullptr, t, 80, 16);
            uint32_t v[16];
            umma::tmem_ld16(tmem_row + 64, v);
            umma::tmem_ld_wait();
            const float hpre = __half2float(__float2half_rn(__uint_as_float(v[0])));
            sigma = expf(hpre);
            if (valid && p.sigma_pre) p.sigma_pre[s] = hpre;
            if (!p.rgb) {   // density query (NeRFNetwork.density, nerf/network.py:286-325): sigma only, CTA-uniform branch
                if (valid) p.sigma[s] = sigma;
                umma::fence_before_sync();
                umma::group_sync(bar_id, 128);
                continue;
            }
            float dx = 0.f, dy = 0.f, dz = 0.f;
            if (valid) { dx = __ldg(p.dirs + (size_t)s * 3); dy = __ldg(p.dirs + (size_t)s * 3 + 1); dz = __ldg(p.dirs + (size_t)s * 3 + 2); }
            float Y[16];
            sh_eval<4, false>(dx, dy, dz, Y, nullptr, nullptr, nullptr);
            const uint4 w0 = make_uint4(pack2(Y[0], Y[1]), pack2(Y[2], Y[3]), pack2(Y[4], Y[5]), pack2(Y[6], Y[7]));
            const uint4 w1 = make_uint4(pack2(Y[8], Y[9]), pack2(Y[10], Y[11]), pack2(Y[12], Y[13]), pack2(Y[14], Y[15]));
            *reinterpret_cast<uint4*>(sCIN + umma::il_offset(t, 0, 80)) = w0;
            *reinterpret_cast<uint4*>(sCIN + umma::il_offset(t, 8, 80)) = w1;
        }
        // ---- colour L1 -> H0 (CIN saved)
        mma_stage_save(tmem_acc, aCIN, 80, 0, aW + B_WC1, 80, 0, 0, 0, aOnes, aB2, 64, mbar, phase, bar_id, t, REC(T_CIN), aCIN, 128 * 80 * 2);
        epilogue_save<2>(tmem_row, 0, true, sH0, nullptr, t, 64, 0);
        // ---- colour L2 -> sigmoid (H0 saved as HC1)
        mma_stage_save(tmem_acc, aH0, 64, 0, aW + B_WC2, 64, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t, REC(T_HC1), aH0, 128 * 64 * 2);
        {
            uint32_t v[16];
            umma::tmem_ld16(tmem_row, v);
            umma::tmem_ld_wait();
            float c[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const float h = __half2float(__float2half_rn(__uint_as_float(v[j])));
                c[j] = __half2float(__float2half_rn(1.0f / (1.0f + expf(-h))));
            }
            if (valid) {
                p.sigma[s] = sigma;
                p.rgb[(size_t)s * 3] = c[0]; p.rgb[(size_t)s * 3 + 1] = c[1]; p.rgb[(size_t)s * 3 + 2] = c[2];
            }
        }
        umma::fence_before_sync();
        umma::group_sync(bar_id, 128);
    }
#undef REC
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // this group's tile saves are complete before its smem goes away
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_slot, 512);
}

// ---- packing ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_grid_table_kernel(const float2* __restrict__ emb, const int32_t* __restrict__ offsets, const int32_t* __restrict__ first, uint32_t L,
                       uint2* __restrict__ out) {
    __shared__ int32_t s_off[65], s_first[64];
    for (uint32_t i = threadIdx.x; i <= L; i += blockDim.x) s_off[i] = offsets[i];
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) s_first[i] = first[i];
    __syncthreads();
    const uint32_t rows = (uint32_t)s_off[L];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
        uint32_t l = 0;
        while (l + 1 < L && (uint32_t)s_off[l + 1] <= r) ++l;
        const uint32_t lo = (uint32_t)s_off[l], size = (uint32_t)s_off[l + 1] - lo, i = r - lo;
        const float2 a = __ldg(emb + r), b = __ldg(emb + lo + (i + 1 == size ? 0 : i + 1));
        out[(uint32_t)s_first[l] + i] = make_uint2(pack2(a.x, a.y), pack2(b.x, b.y));
    }
}

struct PackMat {
    int src;                 // index into the eight weights
    uint32_t ld, rows, cols; // source leading dimension and valid extent of the (sub)matrix AFTER col0
    uint32_t col0;
    uint32_t n_pad, k_pad;   // destination operand [n_pad x k_pad], interleaved
    uint32_t transpose;      // dst(n, k) = src[k][col0 + n] instead of src[n][col0 + k]
    uint32_t perm_s3;        // source row of logical row r: r < 64 -> r + 1, r == 64 -> 0 (geo_feat rows first, log-density last)
    uint32_t dst;            // byte offset in the blob
    uint32_t which;          // 0 = forward blob, 1 = backward blob
};
// weights: 0 Wa1 [64,96], 1 Wa2 [64,64], 2 Wa3 [2,64], 3 Ws1 [64,65], 4 Ws2 [64,64], 5 Ws3 [65,64], 6 Wc1 [64,84], 7 Wc2 [3,64]
__constant__ PackMat c_pack[17] = {
    {0, 96, 64, 32, 0, 64, 32, 0, 0, B_WA1, 0},  {1, 64, 64, 64, 0, 64, 64, 0, 0, B_WA2, 0},  {2, 64, 2, 64, 0, 16, 64, 0, 0, B_WA3, 0},
    {3, 65, 64, 32, 0, 64, 32, 0, 0, B_WS1A, 0}, {3, 65, 64, 32, 32, 64, 32, 0, 0, B_WS1B, 0}, {4, 64, 64, 64, 0, 64, 64, 0, 0, B_WS2, 0},
    {5, 64, 65, 64, 0, 80, 64, 0, 1, B_WS3, 0},  {6, 84, 64, 80, 0, 64, 80, 0, 0, B_WC1, 0},  {7, 64, 3, 64, 0, 16, 64, 0, 0, B_WC2, 0},
    // backward: dst(n, k) = W[k][col0 + n]
    {7, 64, 3, 64, 0, 64, 16, 1, 0, BW1_C2, 1},  {6, 84, 64, 64, 16, 64, 64, 1, 0, BW1_C1G, 1}, {5, 64, 65, 64, 0, 64, 80, 1, 1, BW1_S3, 1},
    {4, 64, 64, 64, 0, 64, 64, 1, 0, BW1_S2, 1}, {3, 65, 64, 64, 0, 64, 64, 1, 0, BW1_S1, 1},
    {2, 64, 2, 64, 0, 64, 16, 1, 0, BW1_BYTES + BW2_A3, 1}, {1, 64, 64, 64, 0, 64, 64, 1, 0, BW1_BYTES + BW2_A2, 1},
    {0, 96, 64, 32, 0, 32, 64, 1, 0, BW1_BYTES + BW2_A1, 1},
};
struct PackPtrs { const float* w[8]; };

__global__ void __launch_bounds__(256)
pack_head_blobs_kernel(PackPtrs ptrs, uint8_t* __restrict__ fwd, uint8_t* __restrict__ bwd) {
    const PackMat m = c_pack[blockIdx.y];
    uint8_t* base = m.which ? bwd : fwd;
    if (!base) return;
    const float* __restrict__ W = ptrs.w[m.src];
    const uint32_t total = m.n_pad * m.k_pad;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const uint32_t n = e / m.k_pad, k = e - n * m.k_pad;
        uint32_t r = m.transpose ? k : n, c = m.transpose ? n : k;   // logical (row, column) of the source sub-matrix
        float v = 0.f;
        bool ok = c < m.cols;
        if (m.perm_s3) { ok = ok && r <= 64; r = (r < 64) ? r + 1 : 0; }
        else ok = ok && r < m.rows;
        if (ok) v = __ldg(W + (size_t)r * m.ld + m.col0 + c);
        *reinterpret_cast<__half*>(base + m.dst + umma::il_offset(n, k, m.k_pad)) = __float2half_rn(v);
    }
}

}  // namespace

int launch_head_train_fwd(const rn_head_train_desc* d, cudaStream_t st) {
    FwdParams p;
    p.xyzs = d->xyzs; p.dirs = d->dirs; p.M = d->M;
    p.table3 = (const __half*)d->grid3d.table_f16; p.offs3 = d->grid3d.offsets; p.poffs3 = d->grid3d.packed_offsets; p.S3 = d->grid3d.S; p.H3 = d->grid3d.H;
    p.table2 = (const __half*)d->grid2d.table_f16; p.offs2 = d->grid2d.offsets; p.poffs2 = d->grid2d.packed_offsets; p.S2 = d->grid2d.S; p.H2 = d->grid2d.H;
    p.blob = (const uint8_t*)d->fwd_blob; p.consts = d->consts;
    p.sigma = d->sigma; p.rgb = d->rgb; p.ambient = d->ambient; p.sigma_pre = d->sigma_pre;
    p.acts = (uint8_t*)d->acts; p.dy_dx2 = (__half*)d->dy_dx2;
    p.bound = d->bound; p.inv2bound = 1.0f / (2.0f * d->bound);
    p.m_valid = d->m_valid;
    cudaError_t e = cudaFuncSetAttribute(head_train_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FWD_SMEM);
    if (e != cudaSuccess) { set_error("head_train_fwd: cannot reserve %u bytes of shared memory: %s", FWD_SMEM, cudaGetErrorString(e)); return (int)e; }
    const uint32_t n_tiles = (d->M + EVAL_TILE - 1) / EVAL_TILE;
    uint32_t grid = (n_tiles + FWD_GROUPS - 1) / FWD_GROUPS;
    if (grid > RN_NUM_SMS) grid = RN_NUM_SMS;
    head_train_fwd_kernel<<<grid, FWD_GROUPS * 128, FWD_SMEM, st>>>(p);
    return finish_launch("rn_head_train_forward");
}

}  // namespace train
}  // namespace rn

using namespace rn;

extern "C" int rn_pack_grid_table(const float* embeddings, const int32_t* offsets, const int32_t* packed_first, uint32_t L, void* out, void* stream) {
    RN_REQUIRE(embeddings && offsets && packed_first && out, "null pointer");
    RN_REQUIRE(L >= 1 && L <= 64, "num_levels must be in [1, 64]");
    train::pack_grid_table_kernel<<<RN_NUM_SMS * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(embeddings), offsets, packed_first, L,
                                                                                  reinterpret_cast<uint2*>(out));
    return finish_launch("rn_pack_grid_table");
}

extern "C" int rn_pack_head_blobs(const float* const* weights8, void* fwd_blob, void* bwd_blob, void* stream) {
    RN_REQUIRE(weights8 && fwd_blob, "null pointer");
    train::PackPtrs ptrs;
    for (int i = 0; i < 8; ++i) { RN_REQUIRE(weights8[i], "null weight pointer"); ptrs.w[i] = weights8[i]; }
    train::pack_head_blobs_kernel<<<dim3(8, 17), 256, 0, (cudaStream_t)stream>>>(ptrs, (uint8_t*)fwd_blob, (uint8_t*)bwd_blob);
    return finish_launch("rn_pack_head_blobs");
}
