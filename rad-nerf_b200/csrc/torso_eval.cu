// torso_eval.cu -- the fused 2-D torso model (NeRFNetwork.forward_torso, nerf/network.py:188-219) on the compacted list of
// masked pixels, one persistent tcgen05 kernel with the same row-per-thread scheme as head_eval.cu:
//
//   x = bg_coord * torso_shrink;  freq encode (2 -> 42, __sinf)              -> F   [128 x 48]  (42 + zero pad)
//   deform MLP 42(+54 pose +8 code, hoisted) -> 64 -> 64 -> 2                 -> dx
//   x' = clamp(x + dx, -1, 1);  2-D grid encode (16 lvl x 4 corners)          -> TIN [128 x 80] = [grid 32 | freq 42 | pad]
//   torso MLP 74(+62 hoisted) -> 32 -> 32 -> 4                                -> sigmoid -> (alpha, rgb)
#include "frame.cuh"
#include "umma.cuh"
#include "gridencoder_impl.cuh"
#include "mlp_tile.cuh"

namespace rn {


namespace {


constexpr uint32_t T_WD1 = 0;                          // [64 x 48]
constexpr uint32_t T_WD2 = T_WD1 + 64 * 48 * 2;        // [64 x 64]
constexpr uint32_t T_WD3 = T_WD2 + 64 * 64 * 2;        // [16 x 64]
constexpr uint32_t T_WT1 = T_WD3 + 16 * 64 * 2;        // [32 x 80]
constexpr uint32_t T_WT2 = T_WT1 + 32 * 80 * 2;        // [32 x 32]
constexpr uint32_t T_WT3 = T_WT2 + 32 * 32 * 2;        // [16 x 32]
static_assert(T_WT3 + 16 * 32 * 2 == TORSO_BLOB_BYTES, "blob layout");

constexpr uint32_t G_F = 0;                            // [128 x 48]
constexpr uint32_t G_H0 = G_F + 128 * 48 * 2;          // [128 x 64]
constexpr uint32_t G_H1 = G_H0 + 128 * 64 * 2;         // [128 x 64]
constexpr uint32_t G_TIN = G_H1 + 128 * 64 * 2;        // [128 x 80]
constexpr uint32_t GROUP_BYTES = G_TIN + 128 * 80 * 2;
constexpr uint32_t TORSO_SMEM = TORSO_BLOB_BYTES + EVAL_GROUPS * GROUP_BYTES;
constexpr uint32_t TMEM_COLS_PER_GROUP = 64;

__global__ void __launch_bounds__(EVAL_GROUPS * 128, 1)
torso_eval_kernel(TorsoEvalParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ FastLevel lv[16];
    __shared__ __align__(8) uint64_t mbar_group[EVAL_GROUPS];
    __shared__ __align__(8) uint64_t mbar_w;
    __shared__ uint32_t tmem_slot;

    const uint32_t n_pix = *p.n_pix;
    const uint32_t n_tiles = (n_pix + EVAL_TILE - 1) / EVAL_TILE;
    if (blockIdx.x * EVAL_GROUPS >= n_tiles) return;

    const uint32_t tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = tid >> 5;
    uint8_t* s_blob = smem;
    uint8_t* s_grp = smem + TORSO_BLOB_BYTES + g * GROUP_BYTES;

    if (tid == 0) {
        for (int i = 0; i < EVAL_GROUPS; ++i) umma::mbar_init(&mbar_group[i], 1);
        umma::mbar_init(&mbar_w, 1);
        umma::fence_mbar_init();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(&mbar_w)), "r"(TORSO_BLOB_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(s_blob)),
                     "l"(p.blob), "r"(TORSO_BLOB_BYTES), "r"(umma::smem_u32(&mbar_w))
                     : "memory");
    }
    if (warp == 1) umma::tmem_alloc(&tmem_slot, 256);
    if (tid >= 64 && tid < 80) {
        grid::LevelMeta m;
        grid::make_level_meta(m, tid - 64, p.offs, p.S, p.H, 2, 1, false);
        if (!make_fast_level(lv[tid - 64], m, (uint32_t)__ldg(p.poffs + (tid - 64)))) __trap();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    umma::mbar_wait(&mbar_w, 0);
    // The hoisted per-frame terms (pose / individual code through deform-L1 and torso-L1) ride through the MMA: the frequency
    // encoding has 6 padding columns (42..47), two of them carry 1.0 and the matching weight columns are patched here with
    // the hi / lo halves of the term (22 bits).  No bias add in the epilogues.
    if (tid < 96) {
        const bool deform = tid < 64;
        const uint32_t n = deform ? tid : tid - 64;
        const float b = __ldg(p.consts + tid);
        const float hi = __half2float(__float2half_rn(b));
        uint8_t* w = s_blob + (deform ? T_WD1 + umma::il_offset(n, 42, 48) : T_WT1 + umma::il_offset(n, 74, 80));
        *reinterpret_cast<uint32_t*>(w) = pack2(hi, b - hi);
    }
    umma::fence_async_smem();
    __syncthreads();

    const uint32_t tmem_acc = tmem_slot + g * TMEM_COLS_PER_GROUP;
    const uint32_t tmem_row = tmem_acc + (((warp & 3u) * 32u) << 16);
    uint64_t* mbar = &mbar_group[g];
    uint32_t phase = 0;
    const uint32_t bar_id = 1 + g;
    uint8_t* sF = s_grp + G_F;
    uint8_t* sH0 = s_grp + G_H0;
    uint8_t* sH1 = s_grp + G_H1;
    uint8_t* sTIN = s_grp + G_TIN;
    const uint32_t aF = umma::smem_u32(sF), aH0 = umma::smem_u32(sH0), aH1 = umma::smem_u32(sH1), aTIN = umma::smem_u32(sTIN);
    const uint32_t aW = umma::smem_u32(s_blob);

    for (uint32_t tile = blockIdx.x * EVAL_GROUPS + g; tile < n_tiles; tile += gridDim.x * EVAL_GROUPS) {
        const uint32_t k = tile * EVAL_TILE + t;
        const bool valid = k < n_pix;
        const int32_t pixel = valid ? __ldg(p.pix + k) : 0;
        const float x0 = __fmul_rn(__ldg(p.bg_coords + (size_t)pixel * 2), p.shrink);
        const float x1 = __fmul_rn(__ldg(p.bg_coords + (size_t)pixel * 2 + 1), p.shrink);

        // ---- frequency encoding (freqencoder.cu:30-58 order: x, then per octave [sin x0, sin x1, cos x0, cos x1]) -> F and TIN[32..]
        {
            float e[48];
            e[0] = x0; e[1] = x1;
#pragma unroll
            for (int f = 0; f < 10; ++f) {
                const float a0 = scalbnf(x0, f), a1 = scalbnf(x1, f);
                e[2 + 4 * f + 0] = __sinf(a0 + 0.0f);
                e[2 + 4 * f + 1] = __sinf(a1 + 0.0f);
                e[2 + 4 * f + 2] = __sinf(a0 + 1.5707963705062866f);
                e[2 + 4 * f + 3] = __sinf(a1 + 1.5707963705062866f);
            }
#pragma unroll
            for (int j = 42; j < 48; ++j) e[j] = j < 44 ? 1.0f : 0.f;   // constant-one columns: carry the hoisted terms
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                const uint4 q = make_uint4(pack2(e[8 * c], e[8 * c + 1]), pack2(e[8 * c + 2], e[8 * c + 3]),
                                           pack2(e[8 * c + 4], e[8 * c + 5]), pack2(e[8 * c + 6], e[8 * c + 7]));
                *reinterpret_cast<uint4*>(sF + umma::il_offset(t, 8 * c, 48)) = q;
                *reinterpret_cast<uint4*>(sTIN + umma::il_offset(t, 32 + 8 * c, 80)) = q;
            }
        }
                // ---- deform L1 (K = 48) -> 64
        mma_stage(tmem_acc, aF, 48, 0, aW + T_WD1, 48, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t);
        epilogue_to_operand<2>(tmem_row, 0, true, sH0, t, 64, 0);
                // ---- deform L2
        mma_stage(tmem_acc, aH0, 64, 0, aW + T_WD2, 64, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t);
        epilogue_to_operand<2>(tmem_row, 0, true, sH1, t, 64, 0);
                // ---- deform L3 (N padded to 16) -> dx -> deformed coordinate -> 2-D grid encode -> TIN[0..31]
        mma_stage(tmem_acc, aH1, 64, 0, aW + T_WD3, 64, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t);
        {
            uint32_t v[16];
            umma::tmem_ld16(tmem_row, v);
            umma::tmem_ld_wait();
            const float dx0 = __half2float(__float2half_rn(__uint_as_float(v[0])));
            const float dx1 = __half2float(__float2half_rn(__uint_as_float(v[1])));
            const float y0 = fminf(fmaxf(__fadd_rn(x0, dx0), -1.f), 1.f), y1 = fminf(fmaxf(__fadd_rn(x1, dx1), -1.f), 1.f);
            float x[2] = {__fmul_rn(__fadd_rn(y0, 1.0f), 0.5f), __fmul_rn(__fadd_rn(y1, 1.0f), 0.5f)};
            fast_encode<2>(x, reinterpret_cast<const uint2*>(p.table), lv, sTIN, t, 80, 0);
        }
        // ---- torso L1 (K = 80) -> 32
        mma_stage(tmem_acc, aTIN, 80, 0, aW + T_WT1, 80, 0, 0, 0, 0, 0, 32, mbar, phase, bar_id, t);
        epilogue_to_operand<1>(tmem_row, 0, true, sH0, t, 32, 0);
                // ---- torso L2 (K = 32) -> 32
        mma_stage(tmem_acc, aH0, 32, 0, aW + T_WT2, 32, 0, 0, 0, 0, 0, 32, mbar, phase, bar_id, t);
        epilogue_to_operand<1>(tmem_row, 0, true, sH1, t, 32, 0);
                // ---- torso L3 (N padded to 16) -> sigmoid -> (alpha, rgb)
        mma_stage(tmem_acc, aH1, 32, 0, aW + T_WT3, 32, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t);
        {
            uint32_t v[16];
            umma::tmem_ld16(tmem_row, v);
            umma::tmem_ld_wait();
            float c[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float h = __half2float(__float2half_rn(__uint_as_float(v[j])));
                c[j] = __half2float(__float2half_rn(1.0f / (1.0f + expf(-h))));
            }
            if (valid) p.out[k] = make_float4(c[0], c[1], c[2], c[3]);
        }
        umma::fence_before_sync();
        umma::group_sync(bar_id, 128);
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_slot, 256);
}

}  // namespace

int launch_torso_eval(const TorsoEvalParams& p, uint32_t max_tiles, cudaStream_t st) {
    static bool configured[64] = {};         // per device: the shared-memory opt-in is an attribute of the function ON a device
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    if (dev_id < 0 || dev_id >= 64 || !configured[dev_id]) {
        cudaError_t e = cudaFuncSetAttribute(torso_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TORSO_SMEM);
        if (e != cudaSuccess) { set_error("torso_eval: cannot reserve %u bytes of shared memory: %s", TORSO_SMEM, cudaGetErrorString(e)); return (int)e; }
        if (dev_id >= 0 && dev_id < 64) configured[dev_id] = true;
    }
    uint32_t grid = (max_tiles + EVAL_GROUPS - 1) / EVAL_GROUPS;
    if (grid > RN_NUM_SMS) grid = RN_NUM_SMS;
    if (grid == 0) grid = 1;
    torso_eval_kernel<<<grid, EVAL_GROUPS * 128, TORSO_SMEM, st>>>(p);
    return finish_launch("torso_eval");
}

}  // namespace rn
