// head_eval.cu -- the fused per-sample network of the head (NeRFNetwork.forward, nerf/network.py:222-283) as ONE
// persistent tcgen05 kernel: for every 128-sample tile
//
//   3-D grid encode (16 lvl x 8 corners, fp16 table)            -> A0  [128 x 32]  fp16, smem
//   ambient MLP  32(+64 audio, hoisted) -> 64 -> 64 -> 2         tcgen05.mma, accumulators in TMEM
//   tanh (fp32)  -> 2-D grid encode (16 lvl x 4 corners)         -> EW  [128 x 32]
//   sigma MLP    64(+eye, hoisted) -> 64 -> 64 -> 65             -> sigma = exp(.), geo_feat [128 x 64]
//   SH(dir) deg 4 (16)                                           -> CIN [128 x 80] = [sh | geo_feat]
//   colour MLP   80(+4 individual code, hoisted) -> 64 -> 3      -> sigmoid
//
// Activations never leave the SM: thread t of a 128-thread tile group owns sample row t of every operand/accumulator
// (TMEM lane t), gathers its own features, and writes fp16 rows straight into the next layer's A operand in the
// interleaved UMMA layout (umma.cuh).  Per-frame-constant inputs (audio code, eye, individual code) are folded into
// fp32 bias vectors once per frame (rn_frame_constants).  All 8 weight matrices (52 KB fp16) are staged once per CTA with
// one TMA bulk copy.  HEAD_GROUPS tile groups per CTA interleave so gathers of one tile overlap MMAs/epilogues of others.
//
// Rounding points follow the reference's fp16 autocast: layer outputs are rounded to fp16 (nn.Linear under autocast),
// the ambient coordinate goes back to fp32 before tanh (network.py:246-247), sigma = exp in fp32 (activation.py:5),
// SH in fp32 then fp16 at the colour layer's input, sigmoid output in fp16.
#include "frame.cuh"
#include "umma.cuh"
#include "sh.cuh"
#include "mlp_tile.cuh"

namespace rn {

namespace {

// blob sub-matrix byte offsets
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

// per-group activation buffers (32 KB).  Aliasing is safe because every stage waits for its MMA (the operand has been read) before
// the next write, which also lets a hidden layer's output overwrite its own input:
//   A0  [128x32]  enc_x           written by the 3-D encode, read by ambient-L1 and sigma-L1
//   H   [128x64]  hidden          every 64-wide layer output, written IN PLACE over the layer's input
//   CIN [128x80]  = A0 + most of H, colour-L1 input [sh | geo_feat], written after sigma-L1 / L3 retired A0 and H
//   EW  [128x32]  enc_w           written by the 2-D encode, read by sigma-L1
// (round 2, late: H0/H1 ping-pong dropped, 40 -> 32 KB per group, so that FIVE groups = 20 warps fit beside the 52 KB of weights)
constexpr uint32_t G_A0 = 0;
constexpr uint32_t G_H = G_A0 + 128 * 32 * 2;
constexpr uint32_t G_CIN = 0;
constexpr uint32_t G_EW = G_H + 128 * 64 * 2;
constexpr uint32_t GROUP_BYTES = G_EW + 128 * 32 * 2;
static_assert(G_EW >= 128 * 80 * 2, "CIN must fit in A0 + H");
constexpr int HEAD_GROUPS = 5;
constexpr uint32_t HEAD_SMEM = HEAD_BLOB_BYTES + HEAD_GROUPS * GROUP_BYTES;
constexpr uint32_t TMEM_COLS_PER_GROUP = 96;         // widest accumulator is 80 columns (sigma-L3); 5 x 96 = 480 of the 512 columns
static_assert(HEAD_GROUPS * TMEM_COLS_PER_GROUP <= 512, "TMEM columns");
// (a register cap through a larger launch bound -- 80 registers, so that the small kernels of other frame lanes fit next to a resident
// CTA -- was measured in round 2: the spills cost more than the co-residency gives, 56.7 vs 52.6 us per launch, 3 200 vs 3 300 frames/s)
#define HEAD_LAUNCH_BOUND (HEAD_GROUPS * 128)

// PROF: the per-phase cycle counters of tools/frame_breakdown.py; compiled out of the production instantiation (the clock reads and
// their bookkeeping were ~3 % of the kernel's instructions even with the counters switched off at run time)
template <bool PROF>
__global__ void __launch_bounds__(HEAD_LAUNCH_BOUND, 1)
head_eval_kernel(HeadEvalParams p, const FrameCur* cur) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ FastLevel lv3[16], lv2[16];
    __shared__ __align__(128) uint8_t s_ones[128 * 16 * 2];        // constant A operand [128 x 16]: columns 0, 1 = 1.0
    __shared__ __align__(128) uint8_t s_biasop[3][64 * 16 * 2];    // B operands [64 x 16]: columns 0, 1 = hi / lo halves of the hoisted term
    __shared__ __align__(8) uint64_t mbar_group[HEAD_GROUPS];
    __shared__ __align__(8) uint64_t mbar_w;
    __shared__ uint32_t tmem_slot;

    long long c_entry = 0;
    if constexpr (PROF) c_entry = clock64();
    const FrameCtl* ctl = &cur->c;
    if (ctl->done) return;
    const uint32_t n_samples = ctl->n_samples;
    const uint32_t n_tiles = (n_samples + EVAL_TILE - 1) / EVAL_TILE;
    // groups in use: all five when the launch is more than one round of four (sustained throughput: 20 warps hide the gather / MMA
    // round trips better, +5-8 % frames/s on one GPU); a launch that fits in one round of four keeps four, because a tile runs faster
    // with fewer co-resident groups and such launches (a rank's share of a ray-sharded frame, the last march iterations) are pure latency
    const uint32_t groups = n_tiles > gridDim.x * 4u ? (uint32_t)HEAD_GROUPS : 4u;
    if (blockIdx.x * groups >= n_tiles) return;

    const uint32_t tid = threadIdx.x, g = tid >> 7, t = tid & 127, warp = tid >> 5;
    uint8_t* s_blob = smem;
    uint8_t* s_grp = smem + HEAD_BLOB_BYTES + g * GROUP_BYTES;

    // ---- one-time CTA setup: barriers, weights via TMA bulk copy, TMEM, level geometry, biases
    if (tid == 0) {
        for (int i = 0; i < HEAD_GROUPS; ++i) umma::mbar_init(&mbar_group[i], 1);
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
    } else if (tid < 128 + 192) {   // hoisted per-frame terms (audio code / eye / individual code), split so that hi + lo carries 22 bits
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
    // the weight blob (52 KB by TMA) is first needed by the first MMA, one 3-D encode (~6 us) from here: it is awaited there, not here
    bool weights_ready = false;
    long long c_setup = 0;
    if constexpr (PROF) c_setup = clock64() - c_entry;

    const uint32_t tmem_acc = tmem_slot + g * TMEM_COLS_PER_GROUP;       // this group's accumulator columns
    const uint32_t tmem_row = tmem_acc + (((warp & 3u) * 32u) << 16);    // this warp's lane quarter
    uint64_t* mbar = &mbar_group[g];
    uint32_t phase = 0;
    const uint32_t bar_id = 1 + g;
    uint8_t* sA0 = s_grp + G_A0;
    uint8_t* sEW = s_grp + G_EW;
    uint8_t* sH0 = s_grp + G_H;      // one hidden buffer: the two names only keep the stage list below readable
    uint8_t* sH1 = s_grp + G_H;
    uint8_t* sCIN = s_grp + G_CIN;
    const uint32_t aA0 = umma::smem_u32(sA0), aEW = umma::smem_u32(sEW), aH0 = umma::smem_u32(sH0), aH1 = umma::smem_u32(sH1),
                   aCIN = umma::smem_u32(sCIN);
    const uint32_t aW = umma::smem_u32(s_blob);
    const uint32_t aOnes = umma::smem_u32(s_ones), aB0 = umma::smem_u32(s_biasop[0]), aB1 = umma::smem_u32(s_biasop[1]), aB2 = umma::smem_u32(s_biasop[2]);
    const uint2* table3 = reinterpret_cast<const uint2*>(p.table3);
    const uint2* table2 = reinterpret_cast<const uint2*>(p.table2);

    long long c_enc3 = 0, c_enc2 = 0, c_mma = 0, c_epi = 0, c_tile = 0, c_last = 0;
    const bool prof = PROF && p.prof != nullptr && t == 0;
#define RN_TICK(acc) if constexpr (PROF) { if (prof) { const long long now = clock64(); acc += now - c_last; c_last = now; } }
    for (uint32_t tile = g < groups ? blockIdx.x * groups + g : n_tiles; tile < n_tiles; tile += gridDim.x * groups) {
        if constexpr (PROF) { if (prof) c_last = clock64(); }
        const long long c_start = c_last;
        const uint32_t s = tile * EVAL_TILE + t;
        const bool valid = s < n_samples;
        const float4 smp = valid ? __ldg(p.samples + s) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t ray = valid ? (uint32_t)__float_as_int(smp.w) : 0u;

        // ---- 3-D encode -> A0
        {
            float x[3] = {__fmul_rn(__fadd_rn(smp.x, p.bound), p.inv2bound), __fmul_rn(__fadd_rn(smp.y, p.bound), p.inv2bound),
                          __fmul_rn(__fadd_rn(smp.z, p.bound), p.inv2bound)};
            fast_encode<3>(x, table3, lv3, sA0, t, 32, 0);
        }
        RN_TICK(c_enc3)
        if (!weights_ready) { umma::mbar_wait(&mbar_w, 0); weights_ready = true; }
        // ---- ambient L1: A0 [128x32] x WA1 -> 64, + hoisted audio term, ReLU -> H0
        mma_stage(tmem_acc, aA0, 32, 0, aW + B_WA1, 32, 0, 0, 0, aOnes, aB0, 64, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        epilogue_to_operand<2>(tmem_row, 0, true, sH0, t, 64, 0);
        RN_TICK(c_epi)
        // ---- ambient L2: H0 -> H1
        mma_stage(tmem_acc, aH0, 64, 0, aW + B_WA2, 64, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        epilogue_to_operand<2>(tmem_row, 0, true, sH1, t, 64, 0);
        RN_TICK(c_epi)
        // ---- ambient L3 (N padded to 16) -> tanh -> 2-D encode -> EW (first half of H0)
        mma_stage(tmem_acc, aH1, 64, 0, aW + B_WA3, 64, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        {
            uint32_t v[16];
            umma::tmem_ld16(tmem_row, v);
            umma::tmem_ld_wait();
            const float a0 = tanhf(__half2float(__float2half_rn(__uint_as_float(v[0]))));
            const float a1 = tanhf(__half2float(__float2half_rn(__uint_as_float(v[1]))));
            float x[2] = {__fmul_rn(__fadd_rn(a0, 1.0f), 0.5f), __fmul_rn(__fadd_rn(a1, 1.0f), 0.5f)};
            fast_encode<2>(x, table2, lv2, sEW, t, 32, 0);
        }
        RN_TICK(c_enc2)
        // ---- sigma L1: [enc_x | enc_w] as two K = 32 slabs, + hoisted eye term, ReLU -> H1
        mma_stage(tmem_acc, aA0, 32, 0, aW + B_WS1A, 32, aEW, aW + B_WS1B, 32, aOnes, aB1, 64, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        epilogue_to_operand<2>(tmem_row, 0, true, sH1, t, 64, 0);
        RN_TICK(c_epi)
        // ---- sigma L2: H1 -> H0
        mma_stage(tmem_acc, aH1, 64, 0, aW + B_WS2, 64, 0, 0, 0, 0, 0, 64, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        epilogue_to_operand<2>(tmem_row, 0, true, sH0, t, 64, 0);
        RN_TICK(c_epi)
        // ---- sigma L3: rows permuted on the host so columns 0..63 = geo_feat, column 64 = log-density
        mma_stage(tmem_acc, aH0, 64, 0, aW + B_WS3, 64, 0, 0, 0, 0, 0, 80, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        float sigma;
        {
            epilogue_to_operand<2>(tmem_row, 0, false, sCIN, t, 80, 16);  // geo_feat -> CIN columns 16..79
            uint32_t v[16];
            umma::tmem_ld16(tmem_row + 64, v);
            umma::tmem_ld_wait();
            sigma = expf(__half2float(__float2half_rn(__uint_as_float(v[0]))));
            // SH(dir), degree 4 -> CIN columns 0..15
            const float* d = p.rays_d + (size_t)ray * 3;
            float Y[16];
            sh_eval<4, false>(__ldg(d), __ldg(d + 1), __ldg(d + 2), Y, nullptr, nullptr, nullptr);
            *reinterpret_cast<uint4*>(sCIN + umma::il_offset(t, 0, 80)) =
                make_uint4(pack2(Y[0], Y[1]), pack2(Y[2], Y[3]), pack2(Y[4], Y[5]), pack2(Y[6], Y[7]));
            *reinterpret_cast<uint4*>(sCIN + umma::il_offset(t, 8, 80)) =
                make_uint4(pack2(Y[8], Y[9]), pack2(Y[10], Y[11]), pack2(Y[12], Y[13]), pack2(Y[14], Y[15]));
        }
        RN_TICK(c_epi)
        // ---- colour L1: CIN [sh | geo] K = 80 -> 64, + hoisted individual-code term, ReLU -> H0
        mma_stage(tmem_acc, aCIN, 80, 0, aW + B_WC1, 80, 0, 0, 0, aOnes, aB2, 64, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
        epilogue_to_operand<2>(tmem_row, 0, true, sH0, t, 64, 0);
        RN_TICK(c_epi)
        // ---- colour L2 (N padded to 16) -> sigmoid
        mma_stage(tmem_acc, aH0, 64, 0, aW + B_WC2, 64, 0, 0, 0, 0, 0, 16, mbar, phase, bar_id, t);
        RN_TICK(c_mma)
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
            if (valid) p.evals[s] = make_float4(sigma, c[0], c[1], c[2]);
        }
        // the next tile's first MMA overwrites the accumulator: retire this tile's TMEM reads first
        umma::fence_before_sync();
        umma::group_sync(bar_id, 128);
        if constexpr (PROF) { if (prof) c_tile += clock64() - c_start; }
    }
    if (PROF && prof) {
        atomicAdd(p.prof + 0, (unsigned long long)c_enc3);
        atomicAdd(p.prof + 1, (unsigned long long)c_enc2);
        atomicAdd(p.prof + 2, (unsigned long long)c_mma);
        atomicAdd(p.prof + 3, (unsigned long long)c_epi);
        atomicAdd(p.prof + 4, (unsigned long long)c_tile);
        atomicAdd(p.prof + 5, 1ull);
        atomicAdd(p.prof + 6, (unsigned long long)c_setup);
    }
#undef RN_TICK

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_slot, 512);
}

}  // namespace

int launch_head_eval(const HeadEvalParams& p, const FrameCur* cur, uint32_t max_tiles, cudaStream_t st) {
    // the opt-in for > 48 KB of dynamic shared memory is a per-device function attribute: set it before every launch (the call is
    // cheap and capture-safe) instead of caching a process-wide flag that would be wrong for a second device
    cudaError_t e = p.prof ? cudaFuncSetAttribute(head_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM)
                           : cudaFuncSetAttribute(head_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_SMEM);
    if (e != cudaSuccess) { set_error("head_eval: cannot reserve %u bytes of shared memory: %s", HEAD_SMEM, cudaGetErrorString(e)); return (int)e; }
    uint32_t grid = (max_tiles + 3) / 4;      // sized for four groups per CTA; the kernel switches to five when that is more than one round
    if (grid > RN_NUM_SMS) grid = RN_NUM_SMS;
    if (grid == 0) grid = 1;
    if (p.prof) head_eval_kernel<true><<<grid, HEAD_GROUPS * 128, HEAD_SMEM, st>>>(p, cur);
    else head_eval_kernel<false><<<grid, HEAD_GROUPS * 128, HEAD_SMEM, st>>>(p, cur);
    return finish_launch("head_eval");
}

}  // namespace rn
