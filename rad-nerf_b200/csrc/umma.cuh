// umma.cuh -- hand-written tcgen05 / TMEM / mbarrier helpers for the fused MLP kernels (sm_100a).
//
// Operand convention used everywhere in this library ("interleaved", i.e. UMMA LayoutType::SWIZZLE_NONE, K-major):
// a [rows x K] fp16 operand is stored as 8-row x 16-byte core matrices,
//      byte_offset(r, k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2,   LBO = 128, SBO = 128 * (K / 8)
// so the K/8 core matrices of one 8-row group are contiguous.  A thread that owns row r writes its K halves as K/8
// 16-byte stores; lanes r..r+7 cover 128 contiguous bytes per store phase (bank-conflict free).  nn.Linear weights
// [out, in] are exactly the B operand (N x K, K-major) in this layout.
// One tcgen05.mma consumes K = 16 (two core matrices along K): the descriptor for k-step s starts at base + s * 2 * LBO.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace rn {
namespace umma {

constexpr uint32_t LBO_BYTES = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, k) of an interleaved [rows x K] fp16 operand
__host__ __device__ __forceinline__ uint32_t il_offset(uint32_t r, uint32_t k, uint32_t K) {
    return (r >> 3) * (LBO_BYTES * (K >> 3)) + (k >> 3) * LBO_BYTES + (r & 7) * 16 + (k & 7) * 2;
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type [61,64) = 0 (no swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t K) {
    const uint32_t sbo = LBO_BYTES * (K >> 3);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(LBO_BYTES >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor for kind::f16, A/B = fp16 K-major, D = fp32, M = 128 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4)              // c_format = F32
           | (0u << 7)            // a_format = F16
           | (0u << 10)           // b_format = F16
           | (0u << 15) | (0u << 16)  // K-major A and B
           | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[128 x N] (+)= A[128 x K] * B[N x K]^T for K a multiple of 16; issued by ONE thread
// A columns [k_begin, k_begin + k_count) of an operand stored with row length K_a_total; B columns [0, k_count) of K_b_total
__device__ __forceinline__ void gemm_issue(uint32_t tmem_d, uint32_t a_smem, uint32_t b_smem, uint32_t K_a_total,
                                           uint32_t K_b_total, uint32_t k_begin, uint32_t k_count, uint32_t N, bool accumulate) {
    const uint32_t idesc = make_idesc_f16(128, N);
    for (uint32_t s = 0; s < k_count; s += 16) {
        const uint64_t da = make_desc(a_smem + ((k_begin + s) >> 3) * LBO_BYTES, K_a_total);
        const uint64_t db = make_desc(b_smem + (s >> 3) * LBO_BYTES, K_b_total);
        mma_f16_ss(tmem_d, da, db, idesc, (accumulate || s > 0) ? 1u : 0u);
    }
}

// ---- MN-major ("transposed") operands ---------------------------------------------------------------------------------
// The SAME interleaved [rows x Kf] tile read with rows as the MMA's K dimension and the Kf features as its M (or N) dimension:
// a core matrix is then 8 K-rows x 16 bytes of 8 consecutive M/N elements, which is exactly how the tile already sits in shared
// memory.  cute's canonical form (mma_traits_sm100.hpp, make_umma_desc<Major::MN>, INTERLEAVE, units of 16 bytes):
//      ((1,n),(8,k)) : ((X,SBO),(1,LBO))   ->   SBO = stride between core matrices along M/N = 128 bytes,
//                                               LBO = stride between 8-row groups along K   = 16 * Kf bytes
// One tcgen05.mma consumes K = 16 rows = two 8-row groups; k-step s starts at base + s * 2 * LBO.
// This is how the weight gradient dW = dZ^T A runs on tensor cores without transposing anything: samples are K.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t Kf) {
    const uint32_t lbo = 16u * Kf;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(LBO_BYTES >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D[128 x N] (+)= X^T Y over `rows` tile rows (a multiple of 16): X = interleaved [rows x Kx] tile whose feature columns
// [mx0, mx0 + 128) become D's rows (columns past Kx read whatever follows in shared memory: those D rows are garbage and must be
// ignored), Y = interleaved [rows x Ky] tile whose columns [ny0, ny0 + N) become D's columns.
__device__ __forceinline__ void gemm_issue_mn(uint32_t tmem_d, uint32_t x_smem, uint32_t Kx, uint32_t mx0, uint32_t y_smem, uint32_t Ky,
                                              uint32_t ny0, uint32_t N, uint32_t rows, bool accumulate) {
    const uint32_t idesc = make_idesc_f16(128, N) | (1u << 15) | (1u << 16);   // A and B both MN-major
    for (uint32_t s = 0; s < rows; s += 16) {
        const uint64_t da = make_desc_mn(x_smem + (mx0 >> 3) * LBO_BYTES + (s >> 3) * 16u * Kx, Kx);
        const uint64_t db = make_desc_mn(y_smem + (ny0 >> 3) * LBO_BYTES + (s >> 3) * 16u * Ky, Ky);
        mma_f16_ss(tmem_d, da, db, idesc, (accumulate || s > 0) ? 1u : 0u);
    }
}

__device__ __forceinline__ void commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy st.shared -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation (one full warp) ---------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- TMEM -> registers: warp w (w % 4) may only touch lanes [32*(w%4), +32); each thread gets its lane's columns -----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(mbar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the waiting thread is parked by the hardware until the phase completes or ~the hint elapses, so a
// wait costs a handful of instructions instead of a spin (ncu, round 2: the plain try_wait / counter loop was 11 % of all warp
// instructions of head_eval_kernel, ~12 iterations per wait, issued in competition with the other tile groups' useful work)
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* mbar, uint32_t parity, uint32_t hint_ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(mbar)), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// bounded: a protocol bug traps instead of hanging the GPU (2^16 waits of up to 20 us each)
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    if (mbar_try_wait_hint(mbar, parity, 20000u)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait_hint(mbar, parity, 20000u)) {
        if (++spins > (1u << 16)) __trap();
    }
}

// sub-CTA barrier for one 128-thread tile group
__device__ __forceinline__ void group_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace umma
}  // namespace rn
