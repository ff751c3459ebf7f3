// selftest.cu -- device self-tests of the hand-written tcgen05 plumbing (umma.cuh): one 128 x N x K fp16 GEMM tile with
// operands staged by the threads themselves in the interleaved (no-swizzle, K-major) layout, accumulator in TMEM.
// Used by tests/test_gpu_fused.py to validate descriptors and layouts in isolation from the fused kernels.
#include "common.cuh"
#include "umma.cuh"

namespace rn {
namespace {

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const __half* __restrict__ A, const __half* __restrict__ W, float* __restrict__ out, uint32_t K, uint32_t N) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    uint8_t* sA = smem;                      // [128 x K] interleaved
    uint8_t* sW = smem + 128 * K * 2;        // [N x K] interleaved
    const uint32_t t = threadIdx.x, warp = t >> 5;

    for (uint32_t k = 0; k < K; k += 8)
        *reinterpret_cast<uint4*>(sA + umma::il_offset(t, k, K)) = *reinterpret_cast<const uint4*>(A + (size_t)t * K + k);
    for (uint32_t r = t; r < N; r += 128)
        for (uint32_t k = 0; k < K; k += 8)
            *reinterpret_cast<uint4*>(sW + umma::il_offset(r, k, K)) = *reinterpret_cast<const uint4*>(W + (size_t)r * K + k);
    if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 128);
    if (t == 0) {
        umma::mbar_init(&mbar, 1);
        umma::fence_mbar_init();
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    if (t == 0) {
        umma::gemm_issue(tmem_base, umma::smem_u32(sA), umma::smem_u32(sW), K, K, 0, K, N, false);
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
    for (uint32_t n0 = 0; n0 < N; n0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(lane_base + n0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)t * N + n0 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 128);
}

// out[m, n] = sum_s X[s, m] * Y[s, n]: both operands MN-major views of interleaved [128 x K] tiles (umma::gemm_issue_mn) --
// the weight-gradient contraction of the fused training step, samples as the MMA's K dimension.
__global__ void __launch_bounds__(128)
umma_mn_selftest_kernel(const __half* __restrict__ X, const __half* __restrict__ Y, float* __restrict__ out, uint32_t Kx, uint32_t Ky,
                        uint32_t passes) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_slot;
    uint8_t* sX = smem;                      // [128 x Kx] interleaved
    uint8_t* sY = smem + 128 * Kx * 2;       // [128 x Ky] interleaved (+ 4 KB of zeroed slack behind it: M is padded to 128)
    const uint32_t t = threadIdx.x, warp = t >> 5;
    for (uint32_t k = 0; k < Kx; k += 8)
        *reinterpret_cast<uint4*>(sX + umma::il_offset(t, k, Kx)) = *reinterpret_cast<const uint4*>(X + (size_t)t * Kx + k);
    for (uint32_t k = 0; k < Ky; k += 8)
        *reinterpret_cast<uint4*>(sY + umma::il_offset(t, k, Ky)) = *reinterpret_cast<const uint4*>(Y + (size_t)t * Ky + k);
    for (uint32_t i = t; i < 4096 / 16; i += 128) reinterpret_cast<uint4*>(sY + 128 * Ky * 2)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) umma::tmem_alloc(&tmem_base_slot, 128);
    if (t == 0) {
        umma::mbar_init(&mbar, 1);
        umma::fence_mbar_init();
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem_base = tmem_base_slot;
    if (t == 0) {
        for (uint32_t p = 0; p < passes; ++p)   // passes > 1: accumulation across "tiles" in TMEM, as the persistent backward does
            umma::gemm_issue_mn(tmem_base, umma::smem_u32(sX), Kx, 0, umma::smem_u32(sY), Ky, 0, Ky, 128, p > 0);
        umma::commit(&mbar);
    }
    umma::mbar_wait(&mbar, 0);
    umma::fence_after_sync();
    const uint32_t lane_base = tmem_base + ((warp * 32u) << 16);
    for (uint32_t n0 = 0; n0 < Ky; n0 += 16) {
        uint32_t v[16];
        umma::tmem_ld16(lane_base + n0, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)t * Ky + n0 + j] = __uint_as_float(v[j]);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem_base, 128);
}

}  // namespace
}  // namespace rn

using namespace rn;

// X [128,Kx], Y [128,Ky] fp16 row-major; out [128,Ky] fp32: rows m < Kx hold passes * sum_s X[s,m] Y[s,n], rows >= Kx are undefined
extern "C" int rn_selftest_umma_mn(const void* X, const void* Y, float* out, uint32_t Kx, uint32_t Ky, uint32_t passes, void* stream) {
    RN_REQUIRE(X && Y && out, "null pointer");
    RN_REQUIRE(Kx % 8 == 0 && Kx >= 8 && Kx <= 128, "Kx must be a multiple of 8 in [8, 128]");
    RN_REQUIRE(Ky % 16 == 0 && Ky >= 16 && Ky <= 128, "Ky must be a multiple of 16 in [16, 128]");
    RN_REQUIRE(passes >= 1 && passes <= 64, "passes in [1, 64]");
    const size_t smem = (size_t)128 * (Kx + Ky) * 2 + 4096;
    cudaFuncSetAttribute(umma_mn_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_mn_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half*)X, (const __half*)Y, out, Kx, Ky, passes);
    return finish_launch("rn_selftest_umma_mn");
}

// A [128,K] fp16 row-major, W [N,K] fp16 row-major (nn.Linear weight layout), out [128,N] fp32 = A @ W^T
extern "C" int rn_selftest_umma(const void* A, const void* W, float* out, uint32_t K, uint32_t N, void* stream) {
    RN_REQUIRE(A && W && out, "null pointer");
    RN_REQUIRE(K % 16 == 0 && K >= 16 && K <= 256, "K must be a multiple of 16 in [16, 256]");
    RN_REQUIRE(N % 16 == 0 && N >= 16 && N <= 128, "N must be a multiple of 16 in [16, 128]");
    const size_t smem = (size_t)(128 + N) * K * 2;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half*)A, (const __half*)W, out, K, N);
    return finish_launch("rn_selftest_umma");
}
