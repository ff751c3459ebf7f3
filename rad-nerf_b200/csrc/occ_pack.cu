// occ_pack.cu -- the occupancy bitfield re-packed for shared memory (north_star: "ray marching ... against a bitfield staged in shared
// memory"; replaces the global-memory probes of kernel_march_rays, raymarching.cu:827-929, in the fused frame).
//
// The density bitfield is C x H^3 bits in Morton order: 256 KiB per cascade at H = 128, more than a CTA's shared memory.  But a
// talking head fills a small box of it (~44 x 31 x 48 cells of 128^3), and every cell outside the bounding box of the occupied cells
// is empty by definition.  rn_occupancy_pack extracts, per cascade, that box as a LINEAR bit array (x fastest; ~8 KiB for the
// synthetic head) plus its cell bounds; the fused marcher copies it into shared memory once per CTA and answers every DDA probe
// from there -- no dependent L2 round trip (~0.5 us each, 15-20 per ray and iteration) and no Morton encode per probe.  A grid whose
// boxes do not fit RN_OCC_PACK_MAX_BYTES keeps the global-memory path (usable = 0).
#include "common.cuh"
#include "occ_pack.cuh"
#include <float.h>

namespace rn {
namespace {

__global__ void occ_init_kernel(OccPack* pk, uint32_t C) {
    const uint32_t l = threadIdx.x;
    if (l == 0) { pk->n_levels = (int32_t)C; pk->total_words = 0; pk->usable = 0; pk->pad = 0; }
    if (l < OCC_MAX_LEVELS) {
        for (int a = 0; a < 3; ++a) { pk->lv[l].lo[a] = INT_MAX; pk->lv[l].dim[a] = -1; }   // dim holds the running max until occ_layout_kernel
        pk->lv[l].word_off = 0; pk->lv[l].pad = 0;
    }
}

// pass 1: per cascade, min / max cell coordinates over the set bits
__global__ void __launch_bounds__(256)
occ_bounds_kernel(const uint8_t* __restrict__ bitfield, uint32_t C, uint32_t H, OccPack* pk) {
    const uint32_t bytes_per_level = H * H * H / 8;
    const uint32_t total = C * bytes_per_level;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {-1, -1, -1};
    uint32_t my_level = 0xffffffffu;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        uint32_t b = __ldg(bitfield + i);
        if (!b) continue;
        const uint32_t level = i / bytes_per_level;
        if (level != my_level) {   // a thread crosses a level boundary at most C - 1 times: flush what it has
            if (my_level != 0xffffffffu && hi[0] >= 0)
                for (int a = 0; a < 3; ++a) { atomicMin(&pk->lv[my_level].lo[a], lo[a]); atomicMax(&pk->lv[my_level].dim[a], hi[a]); }
            my_level = level;
            for (int a = 0; a < 3; ++a) { lo[a] = INT_MAX; hi[a] = -1; }
        }
        const uint32_t base = (i - level * bytes_per_level) * 8;
        while (b) {
            const uint32_t m = base + (uint32_t)__ffs((int)b) - 1;
            b &= b - 1;
            const int c[3] = {(int)compact3(m), (int)compact3(m >> 1), (int)compact3(m >> 2)};
            for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], c[a]); hi[a] = max(hi[a], c[a]); }
        }
    }
    if (my_level != 0xffffffffu && hi[0] >= 0)
        for (int a = 0; a < 3; ++a) { atomicMin(&pk->lv[my_level].lo[a], lo[a]); atomicMax(&pk->lv[my_level].dim[a], hi[a]); }
}

// pass 2 (one thread): box dimensions, word offsets, the world-space box of all occupied cells inflated by one cell per side
__global__ void occ_layout_kernel(OccPack* pk, uint32_t C, uint32_t H, float bound, uint32_t max_words) {
    if (threadIdx.x || blockIdx.x) return;
    float wlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, whi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    uint32_t words = 0;
    for (uint32_t l = 0; l < C; ++l) {
        OccLevel& L = pk->lv[l];
        if (L.dim[0] < 0) {
            for (int a = 0; a < 3; ++a) { L.lo[a] = 0; L.dim[a] = 0; }
            L.word_off = (int32_t)words;
            continue;
        }
        const float b = fminf(exp2f((float)l), bound), cell = 2.0f * b / (float)H;
        for (int a = 0; a < 3; ++a) {
            const int hi = L.dim[a];
            wlo[a] = fminf(wlo[a], (float)L.lo[a] * cell - b - cell);
            whi[a] = fmaxf(whi[a], (float)(hi + 1) * cell - b + cell);
            L.dim[a] = hi - L.lo[a] + 1;
        }
        L.word_off = (int32_t)words;
        words += ((uint32_t)L.dim[0] * (uint32_t)L.dim[1] * (uint32_t)L.dim[2] + 31u) / 32u;
    }
    pk->total_words = (int32_t)words;
    pk->usable = (words > 0 && words <= max_words) ? 1 : 0;
    // an empty grid leaves (+max, -max): every ray is pruned, as it would find nothing
    for (int a = 0; a < 3; ++a) { pk->aabb[a] = wlo[a]; pk->aabb[3 + a] = whi[a]; }
}

// pass 3: one warp per output word (32 cells of a box row-major run) reads the Morton bitfield, ballot assembles the word
__global__ void __launch_bounds__(256)
occ_fill_kernel(const uint8_t* __restrict__ bitfield, uint32_t C, uint32_t H, OccPack* pk, uint32_t max_words) {
    if (!pk->usable) return;
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t H3 = H * H * H;
    for (uint32_t l = 0; l < C; ++l) {
        const OccLevel L = pk->lv[l];
        const uint32_t cells = (uint32_t)L.dim[0] * (uint32_t)L.dim[1] * (uint32_t)L.dim[2], words = (cells + 31u) / 32u;
        for (uint32_t w = warp; w < words; w += n_warps) {
            const uint32_t i = w * 32 + lane;
            bool bit = false;
            if (i < cells) {
                const uint32_t bx = i % (uint32_t)L.dim[0], r = i / (uint32_t)L.dim[0], by = r % (uint32_t)L.dim[1], bz = r / (uint32_t)L.dim[1];
                const uint32_t m = l * H3 + morton_encode((uint32_t)L.lo[0] + bx, (uint32_t)L.lo[1] + by, (uint32_t)L.lo[2] + bz);
                bit = (__ldg(bitfield + (m >> 3)) >> (m & 7u)) & 1u;
            }
            const uint32_t word = __ballot_sync(0xffffffffu, bit);
            if (lane == 0 && (uint32_t)L.word_off + w < max_words) pk->bits[L.word_off + w] = word;
        }
    }
}

}  // namespace
}  // namespace rn

using namespace rn;

extern "C" uint32_t rn_occupancy_pack_bytes(void) { return (uint32_t)(sizeof(OccPack) + RN_OCC_PACK_MAX_BYTES); }

// bitfield [C * H^3 / 8] (Morton order, LSB first) -> pack (rn_occupancy_pack_bytes() bytes, 16-byte aligned).  Header fields the
// host may read back: total_words (int32 at byte 4), usable (int32 at byte 8); the inflated world box (6 floats) starts at byte 16.
extern "C" int rn_occupancy_pack(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, void* pack, void* stream) {
    RN_REQUIRE(bitfield && pack, "null pointer");
    RN_REQUIRE(C >= 1 && C <= OCC_MAX_LEVELS && H >= 2 && H <= 1024 && (H * H * H) % 8 == 0, "bad cascade / grid size");
    RN_REQUIRE(((uintptr_t)pack & 15) == 0, "pack must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    OccPack* pk = (OccPack*)pack;
    const uint32_t max_words = RN_OCC_PACK_MAX_BYTES / 4;
    occ_init_kernel<<<1, 32, 0, st>>>(pk, C);
    occ_bounds_kernel<<<wave_grid((uint64_t)C * H * H * H / 8, 256 * 8, 4), 256, 0, st>>>(bitfield, C, H, pk);
    occ_layout_kernel<<<1, 32, 0, st>>>(pk, C, H, bound, max_words);
    occ_fill_kernel<<<RN_NUM_SMS * 2, 256, 0, st>>>(bitfield, C, H, pk, max_words);
    return finish_launch("rn_occupancy_pack");
}
