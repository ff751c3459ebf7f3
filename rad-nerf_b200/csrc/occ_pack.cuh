// occ_pack.cuh -- layout of the box-packed occupancy bits (occ_pack.cu builds it, the fused marcher of frame_ctl.cu reads it)
#pragma once
#include <stdint.h>

namespace rn {

constexpr uint32_t OCC_MAX_LEVELS = 16;
constexpr uint32_t RN_OCC_PACK_MAX_BYTES = 48 * 1024;   // what a marcher CTA stages next to its 24 KiB of sample staging

struct OccLevel {          // one cascade: cells [lo, lo + dim) hold every set bit of the level
    int32_t lo[3];
    int32_t dim[3];        // 0 = the level has no occupied cell
    int32_t word_off;      // first 32-bit word of the level's bits inside OccPack::bits
    int32_t pad;
};

struct OccPack {
    int32_t n_levels;
    int32_t total_words;   // 32-bit words of packed bits over all levels
    int32_t usable;        // 1: the bits fit RN_OCC_PACK_MAX_BYTES and were written; 0: use the Morton bitfield in global memory
    int32_t pad;
    float aabb[6];         // world-space box of every occupied cell, inflated by one cell (rn_frame_head_desc.occ_aabb semantics)
    float pad2[2];
    OccLevel lv[OCC_MAX_LEVELS];
    uint32_t bits[1];      // total_words words, bit (bz * dim[1] + by) * dim[0] + bx of a level, LSB first
};

}  // namespace rn
