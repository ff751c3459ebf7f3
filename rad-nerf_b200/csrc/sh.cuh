// sh.cuh -- generated real-SH basis (sh_basis.inc, tools/gen_sh.py) wrapped as sh_eval<DEG, GRAD>.
#pragma once
namespace rn {
#include "sh_basis.inc"

template <int DEG, bool GRAD>
__device__ __forceinline__ void sh_eval(float x, float y, float z, float* Y, float* gx, float* gy, float* gz) {
    sh_band0<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 1) sh_band1<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 2) sh_band2<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 3) sh_band3<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 4) sh_band4<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 5) sh_band5<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 6) sh_band6<GRAD>(x, y, z, Y, gx, gy, gz);
    if constexpr (DEG > 7) sh_band7<GRAD>(x, y, z, Y, gx, gy, gz);
}
}  // namespace rn
