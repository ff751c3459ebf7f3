// gridencoder_d4.cu -- instantiates the grid-encoder kernels for input_dim = 4 (see gridencoder_impl.cuh).
#include "gridencoder_impl.cuh"
namespace rn { namespace grid {
RN_GRID_DEFINE_D(4)
} }
