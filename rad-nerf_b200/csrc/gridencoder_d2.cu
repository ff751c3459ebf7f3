// gridencoder_d2.cu -- instantiates the grid-encoder kernels for input_dim = 2 (see gridencoder_impl.cuh).
#include "gridencoder_impl.cuh"
namespace rn { namespace grid {
RN_GRID_DEFINE_D(2)
} }
