// gridencoder_d3.cu -- instantiates the grid-encoder kernels for input_dim = 3 (see gridencoder_impl.cuh).
#include "gridencoder_impl.cuh"
namespace rn { namespace grid {
RN_GRID_DEFINE_D(3)
} }
