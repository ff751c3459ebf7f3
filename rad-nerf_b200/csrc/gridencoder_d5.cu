// gridencoder_d5.cu -- instantiates the grid-encoder kernels for input_dim = 5 (see gridencoder_impl.cuh).
#include "gridencoder_impl.cuh"
namespace rn { namespace grid {
RN_GRID_DEFINE_D(5)
} }
