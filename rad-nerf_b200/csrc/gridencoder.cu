// gridencoder.cu -- C-ABI entry points of the grid encoder; kernels live in gridencoder_impl.cuh and are
// instantiated per input dimension in gridencoder_d{2,3,4,5}.cu.
#include "gridencoder_impl.cuh"

using namespace rn;
using namespace rn::grid;

template <typename A, typename F2, typename F3, typename F4, typename F5>
static int by_dim(uint32_t D, const A& a, F2 f2, F3 f3, F4 f4, F5 f5) {
    switch (D) {
        case 2: return f2(a);
        case 3: return f3(a);
        case 4: return f4(a);
        case 5: return f5(a);
        default: return RN_E_UNSUPPORTED;
    }
}

extern "C" int rn_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets,
                                      void* outputs, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                      uint32_t H, void* dy_dx, uint32_t gridtype, uint32_t align_corners,
                                      uint32_t interp, uint32_t dtype, uint32_t out_layout, void* stream) {
    RN_REQUIRE(L >= 1 && L <= (uint32_t)MAX_LEVELS, "num_levels must be in [1, 64]");
    RN_REQUIRE(dtype == RN_F32 || dtype == RN_F16, "dtype must be RN_F32 or RN_F16");
    RN_REQUIRE(out_layout <= 1 && gridtype <= 1 && interp <= 1, "bad enum argument");
    if (B == 0) return RN_OK;
    RN_REQUIRE(inputs && embeddings && offsets && outputs, "null pointer");
    const FwdArgs a{inputs, embeddings, offsets, outputs, dy_dx, B, C, L, S, H, gridtype, align_corners, interp,
                    dtype, out_layout, (cudaStream_t)stream};
    const int rc = by_dim(D, a, forward_d<2>, forward_d<3>, forward_d<4>, forward_d<5>);
    if (rc == RN_E_UNSUPPORTED)
        set_error("rn_grid_encode_forward: GridEncoding: D must be 2..5 and C must be 1, 2, 4, or 8 (got D=%u C=%u)", D, C);
    return rc;
}

extern "C" int rn_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings,
                                       const int32_t* offsets, void* grad_embeddings, uint32_t B, uint32_t D,
                                       uint32_t C, uint32_t L, float S, uint32_t H, const void* dy_dx,
                                       void* grad_inputs, uint32_t gridtype, uint32_t align_corners, uint32_t interp,
                                       uint32_t dtype, uint32_t grad_layout, uint32_t grad_emb_dtype, void* stream) {
    (void)embeddings;
    RN_REQUIRE(L >= 1 && L <= (uint32_t)MAX_LEVELS, "num_levels must be in [1, 64]");
    RN_REQUIRE(dtype <= 1 && grad_emb_dtype <= 1 && grad_layout <= 1 && gridtype <= 1 && interp <= 1, "bad enum argument");
    RN_REQUIRE(!(dtype == RN_F32 && grad_emb_dtype == RN_F16), "fp32 gradients need an fp32 table gradient");
    if (B == 0) return RN_OK;
    RN_REQUIRE(grad && inputs && offsets && grad_embeddings, "null pointer");
    RN_REQUIRE((dy_dx == nullptr) == (grad_inputs == nullptr), "dy_dx and grad_inputs must be given together");
    // the head's spatial encoder (3-D, 2 features, linear, fp32 target, no input gradient): restructured scatter, csrc/grid_bwd3.cu
    if (D == 3 && C == 2 && interp == 0 && !align_corners && grad_layout == RN_LAYOUT_BLC && grad_emb_dtype == RN_F32 && !dy_dx && L <= 32)
        return rn_grid_backward3(grad, inputs, offsets, (float*)grad_embeddings, B, L, S, H, gridtype, dtype, RN_BWD3_PRODUCTION, 0xffffffffu,
                                 RN_BWD3_AGG_LEVELS, 0, 0, stream);
    const BwdArgs a{grad, inputs, offsets, grad_embeddings, dy_dx, grad_inputs, B, C, L, S, H, gridtype, align_corners,
                    interp, dtype, grad_layout, grad_emb_dtype, (cudaStream_t)stream};
    const int rc = by_dim(D, a, backward_d<2>, backward_d<3>, backward_d<4>, backward_d<5>);
    if (rc == RN_E_UNSUPPORTED)
        set_error("rn_grid_encode_backward: GridEncoding: D must be 2..5 and C must be 1, 2, 4, or 8 (got D=%u C=%u)", D, C);
    return rc;
}

extern "C" int rn_grad_total_variation(const void* inputs, const void* embeddings, void* grad, const int32_t* offsets,
                                       float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                                       uint32_t H, uint32_t gridtype, uint32_t align_corners, uint32_t dtype,
                                       void* stream) {
    RN_REQUIRE(L >= 1 && L <= (uint32_t)MAX_LEVELS, "num_levels must be in [1, 64]");
    RN_REQUIRE(dtype <= 1 && gridtype <= 1, "bad enum argument");
    if (B == 0) return RN_OK;
    RN_REQUIRE(inputs && embeddings && grad && offsets, "null pointer");
    const TvArgs a{inputs, embeddings, grad, offsets, weight, B, C, L, S, H, gridtype, align_corners, dtype,
                   (cudaStream_t)stream};
    const int rc = by_dim(D, a, tv_d<2>, tv_d<3>, tv_d<4>, tv_d<5>);
    if (rc == RN_E_UNSUPPORTED)
        set_error("rn_grad_total_variation: GridEncoding: D must be 2..5 and C must be 1, 2, 4, or 8 (got D=%u C=%u)", D, C);
    return rc;
}

extern "C" int rn_grid_level_geometry(float S, uint32_t H, uint32_t L, float* scales_out, uint32_t* resolutions_out,
                                      void* stream) {
    if (L == 0) return RN_OK;
    level_geometry_kernel<<<div_up(L, 64u), 64, 0, (cudaStream_t)stream>>>(S, H, L, scales_out, resolutions_out);
    return finish_launch("rn_grid_level_geometry");
}
