#include <algorithm>
// frame.cu -- C-ABI entry points of the fused inference frame (see frame.cuh for the pipeline).
#include "frame.cuh"
#include "march.cuh"
#include <cstring>
#include <vector>

using namespace rn;

static void* g_head_prof = nullptr;
// diagnostics: device buffer of 8 uint64 cycle counters filled by head_eval (enc3, enc2, mma, epilogue, tile, groups); NULL disables
extern "C" void rn_debug_set_head_prof(void* p) { g_head_prof = p; }

namespace rn { void set_audio_prof(void* p); }
extern "C" void rn_debug_set_audio_prof(void* p) { rn::set_audio_prof(p); }

extern "C" uint64_t rn_frame_workspace_bytes(uint32_t N) {
    FrameWorkspace w;
    return (uint64_t)carve(w, nullptr, N);
}
extern "C" uint32_t rn_head_blob_bytes(void) { return HEAD_BLOB_BYTES; }
extern "C" uint32_t rn_torso_blob_bytes(void) { return TORSO_BLOB_BYTES; }

extern "C" int rn_frame_conditioning(const rn_conditioning_desc* d, void* stream) {
    RN_REQUIRE(d, "null descriptor");
    RN_REQUIRE(d->head_consts && d->w_amb1 && d->w_sig1 && d->w_col1, "null pointer");
    RN_REQUIRE(!d->auds || (d->F >= 1 && d->F <= 8 && d->Cin >= 1 && d->Cin <= 44), "auds must be [F<=8, Cin<=44, 16]");
    RN_REQUIRE(!d->auds || d->att == 0 || d->F == 8, "attention needs the 8-frame window");
    RN_REQUIRE(!d->smooth || d->enc_a_state, "lip smoothing needs a state buffer");
    RN_REQUIRE(!d->w_def1 || (d->w_tor1 && (d->pose6 || d->pose44) && d->torso_consts), "incomplete torso conditioning");
    static_assert(sizeof(AudioParams) == sizeof(rn_conditioning_desc), "descriptor mirrors AudioParams");
    AudioParams p;
    memcpy(&p, d, sizeof(p));
    return launch_audio_frame(p, (cudaStream_t)stream);
}

#define RN_CUDA(call)                                                                    \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            rn::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return (int)e_;                                                              \
        }                                                                                \
    } while (0)

static uint32_t g_debug_iters = 0;
static bool g_use_while_node = true;
extern "C" void rn_debug_set_while_node(int on) { g_use_while_node = on != 0; }
extern "C" void rn_debug_set_max_iters(uint32_t n) { g_debug_iters = n; }

static int frame_head_impl(const rn_frame_head_desc* d, cudaStream_t st, cudaEvent_t* ev /* nullable: 3*max_steps+1 events */) {
    RN_REQUIRE(d, "null descriptor");
    if (d->N == 0) return RN_OK;
    RN_REQUIRE(d->rays_o && d->rays_d && d->aabb && d->bitfield && d->weights_sum && d->depth && d->image && d->nears && d->fars,
               "null pointer");
    RN_REQUIRE(d->workspace && d->workspace_bytes >= rn_frame_workspace_bytes(d->N), "workspace too small");
    RN_REQUIRE(d->grid3d.packed_offsets && d->grid2d.packed_offsets, "grid tables need packed_offsets");
    RN_REQUIRE(d->grid3d.table_f16 && d->grid3d.offsets && d->grid2d.table_f16 && d->grid2d.offsets && d->head_blob && d->head_consts,
               "null network pointer");
    RN_REQUIRE(d->max_steps >= 1 && d->max_steps <= (uint32_t)FRAME_MAX_ITERS, "max_steps must be in [1, 64] for the fused frame");
    RN_REQUIRE(d->cascade >= 1 && d->cascade <= 16 && d->grid_size >= 1, "bad cascade/grid_size");
    RN_REQUIRE(((uintptr_t)d->head_blob & 15) == 0, "head_blob must be 16-byte aligned");
    RN_REQUIRE(!d->occ_pack || (((uintptr_t)d->occ_pack & 15) == 0 && d->occ_words * 4u <= rn_occupancy_pack_bytes()), "occ_pack misaligned or occ_words too large");
    FrameWorkspace w;
    carve(w, (uint8_t*)d->workspace, d->N);
    int rc = launch_frame_init(d->rays_o, d->rays_d, d->aabb, d->occ_aabb, d->N, d->min_near, d->max_steps, d->nears, d->fars, w, d->weights_sum,
                               d->depth, d->image, st);
    if (rc) return rc;
    const MarchParams mp = make_march_params(d->bound, d->dt_gamma, d->max_steps, d->cascade, d->grid_size, d->bitfield);
    HeadEvalParams hp;
    hp.table3 = (const __half*)d->grid3d.table_f16; hp.offs3 = d->grid3d.offsets; hp.poffs3 = d->grid3d.packed_offsets; hp.S3 = d->grid3d.S; hp.H3 = d->grid3d.H;
    hp.table2 = (const __half*)d->grid2d.table_f16; hp.offs2 = d->grid2d.offsets; hp.poffs2 = d->grid2d.packed_offsets; hp.S2 = d->grid2d.S; hp.H2 = d->grid2d.H;
    hp.blob = (const uint8_t*)d->head_blob; hp.consts = d->head_consts; hp.rays_d = d->rays_d;
    hp.samples = w.samples; hp.evals = w.evals; hp.bound = d->bound; hp.inv2bound = 1.0f / (2.0f * d->bound);
    hp.prof = (unsigned long long*)g_head_prof;
    const uint32_t max_tiles = (d->N + EVAL_TILE - 1) / EVAL_TILE;  // n_alive * n_step <= N in every iteration
    // n_step >= 1, so the reference's loop runs at most max_steps iterations.  All iterations are the same three launches:
    // the kernels read the iteration index from the workspace and the loop controller (last CTA of composite) advances it.
    auto iteration = [&](cudaStream_t s, uint32_t ev_it, bool first, unsigned long long cond) -> int {
        int r;
        if ((r = launch_march_compact(d->N, w, d->rays_o, d->rays_d, d->fars, mp, d->noises, d->occ_pack, d->occ_words, d->occ_aabb, s))) return r;
        if (ev) cudaEventRecord(ev[3 * ev_it + 1], s);
        if (first && d->consts_ready_event) cudaStreamWaitEvent(s, (cudaEvent_t)d->consts_ready_event, 0);
        if ((r = launch_head_eval(hp, w.cur, max_tiles, s))) return r;
        if (ev) cudaEventRecord(ev[3 * ev_it + 2], s);
        if ((r = launch_composite_compact(d->N, d->max_steps, d->T_thresh, w, d->weights_sum, d->depth, d->image, cond, s))) return r;
        if (ev) cudaEventRecord(ev[3 * ev_it + 3], s);
        return RN_OK;
    };
    if (ev) cudaEventRecord(ev[0], st);

    // ---- stream capture: `capture_unroll` plain iterations, then ONE conditional WHILE node whose body is an iteration; the
    //      loop controller sets the node's condition from the device.  Replaces max_steps unrolled iterations, most of which
    //      would be launches that find `done` and exit (measured: 33 such launches, 44 us of a 530 us frame).
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaGraph_t graph = nullptr;
    if (!ev && g_use_while_node && d->max_steps > 1 && cudaStreamGetCaptureInfo(st, &cs, nullptr, &graph, nullptr, nullptr) == cudaSuccess &&
        cs == cudaStreamCaptureStatusActive && graph) {
        cudaGraphConditionalHandle handle;
        RN_CUDA(cudaGraphConditionalHandleCreate(&handle, graph, 0 /* no body run unless a controller asks for it */, cudaGraphCondAssignDefault));
        const uint32_t unroll = std::min(std::max(d->capture_unroll, 1u), d->max_steps);
        for (uint32_t it = 0; it < unroll; ++it)
            if ((rc = iteration(st, it, it == 0, (unsigned long long)handle))) return rc;
        if (unroll == d->max_steps) return RN_OK;
        const cudaGraphNode_t* deps = nullptr;
        size_t n_deps = 0;
        RN_CUDA(cudaStreamGetCaptureInfo(st, &cs, nullptr, &graph, &deps, &n_deps));
        cudaGraphNodeParams np = {};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        RN_CUDA(cudaGraphAddNode(&node, graph, deps, n_deps, &np));
        cudaGraph_t body = np.conditional.phGraph_out[0];
        static cudaStream_t body_stream = nullptr;   // capture-only helper stream
        if (!body_stream) RN_CUDA(cudaStreamCreateWithFlags(&body_stream, cudaStreamNonBlocking));
        RN_CUDA(cudaStreamBeginCaptureToGraph(body_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
        rc = iteration(body_stream, 0, false, (unsigned long long)handle);
        cudaGraph_t out = nullptr;
        RN_CUDA(cudaStreamEndCapture(body_stream, &out));
        if (rc) return rc;
        RN_CUDA(cudaStreamUpdateCaptureDependencies(st, &node, 1, cudaStreamSetCaptureDependencies));
        return RN_OK;
    }
    const uint32_t n_iters = g_debug_iters ? std::min(g_debug_iters, d->max_steps) : d->max_steps;
    for (uint32_t it = 0; it < n_iters; ++it)
        if ((rc = iteration(st, it, it == 0, 0ull))) return rc;
    return RN_OK;
}

extern "C" int rn_frame_head(const rn_frame_head_desc* d, void* stream) { return frame_head_impl(d, (cudaStream_t)stream, nullptr); }

// diagnostics / bench: the same launch sequence bracketed by CUDA events on the launching stream.  SYNCHRONISES the stream.
// ms [3 * max_steps] = (march, eval, composite) per iteration; n_samples [max_steps] = samples evaluated per iteration.
extern "C" int rn_frame_head_timed(const rn_frame_head_desc* d, void* stream, float* ms, uint32_t* n_samples) {
    RN_REQUIRE(d && ms && n_samples, "null pointer");
    const uint32_t n_ev = 3 * d->max_steps + 1;
    std::vector<cudaEvent_t> ev(n_ev);
    for (auto& e : ev) cudaEventCreate(&e);
    int rc = frame_head_impl(d, (cudaStream_t)stream, ev.data());
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (rc == RN_OK && e != cudaSuccess) { set_error("rn_frame_head_timed: %s", cudaGetErrorString(e)); rc = (int)e; }
    if (rc == RN_OK) {
        for (uint32_t i = 0; i + 1 < n_ev; ++i) cudaEventElapsedTime(ms + i, ev[i], ev[i + 1]);
        FrameWorkspace w;
        carve(w, (uint8_t*)d->workspace, d->N);
        std::vector<FrameCtl> ctl(d->max_steps);
        cudaMemcpy(ctl.data(), w.ctl, sizeof(FrameCtl) * d->max_steps, cudaMemcpyDeviceToHost);
        for (uint32_t it = 0; it < d->max_steps; ++it) n_samples[it] = ctl[it].done ? 0u : ctl[it].n_samples;
    }
    for (auto& e2 : ev) cudaEventDestroy(e2);
    return rc;
}

extern "C" int rn_frame_torso(const rn_frame_torso_desc* d, void* stream) {
    RN_REQUIRE(d, "null descriptor");
    if (d->N == 0) return RN_OK;
    RN_REQUIRE(d->bg_coords && d->density_grid_torso && d->torso_alpha && d->torso_color && d->torso_blob && d->torso_consts, "null pointer");
    RN_REQUIRE(d->workspace && d->workspace_bytes >= rn_frame_workspace_bytes(d->N), "workspace too small");
    RN_REQUIRE(d->grid2d.table_f16 && d->grid2d.offsets && d->grid2d.packed_offsets, "null network pointer");
    RN_REQUIRE(((uintptr_t)d->torso_blob & 15) == 0, "torso_blob must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    FrameWorkspace w;
    carve(w, (uint8_t*)d->workspace, d->N);
    cudaMemsetAsync(w.tmisc, 0, 32, st);
    int rc = launch_torso_mask(d->bg_coords, d->density_grid_torso, d->grid_size, d->thresh, d->N, w, st);
    if (rc) return rc;
    TorsoEvalParams tp;
    tp.table = (const __half*)d->grid2d.table_f16; tp.offs = d->grid2d.offsets; tp.poffs = d->grid2d.packed_offsets; tp.S = d->grid2d.S; tp.H = d->grid2d.H;
    tp.blob = (const uint8_t*)d->torso_blob; tp.consts = d->torso_consts; tp.bg_coords = d->bg_coords; tp.pix = w.torso_pix;
    tp.n_pix = w.tmisc; tp.out = w.torso_out; tp.shrink = d->shrink;
    if ((rc = launch_torso_eval(tp, (d->N + EVAL_TILE - 1) / EVAL_TILE, st))) return rc;
    return launch_torso_scatter(d->N, w, d->torso_alpha, d->torso_color, st);
}

extern "C" int rn_frame_finalize(uint32_t N, const float* weights_sum, float* depth, float* image, const float* nears,
                                 const float* fars, const float* bg_color, float bg_scalar, const float* torso_alpha,
                                 const float* torso_color, float* torso_bg_out, void* stream) {
    if (N == 0) return RN_OK;
    RN_REQUIRE(weights_sum && depth && image && nears && fars, "null pointer");
    RN_REQUIRE(!torso_alpha || torso_color, "torso_color missing");
    return launch_finalize(N, weights_sum, depth, image, nears, fars, bg_color, bg_scalar, torso_alpha, torso_color, torso_bg_out,
                           (cudaStream_t)stream);
}
