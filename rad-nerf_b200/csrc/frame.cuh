// frame.cuh -- shared declarations of the fused inference-frame renderer (frame_ctl.cu, head_eval.cu, torso_eval.cu).
//
// One frame of NeRFRenderer.run_cuda's inference branch (nerf/renderer.py:225-316) without a host round trip:
//
//   frame_init      near/far + per-ray state; ctl[0] = {n_alive = N, n_step = 1}
//   for it in 0 .. max_steps-1 (launched unconditionally; iterations after `done` exit at once):
//     march_compact   alive rays march <= n_step samples; samples are COMPACTED (block scan + one atomic per CTA)
//     head_eval       persistent tcgen05 kernel: 3-D encode -> ambient MLP -> tanh -> 2-D encode -> sigma MLP ->
//                     exp / SH -> colour MLP -> sigmoid, 128-sample tiles, activations never leave the SM
//     composite_compact  per-ray accumulate, termination, survivor compaction; the last CTA writes ctl[it+1]
//                     (n_alive, n_step = clamp(N / n_alive, 1, 8), step += n_step) -- the reference's host loop
//   torso_mask / torso_eval / finalize
//
// The reference's schedule (which couples all rays through n_alive) is reproduced exactly on the device, so every ray
// receives the same samples as under the reference's Python loop.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace rn {

struct FrameCtl {            // one per iteration, written by the previous iteration's last CTA
    uint32_t n_alive;        // rays in alive list `it & 1`
    uint32_t n_step;         // samples per ray this iteration
    uint32_t step;           // the reference's `step` before this iteration
    uint32_t done;           // nothing left to do
    uint32_t n_samples;      // samples emitted by march_compact(it)          (atomic)
    uint32_t next_alive;     // survivors appended by composite_compact(it)   (atomic)
    uint32_t blocks_done;    // ticket counter for "last CTA"                 (atomic)
    uint32_t total_samples;  // running total over the frame (statistics)
};
static_assert(sizeof(FrameCtl) == 32, "FrameCtl layout");

// State of the loop iteration in flight, at a FIXED address: the loop kernels read it directly (one load, no dependent
// index -> entry chain), the loop controller (last CTA of composite_compact) files it into the per-iteration history
// `ctl[it]` (diagnostics: frame_stats, rn_frame_head_timed) and replaces it with the next iteration's.
struct FrameCur {
    FrameCtl c;
    uint32_t it;             // index of this iteration
    uint32_t pad[7];
};
static_assert(sizeof(FrameCur) == 64, "FrameCur layout");

constexpr int FRAME_MAX_ITERS = 64;  // >= max_steps supported by the fused path
constexpr int EVAL_GROUPS = 3;       // 128-thread tile groups per CTA in the eval kernels
constexpr int EVAL_TILE = 128;

// byte sizes of the host-prepared fp16 weight blobs (interleaved UMMA operand layout, see umma.cuh)
constexpr uint32_t HEAD_BLOB_BYTES = (64 * 32 + 64 * 64 + 16 * 64 + 64 * 32 + 64 * 32 + 64 * 64 + 80 * 64 + 64 * 80 + 16 * 64) * 2;
constexpr uint32_t TORSO_BLOB_BYTES = (64 * 48 + 64 * 64 + 16 * 64 + 32 * 80 + 32 * 32 + 16 * 32) * 2;


struct MarchParams;

// workspace layout (all offsets 256-byte aligned); see carve() in frame_ctl.cu
struct FrameWorkspace {
    FrameCtl* ctl;        // [FRAME_MAX_ITERS + 1]
    FrameCur* cur;        // the iteration in flight (zeroed per frame): the same launches serve the unrolled sequence and the WHILE-node body
    uint32_t* tmisc;      // [8] zeroed by rn_frame_torso: 0 = n_torso (own block: the torso runs concurrently with the head loop)
    uint32_t* stats;      // [8] never reset by the library: 0 = loop iterations executed since the workspace was zeroed
    int32_t* alive[2];    // [N] each
    float* rays_t;        // [N]
    uint32_t* ray_cnt;    // [N]    samples the alive slot's ray got this iteration
    uint32_t* sample_idx; // [8][N] position of its k-th sample in the compacted (k-major per CTA) sample list
    float4* samples;      // [N]  xyz + ray id bits
    float2* deltas;       // [N]  (dt, t after)
    float4* evals;        // [N]  (sigma, r, g, b)
    int32_t* torso_pix;   // [N]
    float4* torso_out;    // [N]  (alpha, r, g, b), indexed by compact slot
};
size_t carve(FrameWorkspace& w, uint8_t* base, uint32_t N);

struct HeadEvalParams {
    const __half* table3; const int32_t* offs3; const int32_t* poffs3; float S3; uint32_t H3;
    const __half* table2; const int32_t* offs2; const int32_t* poffs2; float S2; uint32_t H2;
    const uint8_t* blob;      // HEAD_BLOB_BYTES, interleaved fp16
    const float* consts;      // [3][64] fp32: ambient-L1 bias, sigma-L1 bias, colour-L1 bias
    const float* rays_d;      // [N,3]
    const float4* samples;    // xyz + ray id
    float4* evals;            // (sigma, r, g, b)
    float bound, inv2bound;
    unsigned long long* prof;  // optional [8] cycle counters (diagnostics), null in production
};

struct TorsoEvalParams {
    const __half* table; const int32_t* offs; const int32_t* poffs; float S; uint32_t H;
    const uint8_t* blob;       // TORSO_BLOB_BYTES
    const float* consts;       // [64] deform-L1 bias, [32] torso-L1 bias
    const float* bg_coords;    // [N,2]
    const int32_t* pix;        // compact list of masked pixels
    const uint32_t* n_pix;     // device count
    float4* out;               // [n_pix] (alpha, r, g, b)
    float shrink;
};

// mirrors rn_conditioning_desc (include/radnerf_b200.h) field for field
struct AudioParams {
    const float* auds;   // [F, Cin, 16]
    uint32_t F, Cin, att, smooth, reserved;
    // all weights/biases below are fp16 COPIES of the fp32 parameters (the reference's autocast casts them per call)
    const __half* conv_w[4]; const __half* conv_b[4];        // AudioNet.encoder_conv.{0,2,4,6}
    const __half* fc_w[2]; const __half* fc_b[2];            // AudioNet.encoder_fc1.{0,2}
    const __half* att_w[5]; const __half* att_b[5];          // AudioAttNet.attentionConvNet.{0,2,4,6,8}
    const __half* att_fc_w; const __half* att_fc_b;          // AudioAttNet.attentionNet.0
    float* enc_a_state;  // [65]: smoothed code of the previous frame + validity flag at [64]
    float lambda;
    const __half* w_amb1; const __half* w_sig1; const __half* w_col1;   // [64,96] [64,65] [64,84] fp16 copies
    const float* eye;        // device [1] or null
    const float* ind_code;   // device [4] or null
    float* head_consts;      // [3][64]
    const __half* w_def1; const __half* w_tor1;  // [64,104] [32,136] fp16 copies, or null (no torso)
    const float* pose6;      // device [6]
    const float* ind_torso;  // device [8] or null
    float* torso_consts;     // [64 + 32]
    const float* pose44;     // device [16] or null: derive the 6-vector on the device (convert_poses, nerf/utils.py:230-237)
    float* pose6_out;        // device [6] or null
};

int launch_frame_init(const float* rays_o, const float* rays_d, const float* aabb, const float* occ_aabb /*nullable*/, uint32_t N,
                      float min_near, uint32_t max_steps, float* nears, float* fars, const FrameWorkspace& w, float* weights_sum, float* depth, float* image, cudaStream_t st);
int launch_march_compact(uint32_t N, const FrameWorkspace& w, const float* rays_o, const float* rays_d, const float* fars,
                         const MarchParams& p, const float* noises, const void* occ_pack /*nullable*/, uint32_t occ_words,
                         const float* occ_aabb /*nullable*/, cudaStream_t st);
// cond_handle: cudaGraphConditionalHandle of the WHILE node that repeats the iteration (0 = none); the loop controller sets it
int launch_composite_compact(uint32_t N, uint32_t max_steps, float T_thresh, const FrameWorkspace& w, float* weights_sum,
                             float* depth, float* image, unsigned long long cond_handle, cudaStream_t st);
int launch_torso_mask(const float* bg_coords, const float* grid, uint32_t G, float thresh, uint32_t N, const FrameWorkspace& w, cudaStream_t st);
int launch_torso_scatter(uint32_t N, const FrameWorkspace& w, float* torso_alpha, float* torso_color, cudaStream_t st);
int launch_finalize(uint32_t N, const float* weights_sum, float* depth, float* image, const float* nears, const float* fars,
                    const float* bg_color, float bg_scalar, const float* torso_alpha, const float* torso_color, float* torso_bg_out,
                    cudaStream_t st);
int launch_head_eval(const HeadEvalParams& p, const FrameCur* cur, uint32_t max_tiles, cudaStream_t st);
int launch_torso_eval(const TorsoEvalParams& p, uint32_t max_tiles, cudaStream_t st);
int launch_audio_frame(const AudioParams& p, cudaStream_t st);

}  // namespace rn
