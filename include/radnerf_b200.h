/*
 * radnerf_b200.h -- C ABI of libradnerf_b200.so: RAD-NeRF's per-ray rendering hot path on B200 (sm_100a).
 *
 * Drop-in boundary.  Every entry point below replaces one function of the reference's four pybind11
 * torch extensions (the `_backend.<fn>` calls made by raymarching/raymarching.py, gridencoder/grid.py,
 * freqencoder/freq.py, shencoder/sphere_harmonics.py).  The reference passes at::Tensor; here every
 * tensor is a raw DEVICE pointer plus extents, and the CUDA stream is explicit (the reference launches
 * on the legacy default stream).  No torch types cross this boundary.
 *
 * Conventions
 *   - return value: 0 = ok, <0 = argument error (RN_E_*), >0 = cudaError_t of the launch.
 *     rn_last_error_string() returns a thread-local, human-readable description of the last failure.
 *   - ownership: the CALLER allocates every output (as the reference's Python wrappers do); kernels
 *     never allocate.  In-place semantics are the reference's (composite_rays, packbits, counter).
 *   - all pointers are device pointers on the current device; `stream` is a cudaStream_t (0 = legacy).
 *   - dtype: RN_F32 (0) or RN_F16 (1) selects the table/feature arithmetic of the grid encoder
 *     (reference: AT_DISPATCH_FLOATING_TYPES_AND_HALF on embeddings, gridencoder.cu:466).
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef RADNERF_B200_H
#define RADNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RN_F32 0
#define RN_F16 1

#define RN_OK 0
#define RN_E_BADARG (-1)     /* null pointer / zero extent where not allowed */
#define RN_E_UNSUPPORTED (-2) /* D, C, degree ... outside the supported template set */

/* output layout of the grid encoder */
#define RN_LAYOUT_LBC 0 /* [L, B, C]   -- what the reference kernel writes (gridencoder.cu:108)           */
#define RN_LAYOUT_BLC 1 /* [B, L*C]    -- what GridEncoder.forward returns after its permute (grid.py:57) */

const char* rn_last_error_string(void);
/* library/ABI version, bumped on any signature change */
int rn_abi_version(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t rn_launch_count(void);
/* account for the kernels of a captured CUDA graph that a replay launches again (host bookkeeping only) */
void rn_note_graph_replay(uint64_t kernels_in_graph);

/* ------------------------------------------------------------------ gridencoder ------------------- */

/* replaces grid_encode_forward (gridencoder/src/gridencoder.h:12, gridencoder.cu:447-470; kernel_grid :87-244).
 * inputs [B,D] f32 in [0,1]; embeddings [rows,C] dtype; offsets [L+1] i32; outputs dtype, layout as given;
 * dy_dx (nullable) [B, L*D*C] dtype.  D in {2,3,4,5}, C in {1,2,4,8}.  gridtype 0 = hash, 1 = tiled;
 * interp 0 = linear, 1 = smoothstep. */
int rn_grid_encode_forward(const float* inputs, const void* embeddings, const int32_t* offsets, void* outputs,
                           uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                           void* dy_dx, uint32_t gridtype, uint32_t align_corners, uint32_t interp,
                           uint32_t dtype, uint32_t out_layout, void* stream);

/* replaces grid_encode_backward (gridencoder.h:13, gridencoder.cu:472-502; kernel_grid_backward :247-339,
 * kernel_input_backward :342-368).
 * grad: dtype, layout `grad_layout` (RN_LAYOUT_LBC as the reference passes it after grid.py:75, or
 * RN_LAYOUT_BLC straight from autograd).  grad_embeddings [rows,C]: accumulated INTO (caller pre-zeroes,
 * grid.py:77), element type `grad_emb_dtype` (RN_F32 accumulates in fp32 -- the product default -- or RN_F16
 * for the reference's half2 atomics).  dy_dx/grad_inputs nullable together; grad_inputs [B,D] dtype. */
int rn_grid_encode_backward(const void* grad, const float* inputs, const void* embeddings, const int32_t* offsets,
                            void* grad_embeddings, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S,
                            uint32_t H, const void* dy_dx, void* grad_inputs, uint32_t gridtype,
                            uint32_t align_corners, uint32_t interp, uint32_t dtype, uint32_t grad_layout,
                            uint32_t grad_emb_dtype, void* stream);

/* replaces grad_total_variation (gridencoder.h:15, gridencoder.cu:505-644).  inputs [B,D] dtype in [0,1];
 * grad [rows,C] dtype accumulated into. */
int rn_grad_total_variation(const void* inputs, const void* embeddings, void* grad, const int32_t* offsets,
                            float weight, uint32_t B, uint32_t D, uint32_t C, uint32_t L, float S, uint32_t H,
                            uint32_t gridtype, uint32_t align_corners, uint32_t dtype, void* stream);

/* helper (no reference counterpart): per-level `scale = exp2f(l*S)*H - 1` and `resolution = ceil(scale)+1`
 * exactly as the device computes them (gridencoder.cu:138-139); lets a CPU checker pin its level geometry. */
int rn_grid_level_geometry(float S, uint32_t H, uint32_t L, float* scales_out, uint32_t* resolutions_out,
                           void* stream);

/* ------------------------------------------------------------------ raymarching: utilities -------- */

/* replaces near_far_from_aabb (raymarching.h:7, raymarching.cu:91-156). rays_o/d [N,3], aabb [6], nears/fars [N] */
int rn_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                          float min_near, float* nears, float* fars, void* stream);
/* replaces sph_from_ray (raymarching.h:8, raymarching.cu:162-209). coords [N,2] */
int rn_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                    void* stream);
/* replaces morton3D (raymarching.h:9, raymarching.cu:214-232). coords [N,3] i32 -> indices [N] i32 */
int rn_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, void* stream);
/* replaces morton3D_invert (raymarching.h:10, raymarching.cu:237-260). indices [N] -> coords [N,3] */
int rn_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, void* stream);
/* replaces packbits (raymarching.h:11, raymarching.cu:267-300). grid [8N] f32 -> bitfield [N] u8, LSB first */
int rn_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, void* stream);
/* as rn_packbits with the threshold min(density_thresh, *mean_density), mean_density a DEVICE scalar (NULL: plain density_thresh):
 * the `min(self.mean_density, self.density_thresh)` of update_extra_state (nerf/renderer.py:471) without reading the mean back */
int rn_packbits_min(const float* grid, uint32_t N, float density_thresh, const float* mean_density, uint8_t* bitfield, void* stream);
/* The grid merge of update_extra_state (nerf/renderer.py:463-468) in one pass: where grid >= 0 and fresh >= 0,
 * grid = max(grid * decay, fresh) (in place); *mean_out = mean(clamp(grid, 0)) over all n cells (deterministic two-stage sum in
 * double).  workspace: rn_occupancy_merge_workspace_bytes() bytes, zeroed once before the first call. */
int rn_occupancy_merge(float* grid, const float* fresh, uint32_t n, float decay, void* workspace, float* mean_out, void* stream);
uint32_t rn_occupancy_merge_workspace_bytes(void);
/* replaces morton3D_dilation (raymarching.h:12, raymarching.cu:304-341). grid [C,H^3] f32 Morton-indexed */
int rn_morton3D_dilation(const float* grid, uint32_t C, uint32_t H, float* grid_dilation, void* stream);

/* ------------------------------------------------------------------ raymarching: training --------- */

/* replaces march_rays_train (raymarching.h:14, raymarching.cu:352-528).
 * xyzs/dirs [M,3], deltas [M,2] (caller pre-zeroes, raymarching.py:231-233), rays [N,3] i32 (id, offset, count),
 * counter [2] i32 (points, rays) accumulated atomically, noises [N]. */
int rn_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                        uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                        const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays,
                        int32_t* counter, const float* noises, void* stream);
/* as rn_march_rays_train, with the sample budget on the DEVICE: buffers hold M samples, a ray is dropped when it would end past
 * min(M, *budget).  Lets a training step captured in a CUDA graph keep fixed buffer shapes while the reference's running
 * estimate (`mean_count`, raymarching.py:213-229, renderer.py:489-493) changes between replays.  budget == NULL: plain M. */
int rn_march_rays_train_budget(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                               uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const int32_t* budget,
                               const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                               int32_t* rays, int32_t* counter, const float* noises, void* stream);
/* replaces march_rays_train_backward (raymarching.h:15, raymarching.cu:535-593); grads accumulated into */
int rn_march_rays_train_backward(const float* grad_xyzs, const float* grad_dirs, const int32_t* rays,
                                 const float* deltas, uint32_t N, uint32_t M, float* grad_rays_o,
                                 float* grad_rays_d, void* stream);
/* replaces composite_rays_train_forward (raymarching.h:16, raymarching.cu:603-698) */
int rn_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* ambient,
                                    const float* deltas, const int32_t* rays, uint32_t M, uint32_t N,
                                    float T_thresh, float* weights_sum, float* ambient_sum, float* depth,
                                    float* image, void* stream);
/* replaces composite_rays_train_backward (raymarching.h:17, raymarching.cu:711-820) */
int rn_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_ambient_sum,
                                     const float* grad_image, const float* sigmas, const float* rgbs,
                                     const float* ambient, const float* deltas, const int32_t* rays,
                                     const float* weights_sum, const float* ambient_sum, const float* image,
                                     uint32_t M, uint32_t N, float T_thresh, float* grad_sigmas,
                                     float* grad_rgbs, float* grad_ambient, void* stream);

/* ------------------------------------------------------------------ raymarching: inference -------- */

/* replaces march_rays (raymarching.h:19, raymarching.cu:827-939). Slot layout n*n_step+k, caller pre-zeroes. */
int rn_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                  const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                  uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                  float* xyzs, float* dirs, float* deltas, const float* noises, void* stream);
/* replaces composite_rays (raymarching.h:20, raymarching.cu:942-1038); in-place on weights_sum/depth/image,
 * rays_alive[n] = -1 marks a terminated ray, rays_t updated otherwise. */
int rn_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, int32_t* rays_alive, float* rays_t,
                      const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                      float* depth, float* image, void* stream);

/* ------------------------------------------------------------------ freqencoder / shencoder ------- */

/* replaces freq_encode_forward (freqencoder.h:7, freqencoder.cu:30-58,96-110). outputs [B,C], C = D + 2*D*deg */
int rn_freq_encode_forward(const float* inputs, uint32_t B, uint32_t D, uint32_t deg, uint32_t C, float* outputs,
                           void* stream);
/* replaces freq_encode_backward (freqencoder.h:9, freqencoder.cu:63-94,113-128). grad_inputs [B,D] written */
int rn_freq_encode_backward(const float* grad, const float* outputs, uint32_t B, uint32_t D, uint32_t deg,
                            uint32_t C, float* grad_inputs, void* stream);
/* replaces sh_encode_forward (shencoder.h:8, shencoder.cu:27-356,400-417). inputs [B,3] f32, outputs [B,deg^2],
 * dy_dx nullable [B,3*deg^2]; degree 1..8 */
int rn_sh_encode_forward(const float* inputs, float* outputs, uint32_t B, uint32_t D, uint32_t degree,
                         float* dy_dx, void* stream);
/* replaces sh_encode_backward (shencoder.h:9, shencoder.cu:358-383,419-439). grad_inputs [B,3] accumulated into */
int rn_sh_encode_backward(const float* grad, const float* inputs, uint32_t B, uint32_t D, uint32_t degree,
                          const float* dy_dx, float* grad_inputs, void* stream);

/* ------------------------------------------------------------------ callers absorbed into the path ---- */

/* replaces the per-frame torch-op chain of get_rays for full frames (nerf/utils.py:248-333):
 * pixel p = (row j, col i) -> direction ((i+0.5-cx)/fx, (j+0.5-cy)/fy, 1) normalised and rotated by pose[:3,:3],
 * origin pose[:3,3].  pose: device float[16] row-major cam2world.  pixel_ids (nullable) int32 [n]: the pixels to
 * generate (row-major ids); NULL = all H*W pixels in order.  rays_o / rays_d [n,3]. */
int rn_get_rays(const float* pose, float fx, float fy, float cx, float cy, uint32_t H, uint32_t W,
                const int32_t* pixel_ids, uint32_t n, float* rays_o, float* rays_d, void* stream);

/* Output stage: fp32 image values in [0,1] -> uint8 on the device, exactly the host expression of the
 * reference's video writer `(pred * 255).astype(np.uint8)` (nerf/utils.py:952-960).  n_values % 16 == 0. */
int rn_image_to_uint8(const float* image, uint8_t* out, uint64_t n_values, void* stream);

/* Multi-GPU frame assembly without a collective library (replaces ncclAllGather + un-permute for the ray-sharded frame,
 * SURVEY 8(e) "Inference frame"): this rank's finished image rows are stored directly into every rank's full-frame buffer
 * through peer mappings (NVLink / NVSwitch).  local [n_local,3] fp32; ids [n_local] pixel index of each row in the full frame,
 * organised in runs of run_pixels consecutive pixels (run_pixels*3 % 4 == 0, run starts 16-byte aligned in the frame);
 * peers: device array of `world` base addresses of the [H*W,3] fp32 frames (e.g. torch symmetric memory).  The caller
 * orders the stores against the readers with a cross-rank barrier. */
int rn_scatter_rows_to_peers(const float* local, const int32_t* ids, uint32_t n_local, uint32_t run_pixels,
                             const uint64_t* peers, uint32_t world, void* stream);
/* Gather-to-root with flags instead of a barrier (the frames are delivered by ONE rank, so only its buffer is assembled):
 * a non-root rank waits until the root has consumed frame seq-1 of frame buffer `slot` (ctrl word 2*slot+1 in ITS OWN control
 * block, written by the root), stores its rows into the root's frame buffer and adds 1 per CTA (rn_scatter_signal_ctas of them) to
 * the root's arrival counter (ctrl word 2*slot; release, system scope).  The root stores its own rows with the same call (no flags).
 * frame_peers / ctrl_peers: DEVICE arrays of `world` base addresses of the [H*W,3] fp32 frame buffers / of the uint64 control blocks
 * (>= 2 * slots words, zero-initialised before the first frame).  seq = 1, 2, ... per slot, the same on every rank. */
int rn_scatter_rows_to_root(const float* local, const int32_t* ids, uint32_t n_local, uint32_t run_pixels, const uint64_t* frame_peers,
                            const uint64_t* ctrl_peers, uint32_t world, uint32_t rank, uint32_t root, uint32_t slot, uint64_t seq,
                            void* stream);
uint32_t rn_scatter_signal_ctas(uint32_t n_local, uint32_t run_pixels);
/* Root side: waits until seq * (world-1) * scatter_ctas arrivals have been counted for `slot`, copies the assembled frame (n_values
 * fp32) to dst -- as fp32, or as uint8 `(v * 255)` truncated when to_uint8 (dst NULL: nothing is copied) -- and then writes seq into
 * every other rank's consumed word.  ctrl_peers NULL: plain copy / conversion without any exchange (one GPU).  ticket: a zeroed
 * device uint32 owned by the caller (one per slot). */
int rn_stage_frame_at_root(const float* frame, void* dst, uint64_t n_values, uint32_t to_uint8, const uint64_t* ctrl_peers, uint32_t world,
                           uint32_t root, uint32_t slot, uint64_t seq, uint32_t scatter_ctas, uint32_t* ticket, void* stream);

/* ------------------------------------------------------------------ fused inference frame ------------ */
/* One call per stage of NeRFRenderer.run_cuda's inference branch (nerf/renderer.py:158-316) with no host round trip:
 * the reference's Python `while step < max_steps` loop, its per-iteration march_rays / NeRFNetwork.forward /
 * composite_rays calls and the `rays_alive[rays_alive >= 0]` host sync run as a fixed launch sequence driven by
 * device-resident loop state; NeRFNetwork.forward / forward_torso run as fused tcgen05 kernels (csrc/head_eval.cu,
 * csrc/torso_eval.cu).  Architecture is fixed to the one nerf/network.py builds: 16-level, 2-feature tiled grids with
 * linear interpolation, hidden width 64 (torso 32), SH degree 4.  Tables are fp16 (the reference's autocast path). */

typedef struct rn_grid_table {
    const void* table_f16;   /* [rows', 4] fp16: packed copy, row = [features of the cell (2) | features of its +x neighbour (2)] */
    const int32_t* offsets;  /* [17] the encoder's offsets (gridencoder/grid.py:118-131): level sizes and wrap rules */
    float S;                 /* log2(per_level_scale) */
    uint32_t H;              /* base resolution */
    const int32_t* packed_offsets; /* [16] first row of each level inside table_f16.  The fused kernels index a wrapped
                                      (size-capped) level as ((i & (size-1)) | first_row), one logic op instead of an
                                      and + 64-bit add per corner, so such a level must start at a multiple of its
                                      power-of-two size: radnerf_b200/frame.py re-packs its fp16 copy accordingly.
                                      A table that violates this traps the kernel.  Never NULL. */
} rn_grid_table;

/* per-frame conditioning: AudioNet + AudioAttNet + lip smoothing + hoisted first-layer bias vectors
 * (nerf/network.py:10-67,170-185; nerf/renderer.py:187-204).  Weight/bias pointers are fp16 COPIES of the fp32
 * parameters (what the reference's autocast produces per call; f32->f16 conversion inside the kernel would run on the
 * quarter-rate conversion pipe); auds, eye, codes and pose stay fp32. */
typedef struct rn_conditioning_desc {
    const float* auds;           /* [F, Cin, 16] or NULL (no audio) */
    uint32_t F, Cin, att, smooth;
    uint32_t reserved;           /* phase: 0 = whole conditioning; 1 = audio features only (AudioNet + attention; the raw code is parked in
                                    head_consts[0..63]); 2 = smoothing + hoisted terms only (reads it back).  1 then 2 == 0; the split lets
                                    frames in flight evaluate their audio nets concurrently and serialises only the EMA tail */
    const void* conv_w[4]; const void* conv_b[4];      /* fp16 */
    const void* fc_w[2]; const void* fc_b[2];
    const void* att_w[5]; const void* att_b[5];
    const void* att_fc_w; const void* att_fc_b;
    float* enc_a_state;          /* [65] in/out: smoothed code [0..63], [64] != 0 once it holds a previous frame (device-side flag,
                                    so a captured CUDA graph serves first and later frames alike) */
    float lambda;
    const void* w_amb1; const void* w_sig1; const void* w_col1;   /* fp16 [64,96] [64,65] [64,84] */
    const float* eye; const float* ind_code;
    float* head_consts;          /* out [3*64] */
    const void* w_def1; const void* w_tor1;                        /* fp16 [64,104] [32,136] or NULL */
    const float* pose6; const float* ind_torso;
    float* torso_consts;         /* out [64+32] */
    const float* pose44;         /* optional device [16] cam2world: when given, the 6-vector (XYZ Euler angles, translation) of the
                                    reference's convert_poses (nerf/utils.py:230-237) is computed ON THE DEVICE from it and `pose6`
                                    is ignored; NULL = use pose6 */
    float* pose6_out;            /* optional device [6]: receives the 6-vector the kernel used (parity checks) */
} rn_conditioning_desc;

typedef struct rn_frame_head_desc {
    uint32_t N, max_steps, cascade, grid_size;
    float bound, min_near, dt_gamma, T_thresh;
    const float* rays_o; const float* rays_d; const float* aabb; const uint8_t* bitfield;
    const float* noises;         /* [N] or NULL: perturbation of the first iteration */
    float* weights_sum; float* depth; float* image; float* nears; float* fars;   /* outputs [N], [N], [N,3], [N], [N] */
    void* workspace; uint64_t workspace_bytes;
    rn_grid_table grid3d, grid2d;
    const void* head_blob;       /* rn_head_blob_bytes() of interleaved fp16 weights (see radnerf_b200/frame.py) */
    const float* head_consts;    /* [3*64] from rn_frame_conditioning */
    void* consts_ready_event;    /* optional cudaEvent_t: rn_frame_conditioning ran on ANOTHER stream and recorded this event; the
                                    head waits for it only before its first network evaluation, so the audio nets overlap the
                                    ray setup and the first march.  NULL = same stream, no wait. */
    uint32_t capture_unroll;     /* only used while `stream` is being captured into a CUDA graph: the first capture_unroll
                                    iterations (0 = 1) of the march/evaluate/composite loop are captured as plain kernel nodes,
                                    the rest is ONE conditional WHILE node whose condition the device-side loop controller sets.
                                    An unrolled iteration that finds the loop finished costs ~1.3 us per kernel, a WHILE
                                    iteration ~7 us more than an unrolled one: pass the iteration count typical of the scene.
                                    Outside capture all max_steps iterations are launched. */
    uint32_t occ_words;          /* with occ_pack: 32-bit words of packed occupancy bits the marcher stages in shared memory (the pack's
                                    total_words as read back by the host when the pack was built; sizes the launch) */
    const float* occ_aabb;       /* optional device [6] (xmin,ymin,zmin,xmax,ymax,zmax): a CONSERVATIVE bounding box of every occupied
                                    cell of the bitfield, all cascades, inflated by at least one cell.  A ray whose [near, far]
                                    segment provably misses it cannot emit a sample, so it is not marched (same outputs; about
                                    70% of the rays of a talking-head frame).  NULL = march every ray. */
    const void* occ_pack;        /* optional: rn_occupancy_pack() output for `bitfield` -- the occupied boxes of every cascade as linear
                                    bit arrays; each marcher CTA copies them into SHARED memory and probes there (same cells, same
                                    samples; replaces the global-memory probes of kernel_march_rays, raymarching.cu:892-905).  Its
                                    bytes 16..39 are the occ_aabb box.  NULL (or a pack marked unusable) = probe the Morton bitfield. */
} rn_frame_head_desc;

/* Box-packed copy of the occupancy bitfield for shared-memory staging (csrc/occ_pack.cu).  pack: rn_occupancy_pack_bytes() bytes,
 * 16-byte aligned.  Header (int32): [0] cascades, [1] total_words, [2] usable; floats [4..9] = world box of all occupied cells inflated
 * by one cell.  Run whenever the bitfield changes. */
int rn_occupancy_pack(const uint8_t* bitfield, uint32_t C, uint32_t H, float bound, void* pack, void* stream);
uint32_t rn_occupancy_pack_bytes(void);

typedef struct rn_frame_torso_desc {
    uint32_t N, grid_size;
    float thresh, shrink;
    const float* bg_coords;      /* [N,2] */
    const float* density_grid_torso; /* [grid_size^2] */
    void* workspace; uint64_t workspace_bytes;
    rn_grid_table grid2d;
    const void* torso_blob; const float* torso_consts;
    float* torso_alpha; float* torso_color;   /* out [N], [N,3] */
} rn_frame_torso_desc;

uint64_t rn_frame_workspace_bytes(uint32_t N);
uint32_t rn_head_blob_bytes(void);
uint32_t rn_torso_blob_bytes(void);
int rn_frame_conditioning(const rn_conditioning_desc* d, void* stream);
int rn_frame_head(const rn_frame_head_desc* d, void* stream);
int rn_frame_torso(const rn_frame_torso_desc* d, void* stream);
/* rn_frame_head bracketed by CUDA events on the launching stream (bench / diagnostics; synchronises the stream):
 * ms [3*max_steps] = (march, eval, composite) per iteration, n_samples [max_steps] = samples evaluated per iteration */
int rn_frame_head_timed(const rn_frame_head_desc* d, void* stream, float* ms, uint32_t* n_samples);
/* diagnostics: device buffer of 8 uint64 cycle counters filled by the head kernel; NULL disables */
void rn_debug_set_head_prof(void* counters);
/* final blend + depth normalisation (nerf/renderer.py:299-310); bg_color [N,3] or NULL (then bg_scalar);
 * torso_alpha/torso_color NULL when there is no torso; torso_bg_out (nullable) receives results['torso_color'] */
int rn_frame_finalize(uint32_t N, const float* weights_sum, float* depth, float* image, const float* nears,
                      const float* fars, const float* bg_color, float bg_scalar, const float* torso_alpha,
                      const float* torso_color, float* torso_bg_out, void* stream);

/* ---- streamed frames: the per-frame host work of radnerf_b200.stream.FrameStreamer in ONE call ------------------------
 * (the frame loop of Trainer.test, nerf/utils.py:905-960, in steady state: every buffer, stream, event and the lane's
 * captured frame graph are fixed; see csrc/pipeline.cu for the choreography).  Events come from rn_event_create.
 * phase bit 0: copy-in, rays, conditioning, frame graph [, scatter to peers]; bit 1: [staging + device->host copy,] ev_done.
 * Both bits in one call when there is no cross-rank barrier to issue in between (one GPU, or the flag-based gather-to-root). */
typedef struct rn_lane_submit {
    void* lane_stream; void* cond_stream; void* copy_stream;
    void* ev_in; void* ev_cond; void* ev_done; void* ev_staged; void* ev_delivered;
    const void* packed_src; void* flat_dst; uint64_t packed_bytes;        /* input block: pinned host or device -> graph input block */
    const float* pose; float fx, fy, cx, cy; uint32_t H, W;                /* rn_get_rays arguments; n_rays = 0 skips it */
    const int32_t* pixel_ids; uint32_t n_rays; uint32_t graph_kernels;
    float* rays_o; float* rays_d;
    const rn_conditioning_desc* cond;
    void* graph_exec;                                                      /* cudaGraphExec_t of the lane's frame */
    const float* image_local; const int32_t* ids; const uint64_t* peers;   /* rn_scatter_rows_to_peers arguments; peers NULL skips it */
    uint32_t n_local, run_pixels, world, phase;
    const void* stage_src; void* stage_dst; void* host_dst; uint64_t image_bytes;   /* host_dst NULL: no delivery */
    uint32_t to_uint8, reserved;   /* to_uint8: stage_src holds image_bytes/4 fp32 values; stage_dst / host_dst receive image_bytes/4 bytes */
    /* gather-to-root exchange (rn_scatter_rows_to_root / rn_stage_frame_at_root); ctrl_peers NULL: the older all-to-all scatter, whose
     * cross-rank barrier the caller issues between phase 1 and phase 2 */
    const uint64_t* ctrl_peers; uint32_t* ticket; uint64_t frame_seq;
    uint32_t rank, root, slot, scatter_ctas;
} rn_lane_submit;
int rn_lane_submit_frame(const rn_lane_submit* s);
int rn_event_create(void** ev);
int rn_event_destroy(void* ev);
int rn_event_synchronize(void* ev);
int rn_stream_wait_event(void* stream, void* ev);
/* sizeof() of a public descriptor struct by name ("rn_lane_submit", ...), 0 if unknown: lets a binding check its mirror */
uint32_t rn_sizeof(const char* name);

/* ------------------------------------------------------------------ training-step tail -------------- */

/* One sweep over every trainable tensor for what the reference ends a training step with (nerf/utils.py:1171-1182):
 * `scaler.step(optimizer)` on torch.optim.Adam (main.py:204; groups from NeRFNetwork.get_params, nerf/network.py:329-361)
 * [+ `optimizer.zero_grad()` of the next step, :1164] and `ema.update()` (torch_ema, :1181-1182).
 * A tensor is cut into chunks of RN_ADAM_CHUNK elements; `first_chunk` of tensor t = sum of ceil(n / RN_ADAM_CHUNK) of the
 * tensors before it, n_chunks = the total.  The descriptor array lives in DEVICE memory, the group table in HOST memory
 * (it changes every step with the learning-rate schedule and travels as a kernel argument). */
#define RN_ADAM_CHUNK 4096u
#define RN_ADAM_MAX_GROUPS 32u
#define RN_ADAM_ZERO_GRADS 1u      /* flags bit 0: leave every gradient zeroed (the next step's zero_grad) */
#define RN_ADAM_GROUPS_ON_DEVICE 2u /* flags bit 1: `groups` is a DEVICE table kept current with rn_adam_groups_store -- for a
                                      step captured in a CUDA graph, whose kernel arguments are frozen while the learning
                                      rate moves on */
typedef struct rn_adam_tensor {
    float* param; float* grad; float* exp_avg; float* exp_avg_sq;   /* [n] fp32 each */
    float* step;                 /* device scalar: optimiser steps this tensor has taken (torch keeps state['step'] the same way
                                    for its fused/capturable Adam); advanced by rn_adam_step unless the step is skipped */
    float* ema;                  /* rn_ema_update only: the shadow copy */
    uint64_t n;
    uint32_t first_chunk, group;
} rn_adam_tensor;
typedef struct rn_adam_group { double lr, beta1, beta2, eps, weight_decay; } rn_adam_group;
/* grad_scale / found_inf: nullable device scalars with torch.amp.GradScaler's meaning (gradients are divided by
 * *grad_scale; *found_inf != 0 skips the whole step -- parameters, moments and step counters untouched, gradients still
 * zeroed when asked).  Arithmetic: torch/optim/adam.py `_single_tensor_adam` (amsgrad / maximize off), fp32. */
int rn_adam_step(const rn_adam_tensor* tensors, uint32_t n_tensors, uint32_t n_chunks, const rn_adam_group* groups,
                 uint32_t n_groups, const float* grad_scale, const float* found_inf, uint32_t flags, void* stream);
/* writes the HOST table `groups` into the device table `groups_dev` [>= n_groups] in stream order (a kernel whose argument
 * is the table: no host memory is read after the call returns) */
int rn_adam_groups_store(const rn_adam_group* groups, uint32_t n_groups, rn_adam_group* groups_dev, void* stream);
/* shadow -= (1 - decay) * (shadow - param) for every tensor (param, ema, n, first_chunk of the descriptors are used) */
int rn_ema_update(const rn_adam_tensor* tensors, uint32_t n_tensors, uint32_t n_chunks, double decay, void* stream);

/* ------------------------------------------------------------------ streaming audio hand-off --------- */

/* replaces ASR.get_next_feat (nerf/asr.py:160-183; consumer nerf/gui.py:186): the [RN_RING_DEPTH, dim, RN_RING_WINDOW] block
 * of one video frame from the ASR feature ring `ring [size, dim]` (asr.py:103) in one launch.  Window k = rows
 * (start[k] + j) mod size, j < 16; start[k] < 0 = an all-zero window (asr.py:109).  snapshot[k] >= 0: the window is a COPY
 * held in slot snapshot[k] of `snapshots [RN_RING_DEPTH, 16, dim]` (the reference's torch.cat for a window that wraps
 * around the ring end); fresh[k] != 0 takes that copy now.  snapshot[k] < 0: read live from the ring (the reference's
 * slice views show later overwrites).  `out` may point anywhere, e.g. into a frame lane's input block. */
#define RN_RING_WINDOW 16
#define RN_RING_DEPTH 8
typedef struct rn_ring_windows { int32_t start[RN_RING_DEPTH]; int32_t snapshot[RN_RING_DEPTH]; int32_t fresh[RN_RING_DEPTH]; } rn_ring_windows;
int rn_feature_window(const float* ring, uint32_t size, uint32_t dim, const rn_ring_windows* windows, float* snapshots,
                      float* out, void* stream);

/* diagnostics (tools/, not part of the operator contract) */
void rn_debug_set_audio_prof(void* stamps);
void rn_debug_set_max_iters(uint32_t n);
void rn_debug_set_while_node(int on);

/* ------------------------------------------------------------------ diagnostics ---------------------- */

/* ------------------------------------------------------------------ fused training step of the head network ------------
 * Replaces, for one batch of march_rays_train samples, NeRFNetwork.forward (nerf/network.py:222-283: two grid encoders, SH, the
 * ambient / sigma / colour MLPs, trunc_exp, sigmoid) and its backward (autograd through the same, nerf/utils.py:1168-1171) --
 * in the reference ~16 cuBLAS GEMMs and ~300 elementwise / cat / repeat / cast launches per step -- by one forward kernel and two
 * backward kernels on tcgen05 (csrc/head_train_fwd.cu, head_train_bwd.cu) plus the grid scatter kernels.
 * All buffers are caller-allocated; sizes come from the rn_head_train_*_bytes queries. */
typedef struct rn_head_train_desc {
    uint32_t M;                  /* sample rows of xyzs / dirs (the tail of the last 128-row tile is padded internally) */
    uint32_t reserved;
    float bound; float reserved1;
    const float* xyzs; const float* dirs;            /* [M,3] fp32: positions in [-bound, bound], unit directions */
    rn_grid_table grid3d, grid2d;                    /* packed fp16 tables (rn_pack_grid_table) of encoder / encoder_ambient */
    const void* fwd_blob;                            /* rn_head_blob_bytes() of interleaved fp16 weights (rn_pack_head_blobs) */
    const void* bwd_blob;                            /* rn_head_train_bwd_blob_bytes(): transposed weights for the data gradients */
    const float* consts;                             /* [3*64] hoisted first-layer terms: W_a1[:,32:96] enc_a, W_s1[:,64] eye, W_c1[:,80:84] ind */
    float* sigma; float* rgb; float* ambient;        /* forward outputs [M], [M,3], [M,2] fp32 (rgb, sigma rounded as the fp16 autocast path) */
    float* sigma_pre;                                /* [M] log-density (input of trunc_exp), kept for its backward */
    void* acts;                                      /* rn_head_train_acts_bytes(M): saved layer inputs, 928 B/sample */
    void* dy_dx2;                                    /* [M, 16*2*2] fp16: d(enc_w)/d(ambient coordinate in [0,1]), gridencoder.cu:200-243 */
    const float* d_sigma; const float* d_rgb; const float* d_ambient;   /* backward inputs [M], [M,3], [M,2] fp32 */
    float* d_table3; float* d_table2;                /* fp32 [rows,2] table gradients, accumulated INTO (caller pre-zeroes) */
    float* d_weights;                                /* [rn_head_train_dw_floats()] fp32, overwritten: nn.Linear-shaped gradients of the
                                                        encoder-fed weight columns + the three column sums that carry the hoisted
                                                        columns (layout: csrc/head_train.cuh G_*) */
    void* workspace; uint64_t workspace_bytes;       /* rn_head_train_workspace_bytes(M) */
    const int32_t* m_valid;                          /* optional DEVICE scalar: only rows [0, min(M, *m_valid)) hold samples (the marcher's
                                                        counter, clamped to its budget); the rest of the buffers is padding -- zero
                                                        positions that would all fall into the same grid cells.  Padding tiles are not
                                                        evaluated (forward outputs there are left untouched: pre-zero them) and are marked
                                                        out of range for the table scatters.  NULL: all M rows are samples. */
} rn_head_train_desc;
uint64_t rn_head_train_acts_bytes(uint32_t M);
uint64_t rn_head_train_workspace_bytes(uint32_t M);
uint32_t rn_head_train_bwd_blob_bytes(void);
uint32_t rn_head_train_dw_floats(void);
int rn_head_train_forward(const rn_head_train_desc* d, void* stream);
int rn_head_train_backward(const rn_head_train_desc* d, void* stream);
/* fp32 GridEncoder table [rows,2] -> the packed fp16 copy the fused kernels gather from: row r of level l (first row
 * packed_first[l]) = (features of row r, features of row (r + 1) mod size_l); radnerf_b200/frame.py pack_table in one launch. */
int rn_pack_grid_table(const float* embeddings, const int32_t* offsets, const int32_t* packed_first, uint32_t L, void* out, void* stream);
/* the eight bias-free Linear weights of the head (fp32, nn.Linear [out,in]: ambient 64x96, 64x64, 2x64; sigma 64x65, 64x64, 65x64;
 * colour 64x84, 3x64) -> the interleaved fp16 operand blobs of the forward (rn_head_blob_bytes) and, if bwd_blob != NULL, of the
 * backward (rn_head_train_bwd_blob_bytes) in one launch. */
int rn_pack_head_blobs(const float* const* weights8, void* fwd_blob, void* bwd_blob, void* stream);

/* one 128 x N x K fp16 GEMM tile through the hand-written tcgen05/TMEM path (out = A @ W^T, fp32 accumulate);
 * A [128,K] fp16 row-major, W [N,K] fp16 row-major.  Validates descriptors/layouts in isolation. */
int rn_selftest_umma(const void* A, const void* W, float* out, uint32_t K, uint32_t N, void* stream);
/* the weight-gradient contraction of the fused training step: out[m,n] = passes * sum_s X[s,m] * Y[s,n] with both operands read
 * MN-major from interleaved [128 x K] tiles (samples = the MMA's K dimension); rows m >= Kx of out [128,Ky] are undefined. */
int rn_selftest_umma_mn(const void* X, const void* Y, float* out, uint32_t Kx, uint32_t Ky, uint32_t passes, void* stream);

/* table-gradient scatter of a 3-D, 2-feature, linearly interpolated grid (replaces kernel_grid_backward,
 * gridencoder.cu:247-339, for the head's spatial encoder): grad [B, L*2] (RN_LAYOUT_BLC) of `dtype`, inputs [B,3] in [0,1],
 * grad_table fp32 [rows,2] accumulated INTO.  variant: bit0 merge the two z-corners of levels that never index z + pair the
 * x-corners into 16-byte atomics, bit1 segmented warp reduction of consecutive samples in one cell on levels < agg_levels,
 * bit2 accumulate the first priv_levels (dense) levels, priv_rows rows in total, in shared memory.  level_mask bit l = do level l.
 * rn_grid_encode_backward dispatches here with the production settings. */
#define RN_BWD3_PRODUCTION 3u   /* z-merge + x-pair + segmented warp aggregation (profiles/r02_bwd3_variants_*.json) */
#define RN_BWD3_AGG_LEVELS 6u
int rn_grid_backward3(const void* grad, const float* inputs, const int32_t* offsets, float* grad_table, uint32_t B, uint32_t L,
                      float S, uint32_t H, uint32_t gridtype, uint32_t dtype, uint32_t variant, uint32_t level_mask,
                      uint32_t agg_levels, uint32_t priv_levels, uint32_t priv_rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADNERF_B200_H */
