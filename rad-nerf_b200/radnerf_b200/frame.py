"""Fused inference frame: host side of rn_frame_* (include/radnerf_b200.h, csrc/frame*.cu, head_eval.cu, torso_eval.cu).

`render_frame(model, ...)` has the contract of NeRFRenderer.run_cuda's inference branch (nerf/renderer.py:158-316: same
arguments, same result keys) but issues a fixed sequence of launches with no host synchronisation: per-frame conditioning
(audio nets, lip smoothing, hoisted first-layer terms) -> device-driven march / fused-network / composite loop -> torso ->
final blend.  Host work per frame is a handful of ctypes calls; everything data-dependent stays on the device.

State cached on the model: `FusedShared` -- packed fp16 copies of the three hash tables, the interleaved fp16 weight blobs
(rebuilt only when a parameter's version counter changes), the lip-smoothing state, the occupied-cell box -- and one
`FusedState` per frame LANE (`model._fused` = lane 0, `lane_state(model, k)`): workspace, hoisted-term vectors, captured
CUDA graphs with their input buffers.  Several lanes = several frames in flight (radnerf_b200.stream.FramePipeline), with
`launch_conditioning` / `render_frame(..., lane=k, external_cond=True)` / `replay_lane` as the building blocks.
"""
import ctypes as C

import numpy as np
import torch

from . import abi

_vp, _u32, _f32, _u64 = C.c_void_p, C.c_uint32, C.c_float, C.c_uint64


class GridTable(C.Structure):
    _fields_ = [("table_f16", _vp), ("offsets", _vp), ("S", _f32), ("H", _u32), ("packed_offsets", _vp)]


class ConditioningDesc(C.Structure):
    _fields_ = [("auds", _vp), ("F", _u32), ("Cin", _u32), ("att", _u32), ("smooth", _u32), ("reserved", _u32),
                ("conv_w", _vp * 4), ("conv_b", _vp * 4), ("fc_w", _vp * 2), ("fc_b", _vp * 2),
                ("att_w", _vp * 5), ("att_b", _vp * 5), ("att_fc_w", _vp), ("att_fc_b", _vp),
                ("enc_a_state", _vp), ("lambda_", _f32),
                ("w_amb1", _vp), ("w_sig1", _vp), ("w_col1", _vp), ("eye", _vp), ("ind_code", _vp), ("head_consts", _vp),
                ("w_def1", _vp), ("w_tor1", _vp), ("pose6", _vp), ("ind_torso", _vp), ("torso_consts", _vp),
                ("pose44", _vp), ("pose6_out", _vp)]


class FrameHeadDesc(C.Structure):
    _fields_ = [("N", _u32), ("max_steps", _u32), ("cascade", _u32), ("grid_size", _u32),
                ("bound", _f32), ("min_near", _f32), ("dt_gamma", _f32), ("T_thresh", _f32),
                ("rays_o", _vp), ("rays_d", _vp), ("aabb", _vp), ("bitfield", _vp), ("noises", _vp),
                ("weights_sum", _vp), ("depth", _vp), ("image", _vp), ("nears", _vp), ("fars", _vp),
                ("workspace", _vp), ("workspace_bytes", _u64),
                ("grid3d", GridTable), ("grid2d", GridTable), ("head_blob", _vp), ("head_consts", _vp), ("consts_ready_event", _vp),
                ("capture_unroll", _u32), ("occ_words", _u32), ("occ_aabb", _vp), ("occ_pack", _vp)]


class FrameTorsoDesc(C.Structure):
    _fields_ = [("N", _u32), ("grid_size", _u32), ("thresh", _f32), ("shrink", _f32),
                ("bg_coords", _vp), ("density_grid_torso", _vp), ("workspace", _vp), ("workspace_bytes", _u64),
                ("grid2d", GridTable), ("torso_blob", _vp), ("torso_consts", _vp), ("torso_alpha", _vp), ("torso_color", _vp)]


abi.register("rn_frame_workspace_bytes", [_u32], _u64)
abi.register("rn_note_graph_replay", [_u64], None)
abi.register("rn_head_blob_bytes", [], _u32)
abi.register("rn_torso_blob_bytes", [], _u32)
abi.register("rn_frame_conditioning", [C.POINTER(ConditioningDesc), _vp])
abi.register("rn_frame_head", [C.POINTER(FrameHeadDesc), _vp])
abi.register("rn_frame_torso", [C.POINTER(FrameTorsoDesc), _vp])
abi.register("rn_frame_finalize", [_u32, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp, _vp, _vp])


def _p(t):
    return None if t is None else t.data_ptr()


def il_pack(W, n_pad, k_pad):
    """[n,k] weight -> fp16 [n_pad,k_pad] (zero padded) in the interleaved UMMA operand layout of csrc/umma.cuh:
    8-row x 16-byte core matrices, K-chunks of an 8-row group contiguous."""
    n, k = W.shape
    Wp = torch.zeros(n_pad, k_pad, dtype=torch.float16, device=W.device)
    Wp[:n, :k] = W.detach().to(torch.float16)
    return Wp.view(n_pad // 8, 8, k_pad // 8, 8).permute(0, 2, 1, 3).contiguous().view(-1)


def supported(model):
    """the fused kernels are specialised for the architecture nerf/network.py builds"""
    try:
        ok = (model.encoder.num_levels == 16 and model.encoder.level_dim == 2 and model.encoder.gridtype == 'tiled'
              and model.encoder.interpolation == 'linear' and not model.encoder.align_corners
              and model.encoder_ambient.num_levels == 16 and model.encoder_ambient.input_dim == 2
              and model.hidden_dim == 64 and model.geo_feat_dim == 64 and model.num_layers == 3 and model.num_layers_color == 2
              and model.hidden_dim_ambient == 64 and model.num_layers_ambient == 3 and model.audio_dim == 64
              and model.encoder_dir.degree == 4 and model.individual_dim == 4 and model.exp_eye and not model.emb
              and model.audio_in_dim <= 44 and model.opt.fp16)
        if model.torso:
            ok = ok and model.individual_dim_torso == 8 and model.torso_encoder.num_levels == 16
        return bool(ok)
    except AttributeError:
        return False


class FusedShared:
    """What every frame in flight shares: fp16 copies of tables / weight blobs (read-only for the kernels), the
    lip-smoothing state that chains consecutive frames, the occupied-cell box."""

    def __init__(self, model):
        self.dev = model.density_bitfield.device
        self.versions = None
        self.generation = 0
        self.enc_a_state = torch.zeros(65, device=self.dev)  # [0..63] smoothed audio code, [64] validity flag (device side)

    def refresh(self, model):
        params = getattr(self, "_params", None)
        if params is None:   # the module tree is fixed: walk it once, then only compare (pointer, version) per frame
            params = [model.encoder.embeddings, model.encoder_ambient.embeddings] + list(model.ambient_net.parameters()) + \
                list(model.sigma_net.parameters()) + list(model.color_net.parameters()) + list(model.audio_net.parameters())
            if getattr(model, 'audio_att_net', None) is not None:
                params += list(model.audio_att_net.parameters())
            if model.torso:
                params += [model.torso_encoder.embeddings] + list(model.torso_deform_net.parameters()) + list(model.torso_net.parameters())
            params += [p for p in (getattr(model, "individual_codes", None), getattr(model, "individual_codes_torso", None)) if p is not None]
            self._params = params
        versions = tuple((p.data_ptr(), p._version) for p in params)
        if versions == self.versions:
            return
        self.versions = versions
        self.generation += 1   # lanes drop their captured graphs: they hold pointers to the old blobs
        self.h16 = {}  # fp16 copies of the small-net parameters used by the conditioning kernel
        self.table3 = pack_table(model.encoder)
        self.table2 = pack_table(model.encoder_ambient)
        a, s, c = model.ambient_net.net, model.sigma_net.net, model.color_net.net
        w_s3 = s[2].weight
        w_s3 = torch.cat([w_s3[1:], w_s3[:1]], 0)  # geo_feat rows first, log-density row last (head_eval.cu)
        self.head_blob = torch.cat([
            il_pack(a[0].weight[:, :32], 64, 32), il_pack(a[1].weight, 64, 64), il_pack(a[2].weight, 16, 64),
            il_pack(s[0].weight[:, 0:32], 64, 32), il_pack(s[0].weight[:, 32:64], 64, 32), il_pack(s[1].weight, 64, 64),
            il_pack(w_s3, 80, 64), il_pack(c[0].weight[:, :80], 64, 80), il_pack(c[1].weight, 16, 64)])
        assert self.head_blob.numel() * 2 == abi.lib().rn_head_blob_bytes()
        if model.torso:
            self.table_t = pack_table(model.torso_encoder)
            d, t = model.torso_deform_net.net, model.torso_net.net
            self.torso_blob = torch.cat([
                il_pack(d[0].weight[:, :42], 64, 48), il_pack(d[1].weight, 64, 64), il_pack(d[2].weight, 16, 64),
                il_pack(t[0].weight[:, :74], 32, 80), il_pack(t[1].weight, 32, 32), il_pack(t[2].weight, 16, 32)])
            assert self.torso_blob.numel() * 2 == abi.lib().rn_torso_blob_bytes()

    def occupied_box(self, model):
        """device [6]: bounding box of all occupied cells of the density bitfield (every cascade, world coordinates), inflated
        by one cell per side -- rn_frame_head_desc.occ_aabb.  It lives inside the box-packed copy of the bitfield that the fused
        marcher stages in shared memory (rn_occupancy_pack, csrc/occ_pack.cu): `self.occ_pack` / `self.occ_words`.  Rebuilt into the
        SAME buffer when the bitfield changes (three small launches and one 16-byte read-back of the pack's size)."""
        bf = model.density_bitfield
        tag = (bf.data_ptr(), bf._version)
        if getattr(self, "_occ_tag", None) != tag:
            L = abi.lib()
            if getattr(self, "occ_pack", None) is None or self.occ_pack.device != bf.device:
                self.occ_pack = torch.zeros(int(L.rn_occupancy_pack_bytes()), dtype=torch.uint8, device=bf.device)
                self._occ_box = self.occ_pack[16:40].view(torch.float32)
            with torch.cuda.device(bf.device):
                abi.check(L.rn_occupancy_pack(abi.ptr(bf), int(model.cascade), int(model.grid_size), float(model.bound), abi.ptr(self.occ_pack),
                                              abi.cur_stream()), "rn_occupancy_pack")
            words, usable = self.occ_pack[4:12].view(torch.int32).tolist()     # outside any capture: a bitfield change is not steady state
            self.occ_words = int(words) if usable else 0
            self._occ_tag = tag
        return self._occ_box


# Staging the box-packed occupancy bits in shared memory (north_star's sketch; csrc/occ_pack.cu, march_compact_kernel) is implemented and
# bit-exact, but it is NOT faster on B200 and therefore off by default (RADNERF_OCC_PACK=1 turns it on): the DDA is bound by the ~100
# dependent instructions of a probe, not by the bitfield load, which hits L1 for the coherent rays of a frame.  Measured per launch at
# 512x512 (tools/frame_timeline.py, profiles/r02_frame_timeline_512_occ_pack.txt): 6.2 / 9.2 / 8.6 / 8.2 / 7.4 us staged vs 4.1 / 8.2 / 7.7 /
# 7.2 / 6.4 us from global memory; frames/s 3 318 vs 3 432.
_USE_OCC_PACK = bool(int(__import__("os").environ.get("RADNERF_OCC_PACK", "0")))


class FusedState:
    """Per-LANE state of the fused renderer (a lane = one frame in flight): workspace, hoisted-term vectors, captured
    graphs and their input buffers, forked stream.  Weights and the smoothing state live in `shared` (attribute
    look-ups fall through to it)."""

    def __init__(self, model, shared=None):
        dev = model.density_bitfield.device
        self.dev = dev
        self.shared = shared if shared is not None else FusedShared(model)
        self.generation = -1
        self.N = 0
        self.workspace = None
        self.head_consts = torch.zeros(3 * 64, device=dev)
        self.torso_consts = torch.zeros(96, device=dev)
        self.frames = 0
        self.graphs = {}      # config key -> (CUDAGraph, static inputs, outputs)
        self.side = torch.cuda.Stream(device=dev)
        self.cond_event = torch.cuda.Event()
        self.use_graph = True
        self.last_static = None   # input buffers of the graph used by the last frame
        self.capture_unroll = 1   # loop iterations captured as plain nodes; set from the warm-up frame before each capture

    def __getattr__(self, name):   # only called when normal look-up fails: shared weights / state
        if name == "shared":
            raise AttributeError(name)
        return getattr(self.shared, name)

    def refresh_weights(self, model):
        self.shared.refresh(model)
        if self.generation != self.shared.generation:
            self.generation = self.shared.generation
            self.graphs.clear()

    def occupied_box(self, model):
        return self.shared.occupied_box(model)

    def ensure_workspace(self, N):
        if N != self.N:
            nbytes = int(abi.lib().rn_frame_workspace_bytes(N))
            self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)  # zeroed: the stats block is never reset by the library
            self.ws_bytes = nbytes
            self.N = N
            self.torso_alpha = torch.empty(N, 1, device=self.dev)
            self.torso_color = torch.empty(N, 3, device=self.dev)
            self.graphs.clear()

    def loop_iterations(self):
        """march/evaluate/composite iterations executed since the workspace was created (device counter; forces a sync).
        A captured frame holds iteration 0 plus ONE conditional WHILE node, so the kernels a replay launches are
        n_captured + 3 * (iterations of that frame - 2): bench.py uses this counter for its `gpu_launches` claim."""
        return int(self.workspace[-256:-252].view(torch.int32).item())

    def ctl(self):
        """device loop state of the last frame as a [65, 8] int32 tensor (n_alive, n_step, step, done, n_samples, ...)"""
        return self.workspace[:65 * 32].view(torch.int32).view(65, 8)


def pack_table(enc):
    """fp16 copy of a GridEncoder's table for the fused kernels (rn_grid_table):
      * every level whose size is a power of two (the size-capped, wrapping levels) starts at a multiple of its size, so a row
        index is ((i & (size-1)) | first_row);
      * a row holds the cell's two features AND the two features of its +x neighbour (row i+1 of the level, wrapped for a
        wrapping level), so the kernels fetch both x-corners of a pair with one 8-byte load.
    Returns (table [rows', 4] fp16, first rows [L] int32)."""
    offs = enc.offsets.cpu().numpy().astype(np.int64)
    emb = enc.embeddings.detach()
    assert emb.shape[1] == 2
    first, cur = [], 0
    for l in range(len(offs) - 1):
        size = int(offs[l + 1] - offs[l])
        align = size if size & (size - 1) == 0 else 8
        cur = (cur + align - 1) // align * align
        first.append(cur)
        cur += size
    out = torch.zeros(cur, 4, dtype=torch.float16, device=emb.device)
    for l, f in enumerate(first):
        lvl = emb[int(offs[l]):int(offs[l + 1])].to(torch.float16)
        out[f:f + lvl.shape[0], :2] = lvl
        out[f:f + lvl.shape[0], 2:] = torch.roll(lvl, -1, 0)   # +x neighbour; the last row's wraps (only used by wrapping levels)
    return out, torch.tensor(first, dtype=torch.int32, device=emb.device)


def _grid_table(enc, packed):
    table, first = packed
    return GridTable(table.data_ptr(), enc.offsets.data_ptr(), float(np.log2(enc.per_level_scale)), int(enc.base_resolution),
                     first.data_ptr())


def conditioning_desc(model, st, auds, eye_t, pose6, pose44=None):
    """pose44: device [4,4] cam2world -- the kernel then derives the 6-vector itself (the reference's convert_poses on the
    device) and writes it back into `pose6` (if given) for inspection"""
    an, at = model.audio_net, getattr(model, "audio_att_net", None)

    def H(param):  # cached fp16 copy (cleared by refresh_weights when any parameter changes)
        t = st.h16.get(id(param))
        if t is None:
            t = st.h16[id(param)] = param.detach().to(torch.float16).contiguous()
        return t.data_ptr()

    cd = ConditioningDesc()
    if auds is not None:
        cd.auds, cd.F, cd.Cin = auds.data_ptr(), auds.shape[0], auds.shape[1]
    cd.att = int(model.att)
    cd.smooth = int(bool(model.smooth_lips))
    for i, k in enumerate((0, 2, 4, 6)):
        cd.conv_w[i], cd.conv_b[i] = H(an.encoder_conv[k].weight), H(an.encoder_conv[k].bias)
    for i, k in enumerate((0, 2)):
        cd.fc_w[i], cd.fc_b[i] = H(an.encoder_fc1[k].weight), H(an.encoder_fc1[k].bias)
    if at is not None:
        for i, k in enumerate((0, 2, 4, 6, 8)):
            cd.att_w[i], cd.att_b[i] = H(at.attentionConvNet[k].weight), H(at.attentionConvNet[k].bias)
        cd.att_fc_w, cd.att_fc_b = H(at.attentionNet[0].weight), H(at.attentionNet[0].bias)
    cd.enc_a_state, cd.lambda_ = st.enc_a_state.data_ptr(), 0.35
    cd.w_amb1 = H(model.ambient_net.net[0].weight)
    cd.w_sig1 = H(model.sigma_net.net[0].weight)
    cd.w_col1 = H(model.color_net.net[0].weight)
    cd.eye = _p(eye_t)
    cd.ind_code = model.individual_codes.data_ptr()  # row 0 (inference uses a fixed code, renderer.py:201-202)
    cd.head_consts = st.head_consts.data_ptr()
    if model.torso:
        cd.w_def1 = H(model.torso_deform_net.net[0].weight)
        cd.w_tor1 = H(model.torso_net.net[0].weight)
        cd.ind_torso, cd.torso_consts = model.individual_codes_torso.data_ptr(), st.torso_consts.data_ptr()
        if pose44 is not None:
            cd.pose44, cd.pose6_out = pose44.data_ptr(), _p(pose6)
        else:
            cd.pose6 = pose6.data_ptr()
    return cd


def lane_state(model, lane=0):
    """FusedState of frame lane `lane` (lane 0 is `model._fused`); all lanes share one FusedShared"""
    abi.lib()
    st0 = getattr(model, "_fused", None)
    if st0 is None:
        st0 = model._fused = FusedState(model)
    if lane == 0:
        return st0
    lanes = getattr(model, "_fused_lanes", None)
    if lanes is None:
        lanes = model._fused_lanes = {}
    if lane not in lanes:
        lanes[lane] = FusedState(model, shared=st0.shared)
    return lanes[lane]


def _sync_smoothing_state(model, st):
    """`model.enc_a` (the reference's attribute) mirrors the device-side state"""
    if model.smooth_lips:
        prev = getattr(model, "enc_a", None)
        if prev is None:
            st.enc_a_state.zero_()
        elif prev.data_ptr() != st.enc_a_state.data_ptr():
            st.enc_a_state[:64].copy_(prev.reshape(-1))
            st.enc_a_state[64] = 1.0


def launch_conditioning(model, lane, auds, eye=None, poses=None, pose44=None):
    """The conditioning kernel of ONE frame on the current stream, writing lane `lane`'s hoisted-term vectors and advancing
    the shared lip-smoothing state.  Callers that keep several frames in flight (radnerf_b200.stream.FramePipeline) run these
    in frame order on one stream and render the frame with render_frame(..., lane=lane, external_cond=True)."""
    st = lane_state(model, lane)
    st.refresh_weights(model)
    _sync_smoothing_state(model, st)
    auds_t = None if auds is None else auds.contiguous().float()
    eye_t = None if eye is None else eye.reshape(-1).float().contiguous()
    pose6 = poses.reshape(-1).float().contiguous() if (model.torso and poses is not None) else None
    if model.torso and pose44 is not None:
        pose44 = pose44.reshape(-1).float().contiguous()   # device-side convert_poses; `pose6` (if any) receives the result
    else:
        pose44 = None
    st._keep = (auds_t, eye_t, pose6, pose44)   # alive until the kernel has run
    # callers that stream frames pass the same (static) buffers every time: build the descriptor once per set of pointers
    key = (None if auds_t is None else (auds_t.data_ptr(), tuple(auds_t.shape)), None if eye_t is None else eye_t.data_ptr(),
           None if pose6 is None else pose6.data_ptr(), None if pose44 is None else pose44.data_ptr(), st.shared.generation)
    cache = st.__dict__.setdefault("_cd_cache", {})
    cd = cache.get(key)
    if cd is None:
        cache.clear()
        cd = cache[key] = conditioning_desc(model, st, auds_t, eye_t, pose6, pose44)
    abi.check(abi.lib().rn_frame_conditioning(C.byref(cd), abi.cur_stream()))
    if model.smooth_lips and auds is not None:
        model.enc_a = st.enc_a_state[:64].view(1, 64)


def replay_lane(model, lane):
    """Replay lane `lane`'s last captured frame graph as is: the caller has written the frame's inputs into that graph's
    input buffers (FusedState.last_static) and run launch_conditioning for the lane.  Returns the graph's output dict, or
    None when there is nothing to replay (no graph yet, or weights / occupancy changed): call render_frame then."""
    st = lane_state(model, lane)
    gen = st.shared.generation
    st.refresh_weights(model)
    st.occupied_box(model)
    entry = st.__dict__.get("last_entry")
    if entry is None or gen != st.shared.generation or not st.graphs:
        return None
    graph, static, outs, n_kernels = entry
    graph.replay()
    abi.lib().rn_note_graph_replay(n_kernels)
    st.frames += 1
    return outs


def advance_conditioning(model, auds, eye=None, poses=None):
    """Run only the per-frame conditioning kernel (audio nets + lip-smoothing EMA) for a frame that is NOT rendered here:
    a rank that renders a slice of a sequence calls this for the frames just before its slice so that its smoothing
    state matches a run over the whole sequence (radnerf_b200.stream.render_sequence)."""
    launch_conditioning(model, 0, auds, eye, poses)


def head_desc(model, st, rays_o, rays_d, noises, dt_gamma, max_steps, T_thresh):
    N, dev = rays_o.shape[0], rays_o.device
    weights_sum = torch.empty(N, device=dev)
    depth = torch.empty(N, device=dev)
    image = torch.empty(N, 3, device=dev)
    nears = torch.empty(N, device=dev)
    fars = torch.empty(N, device=dev)
    hd = FrameHeadDesc()
    hd.N, hd.max_steps, hd.cascade, hd.grid_size = N, int(max_steps), int(model.cascade), int(model.grid_size)
    hd.bound, hd.min_near, hd.dt_gamma, hd.T_thresh = float(model.bound), float(model.min_near), float(dt_gamma), float(T_thresh)
    hd.rays_o, hd.rays_d = rays_o.data_ptr(), rays_d.data_ptr()
    hd.aabb, hd.bitfield, hd.noises = model.aabb_infer.data_ptr(), model.density_bitfield.data_ptr(), _p(noises)
    hd.weights_sum, hd.depth, hd.image = weights_sum.data_ptr(), depth.data_ptr(), image.data_ptr()
    hd.nears, hd.fars = nears.data_ptr(), fars.data_ptr()
    hd.workspace, hd.workspace_bytes = st.workspace.data_ptr(), st.ws_bytes
    hd.grid3d, hd.grid2d = _grid_table(model.encoder, st.table3), _grid_table(model.encoder_ambient, st.table2)
    hd.head_blob, hd.head_consts = st.head_blob.data_ptr(), st.head_consts.data_ptr()
    hd.capture_unroll = st.capture_unroll
    hd.occ_aabb = st.occupied_box(model).data_ptr()
    if st.shared.occ_words and _USE_OCC_PACK:
        hd.occ_pack, hd.occ_words = st.shared.occ_pack.data_ptr(), st.shared.occ_words
    return hd, (weights_sum, depth, image, nears, fars)


def _launch(model, st, rays_o, rays_d, auds, bg_coords, pose6, eye_t, bg_t, bg_scalar, noises, dt_gamma, max_steps, T_thresh,
            external_cond=False):
    """the frame's launch sequence (capturable: no host sync, outputs allocated with torch.empty)"""
    L = abi.lib()
    N = rays_o.shape[0]
    dev = rays_o.device
    stream = abi.cur_stream()

    # ---- forked stream: per-frame conditioning (audio nets), then the whole torso branch.  The audio nets overlap ray setup
    #      + the first march (the head waits for `cond_event` only before its first network evaluation); the torso branch
    #      only meets the head in the final blend, so it fills the SMs the march / composite phases of the head loop leave
    #      idle.  Works the same under CUDA-graph capture (fork / join edges).
    #      With external_cond the conditioning kernel already ran (launch_conditioning, frame pipelining) and only the
    #      torso branch is forked.
    cur = torch.cuda.current_stream()
    results = {}
    torso_bg = torch.empty(N, 3, device=dev) if model.torso else None
    st.side.wait_stream(cur)
    with torch.cuda.stream(st.side):
        if not external_cond:
            cd = conditioning_desc(model, st, auds, eye_t, pose6)
            abi.check(L.rn_frame_conditioning(C.byref(cd), abi.cur_stream()))
            st.cond_event.record(st.side)
        if model.torso:
            td = FrameTorsoDesc()
            td.N, td.grid_size = N, int(model.grid_size)
            td.thresh, td.shrink = float(min(model.density_thresh_torso, model.mean_density_torso)), float(model.opt.torso_shrink)
            td.bg_coords, td.density_grid_torso = bg_coords.data_ptr(), model.density_grid_torso.data_ptr()
            td.workspace, td.workspace_bytes = st.workspace.data_ptr(), st.ws_bytes
            td.grid2d = _grid_table(model.torso_encoder, st.table_t)
            td.torso_blob, td.torso_consts = st.torso_blob.data_ptr(), st.torso_consts.data_ptr()
            td.torso_alpha, td.torso_color = st.torso_alpha.data_ptr(), st.torso_color.data_ptr()
            abi.check(L.rn_frame_torso(C.byref(td), abi.cur_stream()))
            results['torso_alpha'] = st.torso_alpha
            results['torso_color'] = torso_bg

    # ---- head
    hd, (weights_sum, depth, image, nears, fars) = head_desc(model, st, rays_o, rays_d, noises, dt_gamma, max_steps, T_thresh)
    hd.consts_ready_event = None if external_cond else st.cond_event.cuda_event
    abi.check(L.rn_frame_head(C.byref(hd), stream))

    cur.wait_stream(st.side)  # join: the final blend needs the torso
    abi.check(L.rn_frame_finalize(N, weights_sum.data_ptr(), depth.data_ptr(), image.data_ptr(), nears.data_ptr(), fars.data_ptr(),
                                  _p(bg_t), bg_scalar, _p(st.torso_alpha) if model.torso else None,
                                  _p(st.torso_color) if model.torso else None, _p(torso_bg), stream))
    results['depth'] = depth
    results['image'] = image
    results['weights_sum'] = weights_sum
    return results


def render_frame(model, rays_o, rays_d, auds, bg_coords, poses, eye=None, index=0, dt_gamma=0, bg_color=None, perturb=False,
                 force_all_rays=False, max_steps=1024, T_thresh=1e-4, lane=0, external_cond=False, **kwargs):
    """Fused inference frame.  With `model._fused.use_graph` (default) the launch sequence is captured once per
    configuration into a CUDA graph and replayed: inputs are copied into the graph's static buffers and the returned
    tensors are the graph's static OUTPUT buffers -- they are overwritten by the next call (consume or clone them first)."""
    if model.training:
        raise RuntimeError("render_frame is the inference path; training goes through run_cuda")
    if not supported(model):
        raise NotImplementedError("model configuration outside the fused kernels' specialisation")
    st = lane_state(model, lane)
    st.refresh_weights(model)
    st.occupied_box(model)   # refreshed here, outside any capture, when the bitfield changed

    prefix = rays_o.shape[:-1]
    rays_o = rays_o.contiguous().view(-1, 3).float()
    rays_d = rays_d.contiguous().view(-1, 3).float()
    bg_coords = bg_coords.contiguous().view(-1, 2).float()
    N = rays_o.shape[0]
    dev = rays_o.device
    st.ensure_workspace(N)

    if not external_cond:
        _sync_smoothing_state(model, st)
    auds_t = None if auds is None else auds.contiguous().float()
    eye_t = None if eye is None else eye.reshape(-1).float().contiguous()
    pose6 = poses.reshape(-1).float().contiguous() if model.torso else None
    bg_t, bg_scalar = None, 1.0
    if bg_color is not None:
        if torch.is_tensor(bg_color):
            bg_t = bg_color.reshape(-1, 3).float()
            bg_t = (bg_t.expand(N, 3) if bg_t.shape[0] != N else bg_t).contiguous()
        else:
            bg_scalar = float(bg_color)

    if st.use_graph and not perturb:
        key = (N, None if auds_t is None else tuple(auds_t.shape), eye_t is not None, bg_t is not None, bg_scalar, float(dt_gamma),
               int(max_steps), float(T_thresh), float(model.mean_density_torso), model.density_bitfield.data_ptr(), bool(external_cond),
               int(st.shared.occ_words))
        entry = st.graphs.get(key)
        if entry is None:
            # small per-frame inputs live in ONE block [pose 4x4 | pose6 | eye | pad | auds] so that a streaming caller
            # (radnerf_b200.stream.FrameStreamer) fills them with a single host->device copy; `pose` is only used by callers
            # that generate the rays on the device
            n_aud = 0 if auds_t is None else auds_t.numel()
            flat = torch.zeros(24 + n_aud, device=dev)
            static = dict(rays_o=torch.empty_like(rays_o), rays_d=torch.empty_like(rays_d), bg_coords=torch.empty_like(bg_coords),
                          flat=flat, pose=flat[:16].view(4, 4), auds=None if auds_t is None else flat[24:24 + n_aud].view(auds_t.shape),
                          eye=None if eye_t is None else flat[22:23], pose6=None if pose6 is None else flat[16:22],
                          bg=None if bg_t is None else torch.empty_like(bg_t))
            for k, v in (("rays_o", rays_o), ("rays_d", rays_d), ("bg_coords", bg_coords), ("auds", auds_t), ("eye", eye_t),
                         ("pose6", pose6), ("bg", bg_t)):
                if v is not None and static[k].data_ptr() != v.data_ptr():
                    static[k].copy_(v)
            # warm-up outside capture (lazy kernel attributes, module loading) on a scratch copy of the smoothing state
            saved = st.enc_a_state.clone()
            _launch(model, st, static["rays_o"], static["rays_d"], static["auds"], static["bg_coords"], static["pose6"], static["eye"],
                    static["bg"], bg_scalar, None, dt_gamma, max_steps, T_thresh, external_cond)
            st.enc_a_state.copy_(saved)
            torch.cuda.synchronize()
            # the warm-up frame tells how many loop iterations this kind of frame needs: capture that many as plain kernel
            # nodes, the (normally idle) remainder of the loop as one conditional WHILE node
            st.capture_unroll = max(1, len(frame_stats(model, st)))
            graph = torch.cuda.CUDAGraph()
            k0 = abi.launch_count()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):  # rn_frame_head captures the loop body on a helper stream
                outs = _launch(model, st, static["rays_o"], static["rays_d"], static["auds"], static["bg_coords"], static["pose6"],
                               static["eye"], static["bg"], bg_scalar, None, dt_gamma, max_steps, T_thresh, external_cond)
            st.enc_a_state.copy_(saved)  # capture does not execute, but keep the invariant explicit
            entry = st.graphs[key] = (graph, static, outs, abi.launch_count() - k0)
        graph, static, outs, n_kernels = entry
        st.last_static = static
        st.last_entry = entry
        # a caller that already wrote into the graph's input buffers (FrameStreamer) passes those very tensors: no copies
        for k, v in (("rays_o", rays_o), ("rays_d", rays_d), ("bg_coords", bg_coords), ("auds", auds_t), ("eye", eye_t),
                     ("pose6", pose6), ("bg", bg_t)):
            if v is not None and static[k].data_ptr() != v.data_ptr():
                static[k].copy_(v)
        graph.replay()
        abi.lib().rn_note_graph_replay(n_kernels)
        results = dict(outs)
    else:
        noises = torch.rand(N, device=dev) if perturb else None
        results = _launch(model, st, rays_o, rays_d, auds_t, bg_coords, pose6, eye_t, bg_t, bg_scalar, noises, dt_gamma, max_steps,
                          T_thresh, external_cond)
    if model.smooth_lips and auds is not None and not external_cond:
        model.enc_a = st.enc_a_state[:64].view(1, 64)
    st.frames += 1
    results['depth'] = results['depth'].view(*prefix)
    results['image'] = results['image'].view(*prefix, 3)
    return results


def frame_stats(model, st=None):
    """(n_alive, n_step, n_samples) per executed iteration of the last fused frame (of lane state `st`, default lane 0) --
    forces a sync; for tests/bench only"""
    ctl = (st if st is not None else model._fused).ctl().cpu().numpy()
    out = []
    for it in range(64):
        n_alive, n_step, step, done, n_samples = [int(v) for v in ctl[it][:5]]
        if done or (it > 0 and n_alive == 0):
            break
        out.append((n_alive, n_step, n_samples))
    return out


abi.register("rn_frame_head_timed", [C.POINTER(FrameHeadDesc), _vp, _vp, _vp])

# algorithmic bytes per sample of the fused head kernel: 3-D fp16 gathers + 2-D fp16 gathers (the feature
# writes/re-reads of the unfused formulation, 64 B each, never happen) + the sample record in and the evaluation out
HEAD_BYTES_PER_SAMPLE = (4 * 3 + 16 * 8 * 2 * 2) + (16 * 4 * 2 * 2) + 16 + 16
HEAD_FLOP_PER_SAMPLE = 47872  # hoisted MLP chain


def roofline_entries(model, f, bg_local, kw, hbm, tflops, reps=10):
    """per-kernel device times of the head loop measured with CUDA events on the launching stream (rn_frame_head_timed)"""
    L = abi.lib()
    # make sure the fused state / conditioning vectors exist
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        model.render(f["ro"], f["rd"], f["auds"], bg_local, f["pose6"], eye=f["eye"], index=0, path="fused", **kw)
    st = model._fused
    rays_o, rays_d = f["ro"][0].contiguous(), f["rd"][0].contiguous()
    hd, keep = head_desc(model, st, rays_o, rays_d, None, kw["dt_gamma"], kw["max_steps"], 1e-4)
    ms = (C.c_float * (3 * kw["max_steps"]))()
    ns = (C.c_uint32 * kw["max_steps"])()
    tot = np.zeros(3)
    samples = 0
    evals = []
    for r in range(reps + 2):
        abi.check(L.rn_frame_head_timed(C.byref(hd), abi.cur_stream(), ms, ns))
        if r < 2:
            continue
        for it in range(kw["max_steps"]):
            if ns[it] == 0 and it > 0:
                break
            tot += np.array([ms[3 * it], ms[3 * it + 1], ms[3 * it + 2]])
            samples += ns[it]
            evals.append((ns[it], ms[3 * it + 1]))
    n_launch = len(evals)
    ev_ms, ev_samples = sum(e[1] for e in evals), sum(e[0] for e in evals)
    N = rays_o.shape[0]
    head = {"kernel": "head_eval_kernel (fused 3-D encode + ambient MLP + 2-D encode + sigma MLP + SH + colour MLP, tcgen05)",
            "bound": "hbm", "units": ev_samples / n_launch, "unit_name": "samples", "bytes_per_unit": HEAD_BYTES_PER_SAMPLE,
            "ms": ev_ms / n_launch, "achieved": ev_samples * HEAD_BYTES_PER_SAMPLE / ev_ms / 1e6, "peak": hbm, "unit": "GB/s",
            "frac": ev_samples * HEAD_BYTES_PER_SAMPLE / ev_ms / 1e6 / hbm, "gunits_per_s": ev_samples / ev_ms / 1e6,
            "launches_per_frame": n_launch / reps, "ms_per_frame": ev_ms / reps,
            "tensor_tflops": ev_samples * HEAD_FLOP_PER_SAMPLE / ev_ms / 1e9,
            "tensor_frac_of_peak": ev_samples * HEAD_FLOP_PER_SAMPLE / ev_ms / 1e9 / tflops}
    # DRAM bytes / instructions per sample from the newest committed `ncu --set full` capture of this kernel (profiles/rNN_head_eval_ncu_full.json)
    try:
        import glob, json, os
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        prof = sorted(glob.glob(os.path.join(root, "profiles", "r*_head_eval_ncu_full.json")))[-1]
        per = json.load(open(prof))["per_sample"]
        head["traffic"] = per["dram_bytes"] * head["units"]
        head["traffic_note"] = ("dram__bytes_read+write per sample (%.1f B, profiles/%s) x samples per launch; far below the algorithmic "
                                "812 B/sample because both hash tables (5.8 MB) stay L2-resident: the gathers are served by L1/L2"
                                % (per["dram_bytes"], os.path.basename(prof)))
        # what really bounds the kernel: instruction issue.  ceiling = SMs x 4 schedulers x 32 lanes x SM clock / thread-instructions per sample
        clock_ghz = float(torch.cuda.get_device_properties(rays_o.device).clock_rate) / 1e6 if hasattr(torch.cuda.get_device_properties(rays_o.device), "clock_rate") else 1.965
        sms = torch.cuda.get_device_properties(rays_o.device).multi_processor_count
        ceiling = sms * 4 * 32 * clock_ghz / per["thread_instructions"]
        head["issue"] = {"bound": "issue", "inst_per_sample": per["thread_instructions"], "ceiling_gsamples_per_s": ceiling,
                         "achieved_gsamples_per_s": head["gunits_per_s"], "frac": head["gunits_per_s"] / ceiling, "sm_clock_ghz": clock_ghz,
                         "l2_bytes_per_sample": per.get("l2_bytes") or None, "l1_hit_pct": per.get("l1_hit_pct"),
                         "note": "the HBM figures above are SURVEY 8(d)'s algorithmic-bytes accounting; DRAM sees %.0f B/sample, the kernel is bound by "
                                 "instruction issue + gather latency (profiles/%s)" % (per["dram_bytes"], os.path.basename(prof))}
    except Exception:
        head["traffic"] = None
    march = {"kernel": "march_compact_kernel (occupancy DDA + sample compaction)", "bound": "hbm", "ms_per_frame": tot[0] / reps,
             "bytes_per_frame": 44.0 * N + 32.0 * samples / reps, "ms": tot[0] / n_launch}
    march.update(achieved=march["bytes_per_frame"] / march["ms_per_frame"] / 1e6, peak=hbm, unit="GB/s")
    march["frac"] = march["achieved"] / hbm
    comp = {"kernel": "composite_compact_kernel (compositing + survivor compaction + loop control)", "bound": "hbm",
            "ms_per_frame": tot[2] / reps, "bytes_per_frame": 24.0 * samples / reps + 56.0 * N, "ms": tot[2] / n_launch}
    comp.update(achieved=comp["bytes_per_frame"] / comp["ms_per_frame"] / 1e6, peak=hbm, unit="GB/s")
    comp["frac"] = comp["achieved"] / hbm
    return [head, march, comp]
