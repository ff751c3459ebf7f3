"""Host-facing frame pipeline of the fused renderer: the loop the reference runs in Trainer.test (nerf/utils.py:905-960) --
per frame: pose / audio window / eye value arrive on the host, the frame is rendered, `preds.detach().cpu()` brings the
image back -- restated as a pipeline with several frames in flight:

    pinned host block --(ONE H2D copy)--> lane's graph input block --> rays on device --> conditioning kernel (frame order,
        own stream) --> lane's frame graph --> [peer stores / NCCL all-gather] --> staging slot [fp32 or uint8]
        --(copy stream, D2H)--> pinned host image slot

* FramePipeline   K frame "lanes" on K streams sharing one set of weights; keeps the lip-smoothing chain exact.
* FrameStreamer   drives the lanes from host inputs.  The first frame of a lane goes through Python (it captures the lane's
                  CUDA graph); from then on a frame is ONE C call (rn_lane_submit_frame, csrc/pipeline.cu) with raw stream /
                  event / graph handles, because issuing it from Python cost more host time than a ray-sharded frame needs
                  on 4-8 GPUs.  `collect()` is the only blocking call.
* render_sequence frame-parallel rendering of a known sequence over ranks (whole frames per GPU, no collective).

Staging and host slots belong to a lane and are reused only after the copy that read them has completed (event-ordered, no
host sync on the submit path)."""
import ctypes as C
from collections import deque

import numpy as np
import torch

from . import abi
from .rays import RayGenerator
from .sharding import FrameSharder


_vp, _u32, _u64, _f32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float


class LaneSubmit(C.Structure):   # rn_lane_submit (include/radnerf_b200.h)
    _fields_ = [("lane_stream", _vp), ("cond_stream", _vp), ("copy_stream", _vp),
                ("ev_in", _vp), ("ev_cond", _vp), ("ev_done", _vp), ("ev_staged", _vp), ("ev_delivered", _vp),
                ("packed_src", _vp), ("flat_dst", _vp), ("packed_bytes", _u64),
                ("pose", _vp), ("fx", _f32), ("fy", _f32), ("cx", _f32), ("cy", _f32), ("H", _u32), ("W", _u32),
                ("pixel_ids", _vp), ("n_rays", _u32), ("graph_kernels", _u32), ("rays_o", _vp), ("rays_d", _vp),
                ("cond", _vp), ("graph_exec", _vp),
                ("image_local", _vp), ("ids", _vp), ("peers", _vp),
                ("n_local", _u32), ("run_pixels", _u32), ("world", _u32), ("phase", _u32),
                ("stage_src", _vp), ("stage_dst", _vp), ("host_dst", _vp), ("image_bytes", _u64),
                ("to_uint8", _u32), ("reserved", _u32),
                ("ctrl_peers", _vp), ("ticket", _vp), ("frame_seq", _u64),
                ("rank", _u32), ("root", _u32), ("slot", _u32), ("scatter_ctas", _u32)]


abi.register("rn_lane_submit_frame", [C.POINTER(LaneSubmit)])
abi.register("rn_event_create", [C.POINTER(_vp)])
abi.register("rn_event_destroy", [_vp])
abi.register("rn_event_synchronize", [_vp])
abi.register("rn_stream_wait_event", [_vp, _vp])


def pack_inputs(pose, auds, pose6=None, eye=None):
    """one pinned fp32 block per frame, laid out like the frame graph's input block: [pose 4x4 | pose6 | eye | pad | auds]"""
    auds = np.asarray(auds, dtype=np.float32)
    buf = torch.zeros(24 + auds.size, dtype=torch.float32).pin_memory()
    b = buf.numpy()
    b[:16] = np.asarray(pose, dtype=np.float32).reshape(-1)
    if pose6 is not None:
        b[16:22] = np.asarray(pose6, dtype=np.float32).reshape(-1)
    if eye is not None:
        b[22] = float(np.asarray(eye, dtype=np.float32).reshape(-1)[0])
    b[24:] = auds.reshape(-1)
    return buf


class FramePipeline:
    """Several frames in flight on ONE GPU.  A frame is a dependency chain (5 x march -> network -> composite, each step
    a wave or less on 148 SMs for most of its duration), so one frame at a time leaves the GPU idle in its latency-bound
    phases; with 2-3 frames on separate streams ("lanes") the network kernel of one frame runs while another frame marches
    or composites.  Measured at 512x512 on one B200: 0.449 -> 0.355 -> 0.343 ms/frame for 1 -> 2 -> 3 lanes.

    Frames stay coupled only through the lip-smoothing EMA: the conditioning kernels (audio nets + smoothing + hoisted terms)
    therefore run in frame order on one dedicated stream, each writing its lane's hoisted-term vectors; the frame itself
    (render_frame(..., lane=k, external_cond=True)) runs on lane k's stream once its conditioning is done."""

    def __init__(self, model, lanes=2):
        self.model, self.lanes = model, lanes
        dev = model.density_bitfield.device
        self.cond_stream = torch.cuda.Stream(device=dev)
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(lanes)]
        self.ev_cond = [torch.cuda.Event() for _ in range(lanes)]
        self.ev_done = [torch.cuda.Event() for _ in range(lanes)]
        self.n = 0

    def next_lane(self):
        return self.n % self.lanes

    def submit(self, rays_o, rays_d, auds, bg_coords, poses, eye=None, post=None, ready_on=None, static_inputs=False, pose44=None, **kw):
        """Enqueue one frame.  The inputs are device tensors that are ready on stream `ready_on` (default: the current
        stream) and stay untouched until the frame has run.  post(out, lane), if given, runs on the lane's stream right
        after the frame (gather, staging, ...).  static_inputs=True: the tensors ARE the input buffers of the lane's captured
        graph (FusedState.last_static), so the graph is replayed without going through render_frame.  Returns (lane, post's
        result or the frame's result dict); the tensors are valid once wait(lane) has been ordered, and until the lane is
        used again."""
        from . import frame as _frame
        k = self.n % self.lanes
        self.n += 1
        src = ready_on if ready_on is not None else torch.cuda.current_stream()
        ls = self.streams[k]
        self.cond_stream.wait_stream(src)
        self.cond_stream.wait_event(self.ev_done[k])   # lane k's previous frame no longer reads its hoisted-term vectors
        with torch.cuda.stream(self.cond_stream):
            _frame.launch_conditioning(self.model, k, auds, eye, poses, pose44=pose44)
            self.ev_cond[k].record(self.cond_stream)
        if src is not ls:
            ls.wait_stream(src)
        ls.wait_event(self.ev_cond[k])
        with torch.cuda.stream(ls), torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=self.model.opt.fp16):
            out = _frame.replay_lane(self.model, k) if static_inputs else None
            if out is None:
                out = _frame.render_frame(self.model, rays_o, rays_d, auds, bg_coords, poses, eye=eye, lane=k, external_cond=True, **kw)
            res = post(out, k) if post is not None else out
            self.ev_done[k].record(ls)
        return k, res

    def wait(self, lane):
        """order the current stream after lane `lane`'s last frame"""
        torch.cuda.current_stream().wait_event(self.ev_done[lane])

    def sync(self):
        for k in range(self.lanes):
            self.wait(k)


class FrameStreamer:
    def __init__(self, model, H, W, intrinsics, bg_coords, auds_shape, use_eye=True, sharder=None, deliver=True, depth=2,
                 output="float32", device_pose6=True, **render_kw):
        """bg_coords: [H*W, 2] on the device (this rank's rows if `sharder` splits the frame); auds_shape: e.g. (8, 44, 16);
        deliver=False skips the device->host stage (ranks other than the one that consumes the frames); depth = frames in
        flight (= lanes of the FramePipeline underneath); output="uint8" converts on the device -- the reference's
        `(pred * 255).astype(np.uint8)` -- and copies a quarter of the bytes to the host."""
        assert output in ("float32", "uint8") and (output == "float32" or (H * W * 3) % 16 == 0)
        self.u8 = output == "uint8"
        # device_pose6: the torso's 6-vector pose (the reference's convert_poses, nerf/utils.py:230-237) is derived from the 4x4
        # pose inside the conditioning kernel; the [16:22] slot of the input block is then an OUTPUT (host values are ignored)
        self.device_pose6 = bool(device_pose6)
        self.model, self.kw, self.depth, self.deliver = model, render_kw, depth, deliver
        self.dev = bg_coords.device
        self.sharder = sharder if sharder is not None else FrameSharder(H, W, 1, 0, self.dev)
        self.raygen = RayGenerator(H, W, intrinsics, self.dev, self.sharder)
        self.bg = self.sharder.shard(bg_coords) if bg_coords.shape[0] == H * W and self.sharder.world > 1 else bg_coords
        self.auds_shape, self.use_eye = tuple(auds_shape), use_eye
        self.n_in = 24 + int(np.prod(auds_shape))
        self.pipe = FramePipeline(model, lanes=depth)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        odt = torch.uint8 if self.u8 else torch.float32
        self.dev_stage = [torch.empty(H * W, 3, device=self.dev, dtype=odt) for _ in range(depth)]
        self.host_out = [torch.empty(H * W, 3, dtype=odt).pin_memory() for _ in range(depth)]
        self.staged = [torch.cuda.Event() for _ in range(depth)]
        self.delivered = [torch.cuda.Event() for _ in range(depth)]
        self.pending = deque()
        self.ring_read = [torch.cuda.Event() for _ in range(depth)]
        self.static = [None] * depth
        self.fast = [None] * depth      # per lane: (LaneSubmit, things it points to) once the lane's graph exists
        self.fast_generation = -1
        self.H, self.W = H, W
        self.h2d_bytes = 4 * self.n_in
        self.d2h_bytes = (3 if self.u8 else 12) * H * W

    def _views(self, flat):
        auds = flat[24:self.n_in].view(self.auds_shape)
        eye = flat[22:23] if self.use_eye else None
        return auds, flat[16:22], eye

    def _post(self, out, k):
        """on lane k's stream, right after the frame: assemble the image, stage it, start the copy-out"""
        if self.sharder.world > 1 and self.sharder.ctrl is not None:
            # flag-based gather-to-root: scatter + (on the root) wait / stage / release in two launches, no barrier
            ls = torch.cuda.current_stream(self.dev)
            if self.deliver:
                ls.wait_event(self.delivered[k])      # the copy that last read this staging slot has drained
            img = self.sharder.gather_to_root(out["image"].view(-1, 3), slot=k, stage_to=self.dev_stage[k] if self.deliver else None,
                                              to_uint8=self.u8)
            if self.deliver:
                self.staged[k].record(ls)
                with torch.cuda.stream(self.copy_stream):
                    self.copy_stream.wait_event(self.staged[k])
                    self.host_out[k].copy_(self.dev_stage[k], non_blocking=True)
                    self.delivered[k].record(self.copy_stream)
            return img
        img = self.sharder.gather(out["image"].view(-1, 3), slot=k)
        if self.deliver:
            ls = torch.cuda.current_stream(self.dev)
            ls.wait_event(self.delivered[k])          # the copy that last read this staging slot has drained
            if self.u8:
                abi.check(abi.lib().rn_image_to_uint8(abi.ptr(img), abi.ptr(self.dev_stage[k]), img.numel(), abi.cur_stream()))
            else:
                self.dev_stage[k].copy_(img)
            self.staged[k].record(ls)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self.staged[k])
                self.host_out[k].copy_(self.dev_stage[k], non_blocking=True)
                self.delivered[k].record(self.copy_stream)
        return img

    def submit(self, packed, ring=None):
        """packed: pinned host block from pack_inputs() (or the same block already on the device).  Enqueues copy-in, ray
        generation, conditioning, the frame and copy-out; never blocks.

        ring: a radnerf_b200.audio_ring.FeatureRing (streaming mode, the reference's `--asr` GUI loop, nerf/gui.py:180-187).
        `packed` then holds only the 24-float head [pose | pose6 | eye | pad]; the frame's audio window never exists on the
        host: the ring's gather kernel writes it straight into the lane's input block."""
        from . import frame as _frame
        k = self.pipe.next_lane()
        if ring is not None:
            assert packed.numel() == 24 and (packed.is_cuda or packed.is_pinned()) and ring.dim * 128 == self.n_in - 24
            cur = torch.cuda.current_stream(self.dev)
            ls = self.pipe.streams[k]
            st = self.static[k]
            armed = self.fast[k] is not None and st is not None
            full = None if armed else torch.empty(self.n_in, dtype=torch.float32, device=self.dev)
            if full is not None:
                full.record_stream(ls)
            ls.wait_stream(cur)                        # the ring rows pushed so far on the caller's stream are in place
            with torch.cuda.stream(ls):                # ... and the lane's previous frame has finished reading its block
                ring.next_window(out=st["flat"][24:self.n_in] if armed else full[24:])
                self.ring_read[k].record(ls)
            cur.wait_event(self.ring_read[k])          # a later push must not overwrite rows this gather still reads
            if armed and self._submit_fast(k, packed, head_only=True):
                return
            if armed:                                  # the fast path was disarmed under us (weights / graph changed)
                full = st["flat"][:self.n_in].clone()
            full[:24].copy_(packed, non_blocking=True)
            packed = full                              # a complete device block: carry on through the Python path
        else:
            assert packed.numel() == self.n_in and (packed.is_cuda or packed.is_pinned())
            if self._submit_fast(k, packed):
                return
        kw = dict(index=0, bg_color=None, perturb=False, **self.kw)
        st = self.static[k]
        if st is None or _frame.lane_state(self.model, k).last_static is not st:
            # first frame of this lane (or its graph was rebuilt): an ordinary call creates the graph and its input buffers
            flat = packed.to(self.dev, non_blocking=True)
            ro, rd = self.raygen(flat[:16].view(4, 4))
            auds, pose6, eye = self._views(flat)
            self.pipe.submit(ro[None], rd[None], auds, self.bg[None], pose6, eye=eye, post=self._post,
                             pose44=flat[:16] if self.device_pose6 else None, **kw)
            self.static[k] = _frame.lane_state(self.model, k).last_static
            if self.static[k] is not None and k == 0:
                self.bg = self.static[0]["bg_coords"]   # same values, already in place for lane 0: no per-frame copy there
        else:
            ls = self.pipe.streams[k]
            with torch.cuda.stream(ls):   # ordered after this lane's previous frame, which read these buffers
                st["flat"].copy_(packed, non_blocking=True)                      # ONE host->device copy per frame
                self.raygen(st["pose"], out=(st["rays_o"], st["rays_d"]))        # rays straight into the graph's inputs
            auds, pose6, eye = self._views(st["flat"])
            self.pipe.submit(st["rays_o"][None], st["rays_d"][None], auds, st["bg_coords"][None], pose6, eye=eye, post=self._post,
                             ready_on=ls, static_inputs=True, pose44=st["pose"] if self.device_pose6 else None, **kw)
        self.pending.append((k, False))
        self._arm_fast(k)

    # ---- steady state: one C call per frame (csrc/pipeline.cu) ---------------------------------------------------------
    def _arm_fast(self, k):
        """after a frame went through the Python path on lane k, everything the lane needs is fixed: describe it once"""
        from . import frame as _frame
        st = _frame.lane_state(self.model, k)
        entry = st.__dict__.get("last_entry")
        static = self.static[k]
        peer = self.sharder.peer if self.sharder.world > 1 else None
        if entry is None or static is None or (self.sharder.world > 1 and peer is None) or not hasattr(entry[0], "raw_cuda_graph_exec"):
            return   # no graph (perturbed frames), or the NCCL gather: stay on the Python path
        graph, _, outs, n_kernels = entry
        torch.cuda.synchronize(self.dev)   # the Python-path frame used torch events; start the C-event bookkeeping from idle
        L = abi.lib()
        ev = []
        for _ in range(5):
            e = _vp()
            abi.check(L.rn_event_create(C.byref(e)))
            ev.append(e.value)
        auds, pose6, eye = self._views(static["flat"])
        cd = _frame.conditioning_desc(self.model, st, auds.contiguous(), None if eye is None else eye.contiguous(),
                                      pose6.contiguous() if self.model.torso else None,
                                      static["pose"].reshape(-1) if (self.model.torso and self.device_pose6) else None)
        s = LaneSubmit()
        s.lane_stream, s.cond_stream, s.copy_stream = self.pipe.streams[k].cuda_stream, self.pipe.cond_stream.cuda_stream, self.copy_stream.cuda_stream
        s.ev_in, s.ev_cond, s.ev_done, s.ev_staged, s.ev_delivered = ev
        s.flat_dst, s.packed_bytes = static["flat"].data_ptr(), 4 * self.n_in
        g = self.raygen
        s.pose, s.fx, s.fy, s.cx, s.cy, s.H, s.W = static["pose"].data_ptr(), g.fx, g.fy, g.cx, g.cy, g.H, g.W
        s.pixel_ids, s.n_rays = (None if g.ids is None else g.ids.data_ptr()), g.n
        s.rays_o, s.rays_d = static["rays_o"].data_ptr(), static["rays_d"].data_ptr()
        s.cond = C.cast(C.pointer(cd), _vp)
        s.graph_exec, s.graph_kernels = graph.raw_cuda_graph_exec(), n_kernels
        image = outs["image"]
        s.image_local = image.data_ptr()
        if peer is not None:
            bufs, hdls, ids32 = peer
            s.ids, s.peers, s.n_local, s.run_pixels, s.world = ids32.data_ptr(), hdls[k % len(bufs)].buffer_ptrs_dev, image.shape[0], self.W, self.sharder.world
            s.stage_src = bufs[k % len(bufs)].data_ptr()
            if self.sharder.ctrl is not None:
                ctrl, ctrl_hdl, tickets, _, ctas = self.sharder.ctrl
                s.ctrl_peers, s.scatter_ctas = ctrl_hdl.buffer_ptrs_dev, ctas
                s.rank, s.root, s.slot = self.sharder.rank, self.sharder.root, k % len(bufs)
                s.ticket = tickets[s.slot:s.slot + 1].data_ptr()
        else:
            s.stage_src = image.data_ptr()
        s.stage_dst, s.host_dst = self.dev_stage[k].data_ptr(), (self.host_out[k].data_ptr() if self.deliver else None)
        s.image_bytes = 12 * self.H * self.W
        s.to_uint8 = 1 if self.u8 else 0
        self.fast[k] = (s, cd, ev, entry, hdls[k % len(bufs)] if peer is not None else None)
        self.fast_generation = st.shared.generation

    def _submit_fast(self, k, packed, head_only=False):
        f = self.fast[k]
        if f is None:
            return False
        from . import frame as _frame
        st = _frame.lane_state(self.model, k)
        m = self.model
        st.shared.refresh(m)
        st.shared.occupied_box(m)
        ea = getattr(m, "enc_a", None)
        if (st.shared.generation != self.fast_generation or st.__dict__.get("last_entry") is not f[3]
                or (m.smooth_lips and (ea is None or ea.data_ptr() != st.enc_a_state.data_ptr()))):
            self.fast = [None] * self.depth   # weights / graph / smoothing state changed under us: back to the Python path
            torch.cuda.synchronize(self.dev)
            return False
        s, barrier = f[0], f[4]
        s.packed_src = packed.data_ptr()
        s.packed_bytes = 4 * (24 if head_only else self.n_in)   # streaming: the audio part is already in the lane's block
        self.__dict__.setdefault("_keep", {})[k] = packed   # the async copy-in reads it: alive until the lane's next frame
        L = abi.lib()
        if barrier is None or s.ctrl_peers:
            if s.ctrl_peers:
                s.frame_seq = self.sharder.next_seq(s.slot)
            s.phase = 3
            abi.check(L.rn_lane_submit_frame(C.byref(s)))
        else:
            s.phase = 1
            abi.check(L.rn_lane_submit_frame(C.byref(s)))
            with torch.cuda.stream(self.pipe.streams[k]):
                barrier.barrier(channel=0)   # every rank's rows have landed in every rank's frame buffer of this lane
            s.phase = 2
            abi.check(L.rn_lane_submit_frame(C.byref(s)))
        self.pipe.n += 1
        self.pending.append((k, True))
        return True

    def in_flight(self):
        return len(self.pending)

    def sync(self):
        """order the CURRENT stream after every frame submitted so far (device-side wait, no host block)"""
        self.pipe.sync()
        cur = abi.cur_stream()
        for f in self.fast:
            if f is not None:
                abi.check(abi.lib().rn_stream_wait_event(cur, f[2][2]))   # the lane's ev_done

    def collect(self):
        """blocks until the oldest submitted frame is on the host; returns the pinned [H*W, 3] image (valid until `depth`
        more frames have been submitted)"""
        slot, fast = self.pending.popleft()
        if not self.deliver:   # nothing to hand over: just wait for the frame itself (keeps the host `depth` frames ahead at most)
            if fast and self.fast[slot] is not None:
                abi.check(abi.lib().rn_event_synchronize(self.fast[slot][2][2]))
            elif not fast:
                self.pipe.ev_done[slot].synchronize()
            return None
        if fast:
            abi.check(abi.lib().rn_event_synchronize(self.fast[slot][2][4]) if self.fast[slot] is not None else 0)
        else:
            self.delivered[slot].synchronize()
        return self.host_out[slot]

    def close(self):
        """wait for everything in flight and release the C-side events"""
        torch.cuda.synchronize(self.dev)
        for f in self.fast:
            if f is not None:
                for e in f[2]:
                    abi.lib().rn_event_destroy(e)
        self.fast = [None] * self.depth
        self.pending.clear()

    def render_all(self, packed_frames):
        """generator over host images, keeping `depth` frames in flight"""
        for p in packed_frames:
            if len(self.pending) == self.depth:
                yield self.collect()
            self.submit(p)
        while self.pending:
            yield self.collect()


# ------------------------------------------------------------------------------------------------ sequences on N GPUs
SMOOTHING_MEMORY = 24   # frames after which the lip-smoothing EMA (lambda = 0.35) has forgotten its start: 0.35**24 ~ 1e-11


def sequence_slice(n_frames, world, rank):
    """contiguous slice [lo, hi) of a sequence owned by `rank` (the first n_frames % world ranks get one frame more)"""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def render_sequence(model, packed_frames, H, W, intrinsics, bg_coords, auds_shape, world=1, rank=0, use_eye=True, **render_kw):
    """Frame-parallel rendering of a known sequence (the offline `test.py` case): rank r renders the contiguous slice
    sequence_slice(len(frames), world, r) of WHOLE frames with a replicated model -- no collective, no ray sharding, so the
    throughput of N GPUs is N times one GPU's (a 512x512 frame split by rays stops scaling at ~0.29 ms of per-frame latency,
    see DESIGN.md 6).  Frames are only coupled by the lip-smoothing EMA; a rank first runs the conditioning kernel (39 us)
    over the SMOOTHING_MEMORY frames in front of its slice, which reproduces the state of a run over the whole sequence
    to below fp32 resolution.  Yields (frame index, pinned host image)."""
    from . import frame as _frame
    lo, hi = sequence_slice(len(packed_frames), world, rank)
    dev = bg_coords.device
    n_in = 24 + int(np.prod(auds_shape))
    model.enc_a = None
    for p in packed_frames[max(0, lo - SMOOTHING_MEMORY):lo]:
        flat = p.to(dev, non_blocking=True)
        _frame.advance_conditioning(model, flat[24:n_in].view(tuple(auds_shape)), flat[22:23] if use_eye else None, flat[16:22])
    streamer = FrameStreamer(model, H, W, intrinsics, bg_coords, auds_shape, use_eye=use_eye, **render_kw)
    idx = lo
    for img in streamer.render_all(packed_frames[lo:hi]):
        yield idx, img
        idx += 1
