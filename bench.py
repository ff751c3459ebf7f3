#!/usr/bin/env python
"""bench.py -- RAD-NeRF head+torso inference throughput (frames/s at 512x512) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path fused|ops]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one full `render()` of a 512x512 frame: near/far, audio conditioning, the march -> network -> composite
loop of the head, the masked torso pass and the final blend (BASELINE.json configs[2], the configuration the metric is
quoted on).  Synthetic poses / audio windows / occupancy, random-init weights of the obama_eo architecture.

Prints ONE JSON line (rank 0).  `value` = frames/s with the frame's inputs (rays, audio window, pose) already in HBM;
`e2e` = frames/s through the public API with HOST inputs (pose, intrinsics, audio window, eye) copied in and the fp32
image copied out every frame; `roofline` = the dominant kernel against the measured B200 peak; `cpu_baseline` = the
oracle port on the host cores.  `--impl reference` times the CPU port alone (the reference has no CPU implementation
of this path -- its extensions are CUDA-only -- so oracle/ is the CPU arm, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "rad-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "frames/sec at 512x512 (head+torso)"
UNIT = "frames/s"
HW = 512
N_FRAMES_DISTINCT = 64  # distinct poses / audio windows cycled through


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default=os.environ.get("RADNERF_PATH", "auto"), choices=["auto", "fused", "ops"])
    ap.add_argument("--hw", type=int, default=HW)
    ap.add_argument("--lanes", type=int, default=4, help="frames in flight per GPU (fused path; 1 = strictly one frame at a time)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE configs[3]) of the N=1 line")
    ap.add_argument("--no-ref-cuda", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ scene
def make_model(device, ops=None, seed=0, fp16=True):
    import torch
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200 import synthetic as syn
    torch.manual_seed(seed)
    opt = Options(torso=True, smooth_lips=True, fp16=fp16, exp_eye=True)
    model = NeRFNetwork(opt, ops=ops)
    # "trained-like" occupancy: analytic head (+neck), dilated and packed with the reference's threshold rule
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))  # ~30% of the rays hit it, like the recorded obama trace
    model.density_grid.copy_(torch.from_numpy(grid))
    model.mean_density = float(np.clip(grid, 0, None).mean())
    model.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(model.mean_density, model.density_thresh))))
    tg = syn.torso_density_grid(128)
    model.density_grid_torso.copy_(torch.from_numpy(tg))
    model.mean_density_torso = float(tg.mean())
    model.eval()
    return model.to(device)


def make_frames(hw, n=N_FRAMES_DISTINCT):
    """host-side per-frame inputs: pose (4x4), 6-vector pose, audio window, eye"""
    import torch
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.posemath import convert_poses
    bank = syn.audio_feature_bank(600, 44, 16, seed=0)
    frames = []
    for i in range(n):
        yaw = 10.0 * np.sin(2 * np.pi * i / n)
        pose = syn.orbit_pose(yaw_deg=yaw, pitch_deg=2.0)
        frames.append(dict(pose=pose, pose6=convert_poses(torch.from_numpy(pose)[None]).numpy(),
                           auds=syn.audio_window(bank, 8 + i, 2), eye=np.array([[0.25]], np.float32)))
    return frames, syn.intrinsics_for(hw, hw), syn.get_bg_coords(hw, hw)


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=1)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_frame_rate(hw, frames_to_time=1, threads=None):
    """the oracle port (C kernels + torch-CPU MLPs) rendering whole frames on the host cores"""
    import torch
    from oracle.cpu_backend import CPUOps
    from oracle import oracle as O
    from radnerf_b200 import synthetic as syn
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    O.set_num_threads(threads)
    model = make_model("cpu", ops=CPUOps(), fp16=False)
    frames, intr, bg = make_frames(hw, 4)
    bg_t = torch.from_numpy(bg)[None]
    times = []
    with torch.no_grad():
        for i in range(frames_to_time + 1):
            f = frames[i % len(frames)]
            ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
            t0 = time.perf_counter()
            out = model.render(torch.from_numpy(ro)[None], torch.from_numpy(rd)[None], torch.from_numpy(f["auds"]), bg_t,
                               torch.from_numpy(f["pose6"]), eye=torch.from_numpy(f["eye"]), index=0, bg_color=None,
                               perturb=False, **model.opt.render_kwargs())
            float(out["image"].sum())
            times.append(time.perf_counter() - t0)
    t = min(times[1:]) if len(times) > 1 else times[0]
    return 1.0 / t, threads, sum(s[2] for s in model.last_frame_stats)


WORKLOAD = ("RAD-NeRF head+torso inference (BASELINE configs[2]), %dx%d, obama_eo shapes (wav2vec 44-d x16, att=2, exp_eye, "
            "ind codes), random-init, synthetic head occupancy")


def run_reference_arm(args):
    """--impl reference: the CPU port on all host threads, same metric/config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    fps, threads, nsamp = cpu_frame_rate(args.hw, frames_to_time=steps)
    sample = "%d full %dx%d head+torso frame(s) on the CPU port (oracle C kernels + torch-CPU MLPs, fp32), %d sample slots/frame" % (
        steps, args.hw, args.hw, nsamp)
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD % (args.hw, args.hw), "frame": [args.hw, args.hw], "rays_per_frame": args.hw * args.hw,
                       "note": "the reference's extensions are CUDA-only; its CPU implementation is the oracle port (oracle/), all host threads"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t0}
    if not args.no_train:
        try:
            line["train"] = cpu_training_rate()
        except Exception as e:  # noqa: BLE001
            line["train"] = {"unavailable": repr(e)[:200]}
    # context, next to the CPU number this arm is about: the reference's OWN CUDA extensions (oracle/_ref, built from
    # /root/reference by oracle/build_ref.py) driven in the reference's op order on this box's GPU, if there is one
    if not args.no_ref_cuda:
        try:
            import torch
            if torch.cuda.is_available():
                line["ref_cuda"] = ref_cuda_frame_rate(torch.device("cuda", 0), args.hw, 30)
                if not args.no_train:
                    from oracle import ref_backend
                    line["ref_cuda_train"] = training_rate(torch.device("cuda", 0), ops=ref_backend.RefOps(train=True), tail="torch")
                    line["ref_cuda_train"]["what"] = ("the same training loop on the reference's CUDA extensions (oracle/_ref) in the "
                                                      "reference's op order, torch.optim.Adam + GradScaler as main.py / nerf/utils.py run them")
        except Exception as e:  # noqa: BLE001
            line["ref_cuda"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"   # keep stdout to the one JSON line (NCCL prints its version banner there)
    import torch
    import torch.distributed as dist
    from radnerf_b200 import abi, synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    abi.lib()  # fail loudly if the CUDA library is missing

    hw = args.hw
    model = make_model(dev)
    frames, intr, bg = make_frames(hw)
    bg_t = torch.from_numpy(bg).to(dev)[None]
    kw = model.opt.render_kwargs()

    path = args.path
    if path == "auto":
        try:
            from radnerf_b200 import frame  # noqa: F401
            path = "fused"
        except Exception:
            path = "ops"

    from radnerf_b200.sharding import FrameSharder
    sharder = FrameSharder(hw, hw, world, rank, dev)
    gather_impl = "none" if world == 1 else ("peer stores over NVLink (symmetric memory) + barrier" if sharder.enable_peer_gather(n_buffers=max(2, args.lanes))
                                              and os.environ.get("RADNERF_GATHER", "peer") == "peer" else "nccl all_gather + un-permute")
    if gather_impl.startswith("nccl"):
        sharder.peer = None

    # ---- device-resident inputs for `value`
    dev_frames = []
    for f in frames:
        ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
        dev_frames.append(dict(ro=sharder.shard(torch.from_numpy(ro).to(dev))[None], rd=sharder.shard(torch.from_numpy(rd).to(dev))[None],
                               auds=torch.from_numpy(f["auds"]).to(dev), pose6=torch.from_numpy(f["pose6"]).to(dev),
                               eye=torch.from_numpy(f["eye"]).to(dev)))
    bg_local = sharder.shard(bg_t[0])[None]

    lanes = max(1, args.lanes) if path == "fused" else 1
    if path == "fused":
        # `value`: the per-frame input blocks (pose, pose6, eye, audio window) are resident on the device; rays are generated
        # on the device from the pose, `lanes` frames are in flight on separate streams (FramePipeline inside FrameStreamer;
        # the lip-smoothing chain is kept by running the conditioning kernels in frame order on their own stream); no
        # host<->device copies in the timed region.
        from radnerf_b200.stream import FrameStreamer, pack_inputs
        packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames]
        packed_dev = [p.to(dev) for p in packed]
        resident = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                                 deliver=False, depth=lanes, **kw)

        def render_resident(i):
            if resident.in_flight() == resident.depth:
                resident.collect()
            resident.submit(packed_dev[i % len(packed_dev)])

        def drain_resident():
            while resident.in_flight():
                resident.collect()
            resident.sync()
    else:
        def render_resident(i):
            f = dev_frames[i % len(dev_frames)]
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=model.opt.fp16):
                out = model.render(f["ro"], f["rd"], f["auds"], bg_local, f["pose6"], eye=f["eye"], index=0, bg_color=None,
                                   perturb=False, path=path, **kw)
            return sharder.gather(out["image"][0])
        drain_resident = None

    # ---- host inputs for `e2e`: one pinned block per frame (pose, pose6, eye, audio window); the public streaming API
    #      (radnerf_b200.stream.FrameStreamer) copies it in, generates the rays on the device, renders, all-gathers the tiles
    #      and copies the image back to pinned host memory -- the device->host copy of frame i overlaps frame i+1 (depth 2)
    if path == "fused":
        streamer = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                                 deliver=(rank == 0), depth=max(2, lanes), **kw)
        h2d_bytes, d2h_bytes = streamer.h2d_bytes, streamer.d2h_bytes

        def render_e2e(i):
            if streamer.in_flight() == streamer.depth:
                streamer.collect()      # frame i-2 is on the host
            streamer.submit(packed[i % len(packed)])

        def drain_e2e():
            while streamer.in_flight():
                streamer.collect()
    else:
        pinned = [dict(pose=torch.from_numpy(f["pose"]).pin_memory(), auds=torch.from_numpy(f["auds"]).pin_memory(),
                       pose6=torch.from_numpy(f["pose6"]).pin_memory(), eye=torch.from_numpy(f["eye"]).pin_memory()) for f in frames]
        host_img = torch.empty(hw * hw, 3, dtype=torch.float32).pin_memory()
        from radnerf_b200.rays import RayGenerator
        raygen = RayGenerator(hw, hw, intr, dev, sharder)
        h2d_bytes = sum(t.numel() * t.element_size() for t in pinned[0].values())
        d2h_bytes = host_img.numel() * host_img.element_size()

        def render_e2e(i):
            p = pinned[i % len(pinned)]
            pose = p["pose"].to(dev, non_blocking=True)
            auds = p["auds"].to(dev, non_blocking=True)
            pose6 = p["pose6"].to(dev, non_blocking=True)
            eye = p["eye"].to(dev, non_blocking=True)
            ro, rd = raygen(pose)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=model.opt.fp16):
                out = model.render(ro[None], rd[None], auds, bg_local, pose6, eye=eye, index=0, bg_color=None, perturb=False,
                                   path=path, **kw)
            img = sharder.gather(out["image"][0])
            if rank == 0:
                host_img.copy_(img, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the frame is only "delivered" once it is on the host

        def drain_e2e():
            pass

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=None):
        model.enc_a = None
        # the clock sampler (nvidia-smi -lms 100) starts before the warm-up so that it is already reporting when the timed
        # region -- a few tenths of a second -- runs; warm-up and timed steps are the same load
        with ClockSampler(local) as cs:
            for i in range(warmup):
                fn(i)
            if drain:
                drain()
            barrier()
            l0 = abi.launch_count()
            fused = getattr(model, "_fused", None)

            def loop_iters():   # over all frame lanes
                sts = ([fused] if fused is not None else []) + list(getattr(model, "_fused_lanes", {}).values())
                return sum(s.loop_iterations() for s in sts if s.workspace is not None)
            it0 = loop_iters()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(warmup + i)
            if drain:
                drain()   # every timed frame is on the host before the clock stops
            e1.record()
            barrier()
        ms = e0.elapsed_time(e1)
        launches = abi.launch_count() - l0
        if fused is not None and fused.use_graph:
            # a captured frame was counted as its `capture_unroll` plain iterations plus ONE pass of the WHILE node's body;
            # replace that one pass by the body's real executions (lower bound when some frame needs < capture_unroll)
            launches += 3 * (max(0, (loop_iters() - it0) - fused.capture_unroll * steps) - steps)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, cs.summary()

    W = max(3, args.warmup)
    ms, launches, clocks = timed(render_resident, args.steps, W, drain_resident)
    ms_e2e, _, _ = timed(render_e2e, args.steps, W, drain_e2e)
    e2e_u8 = None
    if path == "fused":
        # the same end-to-end loop with the output stage on the device (uint8 frames, what the reference's video writer consumes)
        streamer8 = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                                  deliver=(rank == 0), depth=max(2, lanes), output="uint8", **kw)

        def render_u8(i):
            if streamer8.in_flight() == streamer8.depth:
                streamer8.collect()
            streamer8.submit(packed[i % len(packed)])

        def drain_u8():
            while streamer8.in_flight():
                streamer8.collect()
        ms_u8, _, _ = timed(render_u8, args.steps, W, drain_u8)
        e2e_u8 = {"value": args.steps / (ms_u8 / 1e3), "unit": UNIT, "h2d_bytes_per_step": streamer8.h2d_bytes,
                  "d2h_bytes_per_step": streamer8.d2h_bytes, "what": "as e2e, but frames are converted to uint8 on the device "
                  "((pred * 255).astype(uint8), the reference's host-side expression) before the copy-out"}
    fps, fps_e2e = args.steps / (ms / 1e3), args.steps / (ms_e2e / 1e3)

    stats = getattr(model, "last_frame_stats", None)
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "impl": "ours",
            "config": {"workload": WORKLOAD % (hw, hw),
                       "frame": [hw, hw], "rays_per_frame": hw * hw, "path": path, "frames_in_flight": lanes,
                       "parallelism": "rays of each frame sharded by interleaved row tiles over %d GPU(s), all-gather of image tiles: %s" % (world, gather_impl),
                       "l2": "every frame regenerates its %.0f MB of rays from the pose and rewrites its %.0f MB workspace; with %d frames in "
                             "flight the streamed working set is %.0f MB (> 126 MB L2 for >= 4 lanes); hash tables (8 MB) and weights are "
                             "re-used across frames by design" % (hw * hw * 24 / 1e6, 21.0 * hw * hw / 262144, lanes,
                                                                  lanes * (hw * hw * 24 / 1e6 + 21.0 * hw * hw / 262144 + 8.5 * hw * hw / 262144))},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "radnerf_b200.stream.FrameStreamer (depth-2: the image copy-out of frame i overlaps frame i+1)" if path == "fused"
                           else "model.render per frame, synchronous copy-out",
                    "ms_per_step": ms_e2e / args.steps}}
    if e2e_u8 is not None:
        line["e2e_uint8"] = e2e_u8

    # ---- N > 1: the same job in FRAME-parallel mode (whole frames per GPU, no collective), reported next to the ray-sharded
    #      headline: ray sharding cuts latency, but a 512x512 frame cannot scale past its ~0.29 ms dependency chain
    if world > 1 and path == "fused":
        from radnerf_b200.stream import FrameStreamer as _FS
        full = _FS(model, hw, hw, intr, bg_t[0], frames[0]["auds"].shape, use_eye=True, deliver=True, depth=max(2, lanes), **kw)
        model.enc_a = None

        def render_fp(i):
            if full.in_flight() == full.depth:
                full.collect()
            full.submit(packed[i % len(packed)])

        def drain_fp():
            while full.in_flight():
                full.collect()
        ms_fp, _, _ = timed(render_fp, args.steps, W, drain_fp)
        line["frame_parallel"] = {"value": world * args.steps / (ms_fp / 1e3), "unit": UNIT, "ms_per_step_per_gpu": ms_fp / args.steps,
                                  "what": "every GPU renders WHOLE frames of its own slice of the sequence end to end (host inputs in, fp32 "
                                          "image back on the host; radnerf_b200.stream.render_sequence), no collective; weak scaling"}
        model.enc_a = None

    if rank == 0:
        from radnerf_b200 import roofline
        line["roofline"], line["kernels"] = roofline.measure(model, dev_frames[0], bg_local, kw, path)
        if world == 1 and not args.no_cpu_baseline:
            fps_cpu, threads, nsamp = cpu_frame_rate(hw, 1)
            line["cpu_baseline"] = {"value": fps_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "1 full %dx%d head+torso frame on the CPU port (oracle C kernels, OpenMP, + torch-CPU "
                                              "MLPs, fp32), %d sample slots" % (hw, hw, nsamp)}
        if world == 1 and not args.no_train:
            try:
                line["train"] = training_rate(dev)
            except Exception as e:  # noqa: BLE001  (context next to the headline, never the reason the line is missing)
                line["train"] = {"unavailable": repr(e)[:200]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_training_rate(n_rays=65536, steps=2, threads=None):
    """BASELINE configs[3] on the host cores: one head training step of the CPU port (oracle C kernels under autograd +
    torch-CPU layers + torch.optim.Adam, fp32; the reference has no CPU implementation of its own).  A bounded sample: one
    warm-up step and `steps` timed steps of the full 2^16-ray batch (~1-2 s each), cold regime only (mean_count unknown)."""
    import torch
    from oracle import cpu_backend
    from oracle import oracle as O
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200.train import train_step
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    O.set_num_threads(threads)
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=False, exp_eye=True), ops=cpu_backend.CPUOps(train=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    b = syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=0), "cpu")
    opt = torch.optim.Adam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)
    train_step(m, b, opt, None, None)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = train_step(m, b, opt, None, None)
    dt = (time.perf_counter() - t0) / steps
    samples = float(m.step_counter[1:1 + steps, 0].float().mean())
    return {"workload": "RAD-NeRF head training step (BASELINE configs[3]): %d rays/batch on the CPU port, fp32, torch.optim.Adam" % n_rays,
            "ms_per_step": dt * 1e3, "rays_per_s": n_rays / dt, "samples_per_step": samples, "msamples_per_s": samples / dt / 1e6,
            "cores": threads, "kind": "port", "steps": steps, "loss": float(loss)}


def training_rate(dev, n_rays=65536, steps=48, ops=None, tail="fused", graphed=False):
    """BASELINE configs[3] on one GPU, steady state: head training step on 2^16 rays (march_rays_train -> encoders + MLPs ->
    composite_rays_train, backward, GradScaler, FusedAdam one-sweep tail), occupancy-grid update every 16 steps inside the
    timed region (nerf/utils.py:1153-1182).  The first 16 steps (unknown mean_count: worst-case buffers + a host read per
    step, raymarching.py:213-256) and the first grid update are warm-up, as in any run longer than 16 steps."""
    import torch
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200.optim import FusedAdam
    from radnerf_b200.train import GraphedTrainStep, train_step
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True), ops=ops)
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    m = m.to(dev)
    m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))     # main.py:210-212
    m.eye_area = torch.full((600, 1), 0.25)
    batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i), dev) for i in range(8)]
    if tail == "fused":
        opt = FusedAdam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
    else:
        opt = torch.optim.Adam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)      # main.py:204
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** (it / 200000))
    scaler = torch.amp.GradScaler("cuda")
    g = [0]
    # graphed=True replays the step as one CUDA graph between grid updates (radnerf_b200.train.GraphedTrainStep).  Measured in
    # round 1: a replay takes 10.3 ms where the op-by-op step takes 8.5 ms, and every grid update costs a 25-160 ms re-capture --
    # the step is bound by its device work, not by launches -- so the op-by-op step is what this leg times
    graphed = GraphedTrainStep(m, opt, scaler) if (graphed and tail == "fused") else None

    def one(i):
        if g[0] % m.opt.update_extra_interval == 0 and g[0] > 0:
            with torch.autocast("cuda", dtype=torch.float16):
                m.update_extra_state()
        g[0] += 1
        loss = graphed(batches[i % 8]) if graphed is not None else train_step(m, batches[i % 8], opt, scaler, None)
        sched.step()
        return loss
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    samples = float(m.step_counter[:, 0].float().mean())
    how = "torch.optim.Adam, op by op" if tail != "fused" else "FusedAdam one-sweep tail, op by op"
    if graphed is not None:
        how = "FusedAdam tail, step replayed as a CUDA graph (%d captures, %d replays in the run%s)" % (
            graphed.captures, graphed.replays, "" if graphed.fallback_reason is None else "; FELL BACK to op-by-op: " + graphed.fallback_reason)
    return {"workload": "RAD-NeRF head training step (BASELINE configs[3]): %d rays/batch, fp16 autocast, grid update every 16 steps, "
                        "%s" % (n_rays, how), "ms_per_step": ms, "rays_per_s": n_rays / ms * 1e3, "samples_per_step": samples,
            "msamples_per_s": samples / ms / 1e3, "steps": steps, "loss": float(loss)}


def ref_cuda_frame_rate(dev, hw, steps):
    """the reference's own compiled kernels driven in the reference's op order on the same GPU (context, not a contract key)"""
    import torch
    from oracle import ref_backend
    from radnerf_b200 import synthetic as syn
    if not ref_backend.available():
        return {"unavailable": "oracle/_ref/*.so not built"}
    model = make_model(dev, ops=ref_backend.RefOps())
    frames, intr, bg = make_frames(hw, 8)
    bg_t = torch.from_numpy(bg).to(dev)[None]
    kw = model.opt.render_kwargs()
    devf = []
    for f in frames:
        ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
        devf.append((torch.from_numpy(ro).to(dev)[None], torch.from_numpy(rd).to(dev)[None], torch.from_numpy(f["auds"]).to(dev),
                     torch.from_numpy(f["pose6"]).to(dev), torch.from_numpy(f["eye"]).to(dev)))

    def one(i):
        ro, rd, a, p6, eye = devf[i % len(devf)]
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.render(ro, rd, a, bg_t, p6, eye=eye, index=0, bg_color=None, perturb=False, **kw)["image"]
    for i in range(5):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(5 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": 1000.0 / ms, "unit": UNIT, "ms_per_step": ms, "what": "reference CUDA extensions (sm_100a build) + torch "
            "Linear layers in the reference's op order, inputs resident", "sample_slots_per_frame": sum(s[2] for s in model.last_frame_stats)}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
