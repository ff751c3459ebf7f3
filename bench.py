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
of this path -- its extensions are CUDA-only -- so oracle/ is the CPU arm, kind "port"), and reports next to it the STOCK reference
classes on the reference's own CUDA extensions (baseline/stock_bench.py).

Timing: the driver's `--steps K` window is repeated R >= 15 times (each repeat: K frames, drained, bracketed by CUDA events and a
barrier; max over ranks per repeat); `ms_per_step` / `value` come from the MEDIAN window, `repeats` and the spread are in the line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# more hardware connections than the default 8: frame lanes x side branches must not share queues (see radnerf_b200/__init__.py);
# has to be in the environment before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
    ap.add_argument("--steps", type=int, default=200, help="frames per timed window (the window is repeated, see --repeats)")
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--repeats", type=int, default=0, help="timed windows of --steps frames (0 = at least 15, enough to cover ~1 s)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the BASELINE configs[0], [1], [4] legs")
    ap.add_argument("--hw", type=int, default=HW)
    ap.add_argument("--lanes", type=int, default=0, help="frames in flight per GPU (fused path; 1 = strictly one frame at a time; 0 = 8: a frame is a "
                    "chain of ~20 dependent launches of which only the network kernel fills the GPU -- and on 4-8 GPUs not even that)")
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
def cpu_frame_rate(hw, frames_to_time=1, warmup=1, threads=None, budget_s=None):
    """the oracle port (C kernels + torch-CPU MLPs) rendering whole frames on the host cores.  -> (frames/s from the MEAN frame time,
    threads, sample slots per frame, frames really timed)"""
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
    t_start = time.perf_counter()
    with torch.no_grad():
        for i in range(warmup + frames_to_time):
            f = frames[i % len(frames)]
            ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
            t0 = time.perf_counter()
            out = model.render(torch.from_numpy(ro)[None], torch.from_numpy(rd)[None], torch.from_numpy(f["auds"]), bg_t,
                               torch.from_numpy(f["pose6"]), eye=torch.from_numpy(f["eye"]), index=0, bg_color=None,
                               perturb=False, **model.opt.render_kwargs())
            float(out["image"].sum())
            if i >= warmup:
                times.append(time.perf_counter() - t0)
            # bounded sample: stop early (never before one timed frame) when the wall-clock budget is used up
            if budget_s is not None and times and time.perf_counter() - t_start > budget_s:
                break
    t = sum(times) / len(times)
    return 1.0 / t, threads, sum(s[2] for s in model.last_frame_stats), len(times)


WORKLOAD = ("RAD-NeRF head+torso inference (BASELINE configs[2]), %dx%d, obama_eo shapes (wav2vec 44-d x16, att=2, exp_eye, "
            "ind codes), random-init, synthetic head occupancy")


def run_reference_arm(args):
    """--impl reference: the CPU port on all host threads, same metric/config; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    warm = max(1, min(args.warmup, 2))
    fps, threads, nsamp, timed = cpu_frame_rate(args.hw, frames_to_time=max(1, args.steps), warmup=warm, budget_s=75.0)
    sample = ("%d full %dx%d head+torso frame(s) after %d warm-up frame(s) on the CPU port (oracle C kernels + torch-CPU MLPs, fp32), mean frame "
              "time, %d sample slots/frame%s" % (timed, args.hw, args.hw, warm, nsamp,
                                                 "" if timed == args.steps else "; stopped at the 75 s budget of this arm (%d steps requested)" % args.steps))
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": timed, "steps_requested": args.steps, "warmup": warm,
            "ms_per_step": 1000.0 / fps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD % (args.hw, args.hw), "frame": [args.hw, args.hw], "rays_per_frame": args.hw * args.hw,
                       "note": "the reference's extensions are CUDA-only; its CPU implementation is the oracle port (oracle/), all host threads"},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_train:
        try:
            line["train"] = cpu_training_rate()
        except Exception as e:  # noqa: BLE001
            line["train"] = {"unavailable": repr(e)[:200]}
    # context, next to the CPU number this arm is about: the reference's OWN classes (baseline/_ref/reference/nerf, unmodified) on the
    # reference's OWN CUDA extensions (oracle/_ref, built from /root/reference by oracle/build_ref.py) on this box's GPU, and the same
    # classes on this repository's drop-in packages.  None of radnerf_b200's model / renderer code is on these paths.
    if not args.no_ref_cuda:
        try:
            import torch
            if torch.cuda.is_available():
                import contextlib
                from baseline import stock_bench
                dev = torch.device("cuda", 0)
                with contextlib.redirect_stdout(sys.stderr):      # the reference's wrappers print at import; stdout carries ONE JSON line
                    line["ref_cuda"] = stock_bench.frame_rate(dev, args.hw, 30, backend="ref")
                    line["stock_on_dropin"] = stock_bench.frame_rate(dev, args.hw, 30, backend="ours")
                    if not args.no_train:
                        line["ref_cuda_train"] = stock_bench.train_rate(dev, backend="ref")
                        line["stock_on_dropin_train"] = stock_bench.train_rate(dev, backend="ours")
        except Exception as e:  # noqa: BLE001
            line["ref_cuda_error"] = repr(e)[:300]
    line["wall_s"] = time.perf_counter() - t0
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
    if args.lanes <= 0:
        args.lanes = 8
    model = make_model(dev)
    frames, intr, bg = make_frames(hw)
    bg_t = torch.from_numpy(bg).to(dev)[None]
    kw = model.opt.render_kwargs()
    path = "fused"    # the only inference path of the product; a broken fused renderer fails the bench (no fallback)
    from radnerf_b200 import frame  # noqa: F401

    from radnerf_b200.sharding import FrameSharder
    sharder = FrameSharder(hw, hw, world, rank, dev)
    gather_impl = "none" if world == 1 else ("gather-to-root: peer stores over NVLink into rank 0's frame buffer (symmetric memory), arrival counter + "
                                              "consumed flags (release/acquire, system scope), no barrier" if sharder.enable_peer_gather(n_buffers=max(2, args.lanes))
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

    lanes = max(1, args.lanes)

    def streamer_loop(st, blocks):
        def render(i):
            if st.in_flight() == st.depth:
                st.collect()
            st.submit(blocks[i % len(blocks)])

        def drain():
            while st.in_flight():
                st.collect()
            st.sync()
        return render, drain

    # `value`: the per-frame input blocks (pose, pose6, eye, audio window) are resident on the device; rays are generated
    # on the device from the pose, `lanes` frames are in flight on separate streams (FramePipeline inside FrameStreamer;
    # the lip-smoothing chain is kept by running the conditioning kernels in frame order on their own stream); no
    # host<->device copies in the timed region.
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames]
    packed_dev = [p.to(dev) for p in packed]
    resident = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                             deliver=False, depth=lanes, **kw)
    render_resident, drain_resident = streamer_loop(resident, packed_dev)

    # ---- host inputs for `e2e`: one pinned block per frame (pose, pose6, eye, audio window); the public streaming API
    #      (radnerf_b200.stream.FrameStreamer) copies it in, generates the rays on the device, renders, all-gathers the tiles
    #      and copies the image back to pinned host memory -- the device->host copy of frame i overlaps frame i+1 (depth 2)
    streamer = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                             deliver=(rank == 0), depth=max(2, lanes), **kw)
    h2d_bytes, d2h_bytes = streamer.h2d_bytes, streamer.d2h_bytes
    render_e2e, drain_e2e = streamer_loop(streamer, packed)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, drain=None, mdl=None, repeats=None):
        """-> dict(ms = median window, windows, launches per window, clocks).  A window = `steps` calls of fn + drain, bracketed by CUDA
        events and a barrier; per window the MAX over ranks is kept."""
        mdl = mdl if mdl is not None else model
        mdl.enc_a = None
        fused = getattr(mdl, "_fused", None)

        def loop_iters():   # over all frame lanes
            sts = ([fused] if fused is not None else []) + list(getattr(mdl, "_fused_lanes", {}).values())
            return sum(s.loop_iterations() for s in sts if s.workspace is not None)
        # the clock sampler (nvidia-smi -lms 100) starts before the warm-up so that it is already reporting when the timed
        # region runs; warm-up and timed steps are the same load
        with ClockSampler(local) as cs:
            for i in range(warmup):
                fn(i)
            if drain:
                drain()
            barrier()
            windows, launches, n = [], 0, warmup
            R = repeats or args.repeats or 15
            r = 0
            while r < R:
                l0, it0 = abi.launch_count(), loop_iters()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    fn(n + i)
                if drain:
                    drain()   # every timed frame is delivered before the clock stops
                e1.record()
                barrier()
                n += steps
                windows.append(e0.elapsed_time(e1))
                lw = abi.launch_count() - l0
                if fused is not None and fused.use_graph:
                    # a captured frame was counted as its `capture_unroll` plain iterations plus ONE pass of the WHILE node's body;
                    # replace that one pass by the body's real executions (lower bound when some frame needs < capture_unroll)
                    lw += 3 * (max(0, (loop_iters() - it0) - fused.capture_unroll * steps) - steps)
                launches += lw
                r += 1
                if r == 3 and not (repeats or args.repeats):
                    # enough windows to cover ~1 s (so that nvidia-smi's 100 ms sampler sees the timed load), the same count on every rank
                    want = torch.tensor([min(400, max(15, int(1000.0 / max(1e-3, float(np.median(windows))))))], device=dev)
                    if world > 1:
                        dist.all_reduce(want, op=dist.ReduceOp.MAX)
                    R = int(want.item())
        w = torch.tensor(windows, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        w = w.cpu().numpy()
        return {"ms": float(np.median(w)), "min": float(w.min()), "max": float(w.max()), "repeats": int(len(w)),
                "launches": launches // max(1, len(w)), "clocks": cs.summary()}

    W = max(3, args.warmup)
    K = args.steps
    t_res = timed(render_resident, K, W, drain_resident)
    t_e2e = timed(render_e2e, K, W, drain_e2e)
    ms, launches, clocks = t_res["ms"], t_res["launches"], t_res["clocks"]
    ms_e2e = t_e2e["ms"]
    e2e_u8 = None
    latency = None
    # the same end-to-end loop with the output stage on the device (uint8 frames, what the reference's video writer consumes)
    streamer8 = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                              deliver=(rank == 0), depth=max(2, lanes), output="uint8", **kw)
    t_u8 = timed(*((lambda rd: (rd[0], K, W, rd[1]))(streamer_loop(streamer8, packed))))
    e2e_u8 = {"value": K / (t_u8["ms"] / 1e3), "unit": UNIT, "h2d_bytes_per_step": streamer8.h2d_bytes,
              "d2h_bytes_per_step": streamer8.d2h_bytes, "repeats": t_u8["repeats"], "what": "as e2e, but frames are converted to uint8 on the device "
              "((pred * 255).astype(uint8), the reference's host-side expression) before the copy-out"}
    # single-frame latency: ONE frame in flight, host block in -> image on the host, next frame only after that
    one = FrameStreamer(model, hw, hw, intr, bg_local[0], frames[0]["auds"].shape, use_eye=True, sharder=sharder,
                        deliver=(rank == 0), depth=1, **kw)

    def render_one(i):
        one.submit(packed[i % len(packed)])
        one.collect()
    t_lat = timed(render_one, K, W, None, repeats=15)
    latency = t_lat["ms"] / K
    fps, fps_e2e = K / (ms / 1e3), K / (ms_e2e / 1e3)

    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "repeats": t_res["repeats"],
            "ms_per_step": ms / K, "window_ms": {"median": ms, "min": t_res["min"], "max": t_res["max"]},
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "impl": "ours",
            "config": {"workload": WORKLOAD % (hw, hw),
                       "frame": [hw, hw], "rays_per_frame": hw * hw, "path": path, "frames_in_flight": lanes,
                       "timing": "median of %d windows of %d frames each (every window drained and bracketed by CUDA events + a barrier, max over ranks)" % (t_res["repeats"], K),
                       "parallelism": "rays of each frame sharded by interleaved row tiles over %d GPU(s), all-gather of image tiles: %s" % (world, gather_impl),
                       "l2": "every frame regenerates its %.0f MB of rays from the pose and rewrites its %.0f MB workspace; with %d frames in "
                             "flight the streamed working set is %.0f MB (> 126 MB L2 for >= 4 lanes); hash tables (8 MB) and weights are "
                             "re-used across frames by design" % (hw * hw * 24 / 1e6, 21.0 * hw * hw / 262144, lanes,
                                                                  lanes * (hw * hw * 24 / 1e6 + 21.0 * hw * hw / 262144 + 8.5 * hw * hw / 262144))},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": fps_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "api": "radnerf_b200.stream.FrameStreamer (%d frames in flight: the image copy-out of frame i overlaps the following frames)" % max(2, lanes),
                    "ms_per_step": ms_e2e / K, "repeats": t_e2e["repeats"], "window_ms": {"median": ms_e2e, "min": t_e2e["min"], "max": t_e2e["max"]}}}
    if e2e_u8 is not None:
        line["e2e_uint8"] = e2e_u8
    if latency is not None:
        line["latency_ms_1lane"] = latency
        line["latency_note"] = ("one frame in flight: host input block in -> fp32 image on the host, the next frame submitted only then "
                                "(streaming / config[4] cares about this, `value` is a %d-frames-in-flight throughput)" % lanes)

    # ---- N > 1: the same job in FRAME-parallel mode (whole frames per GPU, no collective), reported next to the ray-sharded
    #      headline: ray sharding cuts latency, but a 512x512 frame cannot scale past its ~0.29 ms dependency chain
    if world > 1:
        from radnerf_b200.stream import FrameStreamer as _FS
        full = _FS(model, hw, hw, intr, bg_t[0], frames[0]["auds"].shape, use_eye=True, deliver=True, depth=max(2, lanes), **kw)
        t_fp = timed(*((lambda rd: (rd[0], K, W, rd[1]))(streamer_loop(full, packed))))
        line["frame_parallel"] = {"value": world * K / (t_fp["ms"] / 1e3), "unit": UNIT, "ms_per_step_per_gpu": t_fp["ms"] / K, "repeats": t_fp["repeats"],
                                  "what": "every GPU renders WHOLE frames of its own slice of the sequence end to end (host inputs in, fp32 "
                                          "image back on the host; radnerf_b200.stream.render_sequence), no collective; weak scaling"}
        model.enc_a = None

    # ---- the other BASELINE configurations, as extra keys (every rank takes part: the frames are ray-sharded like the headline)
    if not args.no_extra_configs:
        try:
            line["configs"] = extra_configs(args, dev, world, rank, timed, streamer_loop)
        except Exception as e:  # noqa: BLE001
            line["configs"] = {"error": repr(e)[:300]}

    # ---- BASELINE configs[3]: the training step, data-parallel over all ranks (every rank its own 2^16 rays)
    if not args.no_train:
        try:
            line["train"] = training_rate(dev, world=world, rank=rank)
        except Exception as e:  # noqa: BLE001  (context next to the headline, never the reason the line is missing)
            line["train"] = {"unavailable": repr(e)[:300]}

    if rank == 0:
        from radnerf_b200 import roofline
        line["roofline"], line["kernels"] = roofline.measure(model, dev_frames[0], bg_local, kw, path)
        if world == 1 and not args.no_cpu_baseline:
            fps_cpu, threads, nsamp, _ = cpu_frame_rate(hw, 1)
            line["cpu_baseline"] = {"value": fps_cpu, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "1 full %dx%d head+torso frame (after 1 warm-up frame) on the CPU port (oracle C kernels, OpenMP, + "
                                              "torch-CPU MLPs, fp32), %d sample slots" % (hw, hw, nsamp)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_configs(args, dev, world, rank, timed, streamer_loop):
    """BASELINE configs[1] (head only, 450x450, wav2vec 44-d; one GPU) and configs[4] (DeepSpeech 29-d model with eye / individual codes,
    1024x1024, audio windows streamed through the device-side FeatureRing; rays sharded over all ranks), end to end through
    FrameStreamer -- host input block in, fp32 image back on rank 0's host."""
    import torch
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.audio_ring import FeatureRing
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200.posemath import convert_poses
    from radnerf_b200.sharding import FrameSharder
    from radnerf_b200.stream import FrameStreamer, pack_inputs
    out = {}
    K, W = args.steps, max(3, args.warmup)

    def build(torso, asr_model, hw, dim, share):
        torch.manual_seed(0)
        m = NeRFNetwork(Options(torso=torso, smooth_lips=True, fp16=True, exp_eye=True, asr_model=asr_model))
        grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
        m.density_grid.copy_(torch.from_numpy(grid))
        m.mean_density = float(np.clip(grid, 0, None).mean())
        m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
        if torso:
            tg = syn.torso_density_grid(128)
            m.density_grid_torso.copy_(torch.from_numpy(tg))
            m.mean_density_torso = float(tg.mean())
        m = m.eval().to(dev)
        sh = FrameSharder(hw, hw, world if share else 1, rank if share else 0, dev)
        if share and world > 1:
            sh.enable_peer_gather(n_buffers=max(2, args.lanes))
        bank = syn.audio_feature_bank(600, dim, 16, seed=0)
        intr = syn.intrinsics_for(hw, hw)
        bg = torch.from_numpy(syn.get_bg_coords(hw, hw)).to(dev)
        blocks = []
        for i in range(16):
            pose = syn.orbit_pose(yaw_deg=10.0 * np.sin(2 * np.pi * i / 16), pitch_deg=2.0)
            blocks.append(pack_inputs(pose, syn.audio_window(bank, 8 + i, 2), convert_poses(torch.from_numpy(pose)[None]).numpy(), np.array([[0.25]], np.float32)))
        st = FrameStreamer(m, hw, hw, intr, sh.shard(bg), (8, dim, 16), use_eye=True, sharder=sh, deliver=(rank == 0), depth=max(2, args.lanes),
                           **m.opt.render_kwargs())
        return m, st, blocks, bank

    if world == 1:
        m, st, blocks, _ = build(False, "cpierse/wav2vec2-large-xlsr-53-esperanto", 450, 44, False)
        r, d = streamer_loop(st, blocks)
        t = timed(r, K, W, d, mdl=m)
        out["configs1_head_only_450"] = {"value": K / (t["ms"] / 1e3), "unit": UNIT, "ms_per_step": t["ms"] / K, "repeats": t["repeats"],
                                         "workload": "BASELINE configs[1]: head-only inference, obama_eo shapes (wav2vec 44-d x16), 450x450, random-init, end to "
                                                     "end through FrameStreamer (host block in, fp32 image out), %d frames in flight" % st.depth}
        del m, st
    # configs[4]: the ASR feature rows live in a device ring (audio_ring.FeatureRing = the reference's ASR.feat_queue); per frame the host sends
    # the 24-float head [pose | pose6 | eye] only, two new feature rows are pushed (25 fps video on 50 Hz features) and the frame's
    # [8, 29, 16] window is gathered on the device straight into the lane's input block
    m, st, blocks, bank = build(True, "deepspeech", 1024, 29, True)
    # reference defaults (nerf/asr.py): feat_buffer_size 4, -m 50 -> a context of 50 feature rows lands every 25 video frames
    ring = FeatureRing(slots=4, context=50, dim=29, device=dev)
    feats = torch.from_numpy((np.random.default_rng(3).standard_normal((4000, 29)) * 3).astype(np.float32)).to(dev)
    heads = [b[:24].clone().pin_memory() for b in blocks]
    state = {"row": 0}

    def push(i):
        if i % 25 == 0:
            r0 = state["row"] % (feats.shape[0] - 50)
            ring.push(feats[r0:r0 + 50])
            state["row"] += 50

    def render(i):
        if st.in_flight() == st.depth:
            st.collect()
        push(i)
        st.submit(heads[i % len(heads)], ring=ring)

    def drain():
        while st.in_flight():
            st.collect()
        st.sync()
    t = timed(render, K, W, drain, mdl=m)
    one = FrameStreamer(m, 1024, 1024, syn.intrinsics_for(1024, 1024), st.bg, (8, 29, 16), use_eye=True, sharder=st.sharder, deliver=(rank == 0), depth=1,
                        **m.opt.render_kwargs())

    def render_one(i):
        push(i)
        one.submit(heads[i % len(heads)], ring=ring)
        one.collect()
    tl = timed(render_one, K, W, None, mdl=m, repeats=15)
    out["configs4_deepspeech_1024_streaming"] = {
        "value": K / (t["ms"] / 1e3), "unit": UNIT, "ms_per_step": t["ms"] / K, "repeats": t["repeats"], "frame_latency_ms": tl["ms"] / K,
        "n_gpus": world, "h2d_bytes_per_step": 96, "d2h_bytes_per_step": st.d2h_bytes,
        "workload": "BASELINE configs[4]: DeepSpeech-feature (29-d) model with eye / individual codes + torso, 1024x1024 streaming render: audio "
                    "windows gathered on the device from the FeatureRing, rays sharded over %d GPU(s), fp32 image back on rank 0's host, %d frames in "
                    "flight; frame_latency_ms = one frame in flight" % (world, st.depth)}
    return out


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


def training_rate(dev, n_rays=65536, steps=64, ops=None, tail="fused", graphed=True, world=1, rank=0):
    """BASELINE configs[3], steady state: head training step on 2^16 rays PER GPU (march_rays_train -> fused encoders + MLPs ->
    composite_rays_train, backward, GradScaler, FusedAdam one-sweep tail), occupancy-grid update every 16 steps inside the timed
    region (nerf/utils.py:1153-1182); with world > 1 data-parallel: every rank draws its own rays, gradients are averaged with
    NCCL all-reduces (two hash tables + one flat bucket), the occupancy update is replicated (radnerf_b200.train).  The first 16
    steps (unknown mean_count: worst-case buffers + a host read per step, raymarching.py:213-256) and the first grid update are
    warm-up, as in any run longer than 16 steps.  graphed: the step is replayed from CUDA graphs (train.GraphedTrainStep)."""
    import torch
    import torch.distributed as dist
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200.optim import FusedAdam
    from radnerf_b200.train import GradSync, GraphedTrainStep, train_step, update_extra_state_replicated
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True), ops=ops)
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    m = m.to(dev)
    m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))     # main.py:210-212
    m.eye_area = torch.full((600, 1), 0.25)
    batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i, seed=rank), dev) for i in range(8)]
    if tail == "fused":
        opt = FusedAdam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
    else:
        opt = torch.optim.Adam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)      # main.py:204
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** (it / 200000))
    scaler = torch.amp.GradScaler("cuda")
    sync = GradSync(m.parameters()) if world > 1 else None
    g = [0]
    graphed = GraphedTrainStep(m, opt, scaler, sync=sync) if (graphed and tail == "fused") else None
    upd = []

    def one(i):
        if g[0] % m.opt.update_extra_interval == 0 and g[0] > 0:
            u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            u0.record()
            with torch.autocast("cuda", dtype=torch.float16):
                update_extra_state_replicated(m) if world > 1 else m.update_extra_state()
            u1.record()
            upd.append((u0, u1))
        g[0] += 1
        loss = graphed(batches[i % 8]) if graphed is not None else train_step(m, batches[i % 8], opt, scaler, sync)
        sched.step()
        return loss
    # warm-up: 16 cold steps, then four occupancy updates -- with random-init weights the re-queried density grid grows for the first
    # few updates (mean_count 0.2 M -> 0.33 M -> 0.69 M samples on the synthetic head) before the sample count per step settles
    for i in range(70):
        one(i)
    upd.clear()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    upd_ms = [a.elapsed_time(b) for a, b in upd]
    samples = float(m.step_counter[:, 0].float().mean())
    res = {}
    if world > 1:
        t = torch.tensor([ms, samples], device=dev, dtype=torch.float64)
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, samples = float(mx[0]), float(t[1])          # slowest rank's step time, samples of ALL ranks per step
        # the gradient exchange alone: the same three all-reduces on the live gradient buffers, back to back
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(20):
            sync.begin_step()
            sync.finish()
        a1.record()
        torch.cuda.synchronize()
        ar = torch.tensor([a0.elapsed_time(a1) / 20], device=dev)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        same = True
        for p in m.parameters():                    # replicas must have stayed bit-identical
            lo, hi = p.detach().clone(), p.detach().clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            same = same and bool(torch.equal(lo, hi))
        res.update(allreduce_us_exposed=float(ar.item()) * 1e3, allreduce_bytes_per_step=sync.bytes_last, replicas_identical=same,
                   parallelism="data-parallel over %d GPUs: %d rays per GPU per step, NCCL all-reduce (AVG) of 2 hash-table gradients + one flat "
                               "bucket, issued between the backward graph and the optimiser graph (not overlapped: allreduce_us_exposed is "
                               "its whole cost per step); replicated occupancy update" % (world, n_rays))
    how = "torch.optim.Adam, op by op" if tail != "fused" else "FusedAdam one-sweep tail, op by op"
    if graphed is not None:
        how = "fused head forward / backward kernels, FusedAdam tail, step replayed from CUDA graph(s) (%d captures, %d replays in the run%s)" % (
            graphed.captures, graphed.replays, "" if graphed.fallback_reason is None else "; FELL BACK to op-by-op: " + graphed.fallback_reason)
    res.update({"workload": "RAD-NeRF head training step (BASELINE configs[3]): %d rays/batch per GPU, fp16 autocast, grid update every 16 steps, "
                            "%s" % (n_rays, how), "n_gpus": world, "ms_per_step": ms, "rays_per_s": world * n_rays / ms * 1e3,
                "samples_per_step": samples, "msamples_per_s": samples / ms / 1e3, "steps": steps, "loss": float(loss),
                "update_extra_state_ms": float(np.mean(upd_ms)) if upd_ms else None, "scaling": "weak"})
    return res


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner, the reference's wrappers at import)
    # are sent to stderr for the duration of the run, the line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = []
    import builtins
    _print = builtins.print

    def emit(*a, **k):
        if k.get("file") in (None, sys.stdout) and len(a) == 1 and isinstance(a[0], str) and a[0].startswith("{"):
            captured.append(a[0])
        else:
            _print(*a, **k)
    builtins.print = emit
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_ours(args)
    finally:
        builtins.print = _print
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for line in captured:
        print(line, flush=True)


if __name__ == "__main__":
    main()
