"""Training step in steady state (BASELINE configs[3]: head training, 2^16 rays / batch, occupancy-grid update every 16 steps).

    python tools/train_bench.py            # one GPU; writes gpurun_out/train_bench.json and prints it

The reference's loop (nerf/utils.py:1153-1182): `update_extra_state()` every 16th step, zero_grad, train_step under
autocast, GradScaler backward/step/update, scheduler step.  The first 16 steps march with unknown `mean_count` (buffers
sized for the worst case, a device->host read and an empty_cache per step, raymarching.py:213-256); from the first grid
update on `mean_count` sizes the buffers and the step runs without that sync.  This tool times BOTH regimes and, in steady
state, both optimiser tails: torch.optim.Adam (what the reference runs) and radnerf_b200.optim.FusedAdam (one sweep).
A second pass brackets the phases of a step with device synchronisations to show where the time goes."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
from radnerf_b200 import synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options
from radnerf_b200.optim import FusedAdam
from radnerf_b200.train import GraphedTrainStep, head_loss, train_step

dev = torch.device("cuda", 0)
n_rays = int(os.environ.get("N_RAYS", 65536))
steps = int(os.environ.get("STEPS", 64))


def make():
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    m = m.to(dev)
    # what main.py:210-212 hands the model for update_extra_state
    m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))
    m.eye_area = torch.full((600, 1), 0.25)
    m.poses = torch.from_numpy(np.stack([syn.orbit_pose(yaw_deg=float(y), pitch_deg=2.0) for y in np.linspace(-10, 10, 16)]))
    if os.environ.get("RADNERF_FUSED_TRAIN", "1") == "0":      # op-by-op network (cuBLAS + elementwise launches): the round-1 step
        m.fused_train = False
    return m


batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i), dev) for i in range(8)]


def run(kind, regime):
    model = make()
    groups = model.get_params(5e-3, 5e-4)
    if kind == "torch":
        opt = torch.optim.Adam(groups, betas=(0.9, 0.99), eps=1e-15)
    else:
        opt = FusedAdam(groups, betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.1 ** (it / 200000))
    scaler = torch.amp.GradScaler("cuda")
    g = 0
    graphed = GraphedTrainStep(model, opt, scaler) if kind == "graphed" else None

    def one(i):
        nonlocal g
        if regime == "steady" and g % model.opt.update_extra_interval == 0 and g > 0:
            with torch.autocast("cuda", dtype=torch.float16):
                model.update_extra_state()
        g += 1
        loss = graphed(batches[i % 8]) if graphed is not None else train_step(model, batches[i % 8], opt, scaler, None)
        sched.step()
        return loss

    for i in range(20 if regime == "steady" else 4):       # steady: 16 cold steps, the first grid update, 4 more
        one(i)
    if regime == "steady":
        assert model.mean_count > 0
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    samples = float(model.step_counter[:, 0].float().mean())
    replay_ms = None
    if graphed is not None:    # replays only: no grid update, no re-capture in between
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for i in range(12):
            graphed(batches[i % 8])
        r1.record()
        torch.cuda.synchronize()
        replay_ms = r0.elapsed_time(r1) / 12
    extra = {} if graphed is None else {"captures": graphed.captures, "replays": graphed.replays, "fallback_reason": graphed.fallback_reason,
                                             "capture_ms": graphed.capture_ms, "replay_only_ms_per_step": replay_ms}
    return {**extra, "optimizer": kind, "regime": regime, "network": "fused kernels" if model.fused_train else "op by op", "ms_per_step": ms, "host_ms_per_step": (time.perf_counter() - t0) * 1e3 / steps,
            "rays_per_s": n_rays / ms * 1e3, "samples_per_step": samples, "msamples_per_s": samples / ms / 1e3,
            "loss": float(loss), "mean_count": int(model.mean_count), "steps": steps}, model, opt, scaler


def phases(model, opt, scaler, reps=8):
    """one step cut into phases, a device synchronisation after each (so a phase costs max(host, device))"""
    acc = {}

    def lap(name, t):
        torch.cuda.synchronize()
        acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t) * 1e3 / reps
        return time.perf_counter()

    model.train()
    for r in range(reps):
        b = batches[r % 8]
        torch.cuda.synchronize()
        t = time.perf_counter()
        opt.zero_grad(set_to_none=False)
        t = lap("zero_grad", t)
        with torch.autocast("cuda", dtype=torch.float16):
            out = model.render(b["rays_o"], b["rays_d"], b["auds"], b["bg_coords"], b["poses"], eye=b["eye"], index=b["index"],
                               bg_color=b["bg_color"], perturb=True, force_all_rays=False, **model.opt.render_kwargs())
            loss = head_loss(out, b["rgb"], b["face_mask"], 0.1)
        t = lap("forward", t)
        scaler.scale(loss).backward()
        t = lap("backward", t)
        scaler.step(opt)
        scaler.update()
        t = lap("optimizer_tail", t)
    return acc


results = []
runs = [tuple(a.split(":")) for a in sys.argv[1:]] or [("torch", "cold"), ("torch", "steady"), ("fused", "steady"), ("graphed", "steady")]
for kind, regime in runs:
    r, model, opt, scaler = run(kind, regime)
    if regime == "steady" and kind != "graphed":
        r["phases_ms"] = phases(model, opt, scaler)
    results.append(r)
    print(json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"n_rays": n_rays, "gpu": torch.cuda.get_device_name(0), "runs": results},
          open(os.path.join(ROOT, "gpurun_out", "train_bench.json"), "w"), indent=1)
