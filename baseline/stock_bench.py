"""Throughput of the reference's OWN classes (baseline/_ref/reference/nerf/*.py, unmodified) on a GPU, for `bench.py --impl reference`:

    frame_rate(...)   NeRFNetwork.render(..., staged=True, **vars(opt)) per frame, the call Trainer.test_step makes
                      (nerf/utils.py:845-870), under torch.autocast as `-O` runs it;
    train_rate(...)   the body of Trainer.train_one_epoch (nerf/utils.py:1153-1182): update_extra_state every 16 steps, zero_grad,
                      Trainer.train_step (called unbound on a namespace holding what it reads), GradScaler backward / step / update,
                      LambdaLR step, `loss.item()`, on torch.optim.Adam(model.get_params(lr, lr_net), betas=(0.9, 0.99), eps=1e-15)
                      (main.py:204).

backend="ref": the classes sit on the reference's own wrappers + compiled CUDA extensions (oracle/_ref/*.so) -- the reference's CUDA
path, the yardstick north_star's ">= 10x" is about.  backend="ours": the same classes, unchanged, on this repository's drop-in
packages -- what a user gets by only putting rad-nerf_b200/ on sys.path.  Nothing of radnerf_b200's model / fused renderer / engine
is on either path; synthetic scene, poses, audio windows and batches are the ones bench.py uses for its own arm.

TEST / BENCH INFRASTRUCTURE: the product never imports this module."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "rad-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from baseline import stock   # noqa: E402


def _scene(net, torso):
    """bench.py's synthetic occupancy (make_model): analytic head, torso silhouette"""
    from radnerf_b200 import synthetic as syn
    dev = net.density_bitfield.device
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    with torch.no_grad():
        net.density_grid.copy_(torch.from_numpy(grid).to(dev))
        net.mean_density = float(np.clip(grid, 0, None).mean())
        net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(net.mean_density, net.density_thresh))).to(dev))
        if torso:
            tg = syn.torso_density_grid(128)
            net.density_grid_torso.copy_(torch.from_numpy(tg).to(dev))
            net.mean_density_torso = float(tg.mean())


def frame_rate(dev, hw, steps, warmup=5, backend="ref", torso=True, asr_model="cpierse/wav2vec2-large-xlsr-53-esperanto", dim=44,
               n_frames=8):
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.posemath import convert_poses
    if not stock.available(backend):
        return {"unavailable": "baseline/_ref (reference Python) or oracle/_ref (reference extensions) not installed"}
    net = stock.build(backend, dev, seed=0, torso=torso, asr_model=asr_model, fp16=True).eval()
    _scene(net, torso)
    bank = syn.audio_feature_bank(600, dim, 16, seed=0)
    intr = syn.intrinsics_for(hw, hw)
    bg = torch.from_numpy(syn.get_bg_coords(hw, hw)).to(dev)[None]
    frames = []
    for i in range(n_frames):
        pose = syn.orbit_pose(yaw_deg=10.0 * np.sin(2 * np.pi * i / 64), pitch_deg=2.0)
        ro, rd = syn.get_rays(pose, intr, hw, hw)
        frames.append((torch.from_numpy(ro).to(dev)[None], torch.from_numpy(rd).to(dev)[None], torch.from_numpy(syn.audio_window(bank, 8 + i, 2)).to(dev),
                       convert_poses(torch.from_numpy(pose)[None]).to(dev), torch.tensor([[0.25]], device=dev)))
    kw = vars(net.opt)

    def one(i):
        ro, rd, a, p6, eye = frames[i % len(frames)]
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return net.render(ro, rd, a, bg, p6, eye=eye, index=[0], staged=True, bg_color=None, perturb=False, **kw)["image"]
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": 1000.0 / ms, "unit": "frames/s", "ms_per_step": ms, "steps": steps, "warmup": warmup, "frame": [hw, hw],
            "what": "stock NeRFNetwork.render (baseline/_ref/reference/nerf, unmodified; staged=True, **vars(opt), fp16 autocast) on %s, "
                    "inputs resident, device-timed" % ("the reference's own CUDA extensions (oracle/_ref/*.so, sm_100a build)" if backend == "ref"
                                                         else "this repository's drop-in operator packages (rad-nerf_b200/)")}


def train_rate(dev, n_rays=65536, steps=48, warmup=70, backend="ref"):
    from radnerf_b200 import synthetic as syn
    if not stock.available(backend):
        return {"unavailable": "baseline/_ref (reference Python) or oracle/_ref (reference extensions) not installed"}
    st = stock.load(backend)
    opt = stock.default_opt(torso=False, smooth_lips=False, fp16=True)
    opt.iters = 200000
    torch.manual_seed(0)
    net = st.NeRFNetwork(opt).to(dev).train()
    # the same initial parameters as bench.py's own training leg (radnerf_b200.model mirrors the reference's state-dict names; only
    # the VALUES are borrowed, so that both runs start from the same density field and march comparable sample counts)
    from radnerf_b200.model import NeRFNetwork as _Mirror, Options as _Options
    torch.manual_seed(0)
    net.load_state_dict(_Mirror(_Options(torso=False, fp16=True, exp_eye=True)).state_dict(), strict=False)
    _scene(net, False)
    net.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))       # main.py:210-212
    net.eye_area = torch.full((600, 1), 0.25)
    batches = []
    for i in range(8):
        b = syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i), dev)
        b["images"] = b["rgb"]
        batches.append(b)
    optim = torch.optim.Adam(net.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)            # main.py:204
    sched = torch.optim.lr_scheduler.LambdaLR(optim, lambda it: 0.1 ** min(it / opt.iters, 1))     # main.py:225
    scaler = torch.amp.GradScaler("cuda")
    me = types.SimpleNamespace(opt=opt, model=net, criterion=torch.nn.MSELoss(reduction="none"), global_step=0, flip_finetune_lips=False)
    Trainer = st.utils.Trainer

    def one(i):
        if me.global_step % opt.update_extra_interval == 0 and me.global_step > 0:
            with torch.autocast("cuda", dtype=torch.float16):
                net.update_extra_state()
        me.global_step += 1
        optim.zero_grad()
        with torch.autocast("cuda", dtype=torch.float16):
            _, _, loss = Trainer.train_step(me, batches[i % 8])
        scaler.scale(loss).backward()
        scaler.step(optim)
        scaler.update()
        sched.step()
        return loss.item()
    for i in range(warmup):
        one(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = one(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    samples = float(net.step_counter[:, 0].float().mean())
    return {"ms_per_step": ms, "rays_per_s": n_rays / ms * 1e3, "samples_per_step": samples, "msamples_per_s": samples / ms / 1e3, "steps": steps,
            "warmup": warmup, "loss": float(loss),
            "what": "stock Trainer.train_step + the body of Trainer.train_one_epoch (nerf/utils.py:718-808, 1153-1182, unmodified classes): "
                    "%d rays/batch, fp16 autocast, update_extra_state every 16 steps, torch.optim.Adam + GradScaler + LambdaLR, on %s"
                    % (n_rays, "the reference's own CUDA extensions (oracle/_ref/*.so)" if backend == "ref" else "this repository's drop-in packages")}
