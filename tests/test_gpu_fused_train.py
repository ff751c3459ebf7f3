"""The fused training step of the head network (radnerf_b200.fused_train: one forward kernel, two tcgen05 backward kernels) against
the op-by-op autograd path -- which is itself pinned to the reference's own classes (tests/test_train_parity.py) and whose operators
are bit-checked against the reference kernels (tests/test_gpu_parity.py, tests/test_stock_reference.py).

Yardsticks: `ops fp32` = the op-by-op step in fp32 (the truth both fp16 paths approximate); `ops fp16` = the op-by-op step under fp16
autocast, i.e. what the reference computes with -O.  The fused step must be as close to the truth as the reference's own fp16 step is."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SCALE = 1024.0     # stands in for GradScaler's loss scale
_report = {}


def _model(seed=0):
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200 import synthetic as syn
    torch.manual_seed(seed)
    model = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    model.density_grid.copy_(torch.from_numpy(grid))
    model.mean_density = float(np.clip(grid, 0, None).mean())
    model.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(model.mean_density, model.density_thresh))))
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # "trained-like" tables: every stage numerically visible
        for enc in (model.encoder, model.encoder_ambient):
            enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1) * 0.5)
    return model.to(DEV).train()


def _step_grads(model, batch, fused, amp):
    from radnerf_b200.train import head_loss
    model.fused_train = "auto" if fused else False
    model.zero_grad(set_to_none=True)
    model.local_step = 0
    with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
        out = model.render(batch["rays_o"], batch["rays_d"], batch["auds"], batch["bg_coords"], batch["poses"], eye=batch["eye"],
                           index=batch["index"], bg_color=batch["bg_color"], perturb=False, force_all_rays=True, **model.opt.render_kwargs())
        loss = head_loss(out, batch["rgb"], batch["face_mask"], 0.1)
    (loss * SCALE).backward()
    grads = {n: (p.grad.detach().double() / SCALE) for n, p in model.named_parameters() if p.grad is not None}
    return float(loss), {k: v.detach() for k, v in out.items()}, grads


def test_fused_head_forward_matches_the_op_by_op_network():
    from radnerf_b200 import fused_train, synthetic as syn
    import raymarching as rm
    model = _model()
    b = syn.batch_to(syn.training_batch(128, 128, 4096, frame_index=2), DEV)
    ro, rd = b["rays_o"][0].contiguous(), b["rays_d"][0].contiguous()
    nears, fars = rm.near_far_from_aabb(ro, rd, model.aabb_train, model.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=DEV)
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars,
                                                   counter, -1, False, 128, True, model.opt.dt_gamma, model.opt.max_steps)
    assert xyzs.shape[0] % 128 == 0 and int(counter[0]) > 5000
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        enc_a = model.encode_audio(b["auds"])
        ind = model.individual_codes[b["index"]]
        s_ref, c_ref, a_ref = model(xyzs, dirs, enc_a, ind, b["eye"])
        s, c, a = fused_train.head_forward(model, xyzs, dirs, enc_a, ind, b["eye"])
        # a ragged batch (not a multiple of the 128-row tile) gives the same rows
        s2, c2, a2 = fused_train.head_forward(model, xyzs[:1000], dirs[:1000], enc_a, ind, b["eye"])
        dens = fused_train.density(model, xyzs, enc_a, b["eye"])
    torch.cuda.synchronize()
    n = int(counter[0])
    rel_s = ((s[:n] - s_ref[:n].float()).abs() / s_ref[:n].float().abs().clamp(min=1e-3)).max().item()
    d_c = (c[:n] - c_ref[:n].float()).abs().max().item()
    d_a = (a[:n] - a_ref[:n].float()).abs().max().item()
    _report["forward"] = dict(rel_sigma=rel_s, d_color=d_c, d_ambient=d_a, samples=n)
    print(_report["forward"])
    assert rel_s <= 4e-3 and d_c <= 2e-3 and d_a <= 2e-3
    assert torch.equal(s2, s[:1000]) and torch.equal(c2, c[:1000]) and torch.equal(a2, a[:1000])
    assert torch.equal(dens, s)


@pytest.mark.parametrize("n_rays", [4096, 16384])
def test_fused_training_step_gradients(n_rays):
    """loss, image and every gradient of one head training step: fused vs op-by-op fp16 vs op-by-op fp32"""
    from radnerf_b200 import synthetic as syn
    model = _model(seed=3)
    b = syn.batch_to(syn.training_batch(192, 192, n_rays, frame_index=5), DEV)
    l32, o32, g32 = _step_grads(model, b, fused=False, amp=False)
    l16, o16, g16 = _step_grads(model, b, fused=False, amp=True)
    lf, of, gf = _step_grads(model, b, fused=True, amp=True)
    assert set(gf) == set(g16) == set(g32)
    rows = {}
    for n in sorted(g32):
        ref = g32[n]
        nr = ref.norm().item()
        rows[n] = dict(norm=nr, rel_fused=(gf[n] - ref).norm().item() / max(nr, 1e-30), rel_ops16=(g16[n] - ref).norm().item() / max(nr, 1e-30),
                       max_fused=(gf[n] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30))
    _report["grads_%d" % n_rays] = dict(loss32=l32, loss16=l16, loss_fused=lf, d_image_fused=(of["image"] - o32["image"]).abs().max().item(),
                                        d_image_ops16=(o16["image"] - o32["image"]).abs().max().item(), rows=rows)
    for n, r in rows.items():
        print("%-40s |g|=%.3e  fused %.2e  ops16 %.2e  (max-norm fused %.2e)" % (n, r["norm"], r["rel_fused"], r["rel_ops16"], r["max_fused"]))
    assert abs(lf - l32) <= 2e-3 * max(1.0, abs(l32)) and (of["image"] - o32["image"]).abs().max().item() <= 2e-3
    for n, r in rows.items():
        # as close to the fp32 step as the reference's own fp16 arithmetic gets (x2 head-room), never worse than 3 % of the norm
        assert r["rel_fused"] <= max(2.0 * r["rel_ops16"], 3e-2), (n, r)


def test_fused_training_learns_like_the_op_by_op_step():
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.optim import FusedAdam
    from radnerf_b200.train import train_step
    curves = {}
    for fused in (False, True):
        model = _model(seed=7)
        model.fused_train = "auto" if fused else False
        batch = syn.batch_to(syn.training_batch(128, 128, 4096, frame_index=3), DEV)
        opt = FusedAdam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
        scaler = torch.amp.GradScaler("cuda")
        torch.manual_seed(11)
        curves[fused] = [float(train_step(model, batch, opt, scaler, lambda_amb=0.1)) for _ in range(60)]
    a, b = np.array(curves[False]), np.array(curves[True])
    _report["learning"] = dict(ops=curves[False][::10], fused=curves[True][::10])
    print(a[::10], b[::10])
    assert np.isfinite(b).all() and b[-1] < 0.92 * b[0]
    assert abs(b[-1] - a[-1]) <= 0.01 * a[0]        # same optimisation trajectory up to fp16 noise (measured: 3e-4 after 60 steps)


def test_zz_write_report():
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    json.dump(_report, open(os.path.join(out, "fused_train_parity.json"), "w"), indent=1, sort_keys=True)
