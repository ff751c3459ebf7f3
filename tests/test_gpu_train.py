"""GPU tests of the training ray path (march_rays_train -> network -> composite_rays_train through the C ABI) and of one
full optimisation step as radnerf_b200.train.train_step runs it (the reference's Trainer.train_step, nerf/utils.py:718)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _head_model(seed=0):
    from radnerf_b200.model import NeRFNetwork, Options
    from radnerf_b200 import synthetic as syn
    torch.manual_seed(seed)
    model = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    model.density_grid.copy_(torch.from_numpy(grid))
    model.mean_density = float(np.clip(grid, 0, None).mean())
    model.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(model.mean_density, model.density_thresh))))
    return model.to(DEV)


@pytest.mark.parametrize("tail", ["torch", "fused"])
def test_training_step_has_finite_gradients_everywhere_and_learns(tail):
    """tail: torch.optim.Adam as main.py:204 builds it, or the one-sweep FusedAdam (radnerf_b200.optim) in its place"""
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.optim import FusedAdam
    from radnerf_b200.train import train_step
    model = _head_model()
    batch = syn.batch_to(syn.training_batch(128, 128, 4096, frame_index=3), DEV)
    if tail == "torch":
        opt = torch.optim.Adam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)
    else:
        opt = FusedAdam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
    scaler = torch.amp.GradScaler("cuda")
    names = ("encoder.embeddings", "encoder_ambient.embeddings", "sigma_net.net.0.weight", "color_net.net.1.weight",
             "ambient_net.net.0.weight", "audio_net.encoder_conv.0.weight", "audio_att_net.attentionNet.0.weight",
             "individual_codes")
    before = {n: dict(model.named_parameters())[n].detach().clone() for n in names}
    losses = [float(train_step(model, batch, opt, scaler, lambda_amb=0.1))]
    assert np.isfinite(losses[0])
    # one step touched every trainable group: tables (sparse rows), MLPs, audio nets, the indexed individual code
    for name in names:
        p = dict(model.named_parameters())[name]
        assert p.grad is not None and torch.isfinite(p.grad).all() and torch.isfinite(p).all(), name
        assert not torch.equal(p.detach(), before[name]), name          # the optimiser moved it
        if tail == "torch":
            assert p.grad.abs().sum() > 0, name
        else:
            assert not p.grad.any(), name                                # the sweep left the gradient zeroed for the next step
    counter = model.step_counter[(model.local_step - 1) % 16]
    assert 0 < int(counter[0]) <= 16 * 4096 and 0 < int(counter[1]) <= 4096     # samples / rays the marcher emitted (raymarching.cu:448-452)
    for _ in range(40):
        losses.append(float(train_step(model, batch, opt, scaler, lambda_amb=0.1)))
    assert all(np.isfinite(losses))
    # fixed batch, 40 Adam steps.  The target colours are noise and 70% of the rays only see the fixed background, so the
    # reachable drop is small (0.304 -> 0.283 measured); what matters is that the optimiser moves the loss the right way
    assert losses[-1] < 0.97 * losses[0], losses


def test_training_render_is_deterministic_for_fixed_noise():
    """two training-mode renders of the same batch with perturb=False give identical images and sample counts (the
    marcher's atomics only decide the ORDER of ray slots, composite results are per ray)"""
    from radnerf_b200 import synthetic as syn
    model = _head_model().train()
    b = syn.batch_to(syn.training_batch(128, 128, 2048, frame_index=1), DEV)
    outs = []
    for _ in range(2):
        with torch.autocast("cuda", dtype=torch.float16):
            o = model.render(b["rays_o"], b["rays_d"], b["auds"], b["bg_coords"], b["poses"], eye=b["eye"], index=b["index"],
                             bg_color=b["bg_color"], perturb=False, force_all_rays=False, **model.opt.render_kwargs())
        outs.append((o["image"].detach().clone(), o["weights_sum"].detach().clone(),
                     int(model.step_counter[(model.local_step - 1) % 16][0])))
    assert outs[0][2] == outs[1][2] and outs[0][2] > 0
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_graphed_step_replays_the_whole_step_and_follows_the_schedule():
    """GraphedTrainStep: eager until mean_count is known, then one capture and replays; the optimiser's device-side step
    counters advance once per call, the learning rate reaches the captured sweep through the device group table (lr = 0
    freezes the parameters), a change of mean_count re-captures, the marcher's counter rows are filled as the eager step
    fills them"""
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.optim import FusedAdam
    from radnerf_b200.train import GraphedTrainStep
    model = _head_model()
    batches = [syn.batch_to(syn.training_batch(128, 128, 4096, frame_index=i), DEV) for i in range(3)]
    opt = FusedAdam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
    scaler = torch.amp.GradScaler("cuda")
    step = GraphedTrainStep(model, opt, scaler)
    losses = [float(step(batches[i % 3])) for i in range(4)]           # mean_count unknown: eager
    assert step.captures == 0 and step.replays == 0
    model.mean_count = int(model.step_counter[:4, 0].float().mean().item() * 1.2)
    model.local_step = 0
    losses += [float(step(batches[i % 3])) for i in range(12)]         # 1 eager (warm), 1 capture, 11 replays
    assert step.fallback_reason is None, step.fallback_reason
    assert step.captures == 1 and step.replays == 11
    assert all(np.isfinite(losses))
    p = model.sigma_net.net[0].weight
    assert float(opt.state[p]["step"]) == 16                            # every call took exactly one optimiser step
    assert model.local_step == 12
    rows = model.step_counter[:12, 0]
    assert int(rows.min()) > 0 and len(set(rows.tolist())) > 1          # every step's (samples, rays) row was recorded
    before = {n: q.detach().clone() for n, q in model.named_parameters() if q.grad is not None}
    for g in opt.param_groups:
        g["lr"] = 0.0
    step(batches[0])
    assert step.captures == 1                                           # a new learning rate is not a new graph
    for n, q in model.named_parameters():
        if n in before:
            assert torch.equal(q.detach(), before[n]), n
    for g, lr in zip(opt.param_groups, [5e-4, 5e-3, 5e-3, 5e-4, 5e-4, 5e-4, 2.5e-3, 5e-4]):
        g["lr"] = lr
    step(batches[1])
    assert not torch.equal(p.detach(), before["sigma_net.net.0.weight"])
    # what update_extra_state does every 16 steps: a new mean_count.  The sample budget is a device scalar, the buffers keep their
    # capacity bucket: no new graph, and the marcher honours the new budget exactly as the eager step does
    padded, capacity = step._capacity()
    model.mean_count += 128
    step(batches[2])
    assert step.captures == 1 and step.fallback_reason is None
    assert int(step.budget.item()) == padded + 128
    model.mean_count = int(0.65 * capacity)                             # inside [60 %, 100 %] of the captured capacity: still no new graph
    step(batches[0])
    assert step.captures == 1 and int(step.budget.item()) == model.mean_count + (128 - model.mean_count % 128)
    model.mean_count = 1024                                             # far below: most rays are dropped; the smallest bucket stays
    row = model.local_step % 16
    step(batches[0])
    assert step.captures == 1 and int(step.budget.item()) == 1152
    assert int(model.step_counter[row, 0]) > 1024                       # the counter still counts every sample the rays wanted
    model.mean_count = step.capacity + 1                                # exceeds the captured capacity: one re-capture
    step(batches[1])
    assert step.captures == 2 and step.fallback_reason is None


def test_sample_budget_drops_the_rays_the_reference_budget_drops():
    """march_rays_train under raymarching.sample_budget(capacity, budget): identical (ray, offset, count) table and identical
    samples in the first `budget` slots as the plain call with mean_count = budget, whatever the capacity"""
    import raymarching as rm
    from raymarching.raymarching import sample_budget
    from radnerf_b200 import synthetic as syn
    model = _head_model()
    b = syn.batch_to(syn.training_batch(128, 128, 4096, frame_index=1), DEV)
    ro, rd = b["rays_o"][0].contiguous(), b["rays_d"][0].contiguous()
    nears, fars = rm.near_far_from_aabb(ro, rd, model.aabb_train, model.min_near)

    def march(mean_count, ctx=None):
        counter = torch.zeros(2, dtype=torch.int32, device=DEV)
        args = (ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars, counter, mean_count, False, 128,
                False, model.opt.dt_gamma, model.opt.max_steps)
        if ctx is None:
            return rm.march_rays_train(*args) + (counter,)
        with ctx:
            return rm.march_rays_train(*args) + (counter,)
    full = march(-1)
    total = int(full[4][0])
    budget = (total // 2) // 128 * 128
    ref = march(budget - 128)           # plain call: buffers of _padded(mean_count) = budget slots
    assert ref[0].shape[0] == budget
    lim = torch.tensor([budget], dtype=torch.int32, device=DEV)
    got = march(budget - 128, sample_budget(budget + 4096, lim))
    assert got[0].shape[0] == budget + 4096
    assert torch.equal(got[4], ref[4]) and int(got[4][0]) == total      # the counters count every wanted sample, kept or not

    def check(out, cap):
        # CTAs reserve their slots with one atomic each, so offsets differ from run to run; the rule does not: a ray's samples are
        # written iff the ray ends inside the budget, and nothing else is ever written
        rays = out[3].long()
        off, cnt = rays[:, 1], rays[:, 2]
        kept = (cnt > 0) & (off + cnt <= budget)
        edge = torch.zeros(int((off + cnt).max()) + 2, device=DEV)
        one = torch.ones(int(kept.sum()), device=DEV)
        edge.index_add_(0, off[kept], one)
        edge.index_add_(0, (off + cnt)[kept], -one)
        expected = edge.cumsum(0)[:cap] > 0.5
        assert torch.equal(expected, out[2][:, 0] != 0)
        assert torch.equal(torch.sort(cnt[torch.argsort(rays[:, 0])]).values, torch.sort(full[3][:, 2].long()).values)
        return int(kept.sum())
    assert check(ref, budget) > 0 and check(got, budget + 4096) > 0
    assert not got[2][budget:].any() and not got[0][budget:].any()
