"""Data-parallel training check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_train_check.py
Every rank trains the same replica on its own ray batch with radnerf_b200.train (GradSync over NCCL).  Checks that after
every step the parameters of all ranks are bit-identical, that the synced gradient equals the mean of the per-rank gradients
(recomputed without sync), and prints device-timed steps/s plus the bytes each rank all-reduces per step."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
import torch.distributed as dist
from radnerf_b200 import synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options
from radnerf_b200.train import GradSync, train_step

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_rays = int(os.environ.get("N_RAYS", 65536))
steps = int(os.environ.get("STEPS", 20))


def make():
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    return m.to(dev)


model = make()
sync = GradSync(model.parameters()) if world > 1 else None
if os.environ.get("TAIL", "torch") == "fused":     # the one-sweep tail (radnerf_b200.optim): deterministic, replicas stay identical
    from radnerf_b200.optim import FusedAdam
    opt = FusedAdam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
else:
    opt = torch.optim.Adam(model.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15)
scaler = torch.amp.GradScaler("cuda")
batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i, seed=rank), dev) for i in range(4)]

# ---- gradient check on step 0 (no optimiser step): synced grad == mean over ranks of local grads
ok_grad = True
if world > 1:
    model.train()
    def local_grads(m, b):
        m.zero_grad(set_to_none=True)
        m.local_step = 0
        with torch.autocast("cuda", dtype=torch.float16):
            out = m.render(b["rays_o"], b["rays_d"], b["auds"], b["bg_coords"], b["poses"], eye=b["eye"], index=b["index"],
                           bg_color=b["bg_color"], perturb=False, force_all_rays=False, **m.opt.render_kwargs())
            loss = ((out["image"] - b["rgb"]) ** 2).mean()
        loss.backward()
    sync.remove()                      # local pass without hooks
    local_grads(model, batches[0])
    mine = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    mean = {}
    for n, g in mine.items():
        t = g.clone(); dist.all_reduce(t); mean[n] = t / world
    sync = GradSync(model.parameters())
    local_grads(model, batches[0]); sync.finish()
    for n, p in model.named_parameters():
        if p.grad is not None:
            # table gradients are accumulated with float atomics: run-to-run order differs, so compare with a tolerance
            err = (p.grad - mean[n]).abs().max().item(); ref = mean[n].abs().max().item() + 1e-12
            if err > 2e-3 * ref + 1e-7:
                ok_grad = False
                print(f"[rank {rank}] grad mismatch {n}: {err:.3e} vs scale {ref:.3e}")
    model.zero_grad(set_to_none=True)

# steady state as in a real run (tools/train_bench.py): 16 steps with unknown mean_count, the first occupancy update -- run so
# that the replicas stay identical (same seed on every rank, max-reduced counters) -- then the timed steps
from radnerf_b200.train import update_extra_state_replicated
model.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))
model.eye_area = torch.full((600, 1), 0.25)
for i in range(16 if os.environ.get("REGIME", "steady") == "steady" else 3):
    train_step(model, batches[i % 4], opt, scaler, sync)
if os.environ.get("REGIME", "steady") == "steady":
    with torch.autocast("cuda", dtype=torch.float16):
        update_extra_state_replicated(model)
    for i in range(4):
        train_step(model, batches[i % 4], opt, scaler, sync)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    loss = train_step(model, batches[i % 4], opt, scaler, sync)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# replicas must have stayed identical
same = True
if world > 1:
    for n, p in model.named_parameters():
        lo, hi = p.detach().clone(), p.detach().clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if not torch.equal(lo, hi):
            same = False
            if rank == 0: print("replicas diverged in", n)
if rank == 0:
    print(json.dumps({"world": world, "rays_per_rank": n_rays, "regime": os.environ.get("REGIME", "steady"), "mean_count": int(model.mean_count), "ms_per_step": float(ms.item()),
                      "rays_per_s_total": world * n_rays / (float(ms.item()) / 1e3), "loss": float(loss),
                      "allreduce_bytes_per_step": None if sync is None else sync.bytes_last,
                      "synced_grad_equals_mean_of_local": ok_grad, "replicas_identical_after_training": same}))
if world > 1:
    dist.destroy_process_group()
