"""Where a steady-state training step spends its device and host time (torch.profiler / CUPTI; no ncu needed).

    python tools/train_timeline.py [steps]      # writes gpurun_out/train_timeline.txt

Same setup as tools/train_profile.py.  Prints per-kernel device totals over `steps` steps, the host wall time of the loop, the
number of launches per step, and the cost of one occupancy update (`update_extra_state`) measured by CUDA events."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile
from radnerf_b200 import synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options
from radnerf_b200.optim import FusedAdam
from radnerf_b200.train import train_step

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
n_rays = int(os.environ.get("N_RAYS", 65536))
torch.manual_seed(0)
m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
m.density_grid.copy_(torch.from_numpy(grid))
m.mean_density = float(np.clip(grid, 0, None).mean())
m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
m = m.to(dev)
if os.environ.get("RADNERF_FUSED_TRAIN", "1") == "0":
    m.fused_train = False
m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))
m.eye_area = torch.full((600, 1), 0.25)
batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i), dev) for i in range(4)]
opt = FusedAdam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
scaler = torch.amp.GradScaler("cuda")
graphed = None
if os.environ.get("GRAPHED", "0") == "1":       # the bench's regime: replayed from a CUDA graph, warmed up through four occupancy updates
    from radnerf_b200.train import GraphedTrainStep
    graphed = GraphedTrainStep(m, opt, scaler)
    _eager_step = train_step

    def train_step(model, batch, o, sc, sync):   # noqa: F811
        return graphed(batch)
for i in range(int(os.environ.get('WARM', 70 if graphed is not None else 20))):
    if i % 16 == 0 and i > 0:
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    train_step(m, batches[i % 4], opt, scaler, None)
torch.cuda.synchronize()

# plain timing first (no profiler attached)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(steps):
    train_step(m, batches[i % 4], opt, scaler, None)
e1.record()
host_issue_ms = (time.perf_counter() - t0) * 1e3 / steps
torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / steps

u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
u0.record()
with torch.autocast("cuda", dtype=torch.float16):
    m.update_extra_state()
u1.record()
upd_host_ms = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize()
upd_ms = u0.elapsed_time(u1)

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        train_step(m, batches[i % 4], opt, scaler, None)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in ev:
    k = e.name[:110]
    a = tot.setdefault(k, [0.0, 0])
    a[0] += e.device_time if hasattr(e, "device_time") else e.cuda_time
    a[1] += 1
lines = ["training step, %d rays, fused_train=%s: %.3f ms/step on the device clock, host issue %.3f ms/step; update_extra_state %.2f ms device "
         "(%.2f ms host)" % (n_rays, m.fused_train, dev_ms, host_issue_ms, upd_ms, upd_host_ms),
         "device activities per step: %.1f, summed device time per step %.3f ms" % (len(ev) / steps, sum(a[0] for a in tot.values()) / steps / 1e3),
         "%-112s %10s %8s" % ("kernel / memcpy", "us/step", "n/step")]
for k, a in sorted(tot.items(), key=lambda kv: -kv[1][0])[:60]:
    lines.append("%-112s %10.1f %8.1f" % (k, a[0] / steps, a[1] / steps))
out = "\n".join(lines)
print(out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "train_timeline%s%s.txt" % ("" if m.fused_train else "_ops", "_graphed" if graphed is not None else "")), "w").write(out + "\n")
