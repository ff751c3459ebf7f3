"""One steady-state training step between cudaProfilerStart/Stop, for a launch list:

    python tools/train_profile.py                                   # must exit 0 on its own first
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/train_launches.csv python tools/train_profile.py
    python tools/launch_summary.py gpurun_out/train_launches.csv profiles/rNN_launches_train_step_summary.txt "<note>"
    # one kernel in depth, e.g. the table backward (3-D: L2-atomic bound at 6 % of HBM peak in round 1):
    ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:grid_backward \
        -o gpurun_out/prof_grid_backward_rNN python tools/train_profile.py
    python tools/ncu_summary.py gpurun_out/prof_grid_backward_rNN.ncu-rep profiles/rNN_grid_backward_ncu_full.json

Setup as tools/train_bench.py (BASELINE configs[3]: 2^16 rays, fp16 autocast, FusedAdam tail): 16 cold steps, the first
occupancy update, 4 steady steps, then the bracketed step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
from radnerf_b200 import synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options
from radnerf_b200.optim import FusedAdam
from radnerf_b200.train import train_step

dev = torch.device("cuda", 0)
n_rays = int(os.environ.get("N_RAYS", 65536))
torch.manual_seed(0)
m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
m.density_grid.copy_(torch.from_numpy(grid))
m.mean_density = float(np.clip(grid, 0, None).mean())
m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
m = m.to(dev)
m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))
m.eye_area = torch.full((600, 1), 0.25)
batches = [syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=i), dev) for i in range(4)]
opt = FusedAdam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
scaler = torch.amp.GradScaler("cuda")
for i in range(20):
    if i == 16:
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    train_step(m, batches[i % 4], opt, scaler, None)
torch.cuda.synchronize()
torch.cuda.profiler.start()
loss = train_step(m, batches[0], opt, scaler, None)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("samples", int(m.step_counter[(m.local_step - 1) % 16][0]), "mean_count", m.mean_count, "loss", float(loss))
