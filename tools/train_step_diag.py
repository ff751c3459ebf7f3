"""Per-step GPU time of the graphed training step in bench.py's regime (70 warm-up steps, occupancy update every 16 steps): one CUDA event
per step, then replay-only loops without updates -- where the difference between a replay and the bench's ms/step goes."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch, numpy as np
from radnerf_b200 import synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options
from radnerf_b200.optim import FusedAdam
from radnerf_b200.train import GraphedTrainStep
dev=torch.device("cuda",0)
torch.manual_seed(0)
m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
m.density_grid.copy_(torch.from_numpy(grid)); m.mean_density = float(np.clip(grid, 0, None).mean())
m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
m = m.to(dev)
m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0)); m.eye_area = torch.full((600, 1), 0.25)
batches = [syn.batch_to(syn.training_batch(512, 512, 65536, frame_index=i, seed=0), dev) for i in range(8)]
opt = FusedAdam(m.get_params(5e-3, 5e-4), betas=(0.9, 0.99), eps=1e-15, zero_grads=True)
scaler = torch.amp.GradScaler("cuda")
graphed = GraphedTrainStep(m, opt, scaler)
g=[0]
def one(i, upd=True):
    if upd and g[0] % 16 == 0 and g[0] > 0:
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    g[0]+=1
    return graphed(batches[i % 8])
for i in range(70): one(i)
torch.cuda.synchronize()
ev=[torch.cuda.Event(enable_timing=True) for _ in range(65)]
host=[]
ev[0].record()
for i in range(64):
    one(i); ev[i+1].record(); host.append(time.perf_counter())
torch.cuda.synchronize()
d=[ev[i].elapsed_time(ev[i+1]) for i in range(64)]
print("per-step GPU ms:", " ".join("%.2f"%x for x in d))
print("mean %.3f median %.3f; samples/step %.0f; captures %d" % (np.mean(d), np.median(d), float(m.step_counter[:,0].float().mean()), graphed.captures))
# replay only, same batch
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(32): one(0, upd=False)
e1.record(); torch.cuda.synchronize()
print("32 replays of batch 0 without updates: %.3f ms/step" % (e0.elapsed_time(e1)/32))
e0.record()
for i in range(32): one(i, upd=False)
e1.record(); torch.cuda.synchronize()
print("32 replays cycling 8 batches without updates: %.3f ms/step" % (e0.elapsed_time(e1)/32))
