"""Driver for `ncu -k regex:head_eval`: renders ONE fixed 512x512 frame repeatedly with direct launches (no CUDA graph), so
that head_eval launch #(16*r + it) is loop iteration `it` of repetition r, and prints the sample count of every iteration
(needed to turn ncu's per-launch DRAM bytes into bytes per sample).  Usage: python tools/profile_head_eval.py [reps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
sys.argv = sys.argv[:1]
import bench
from radnerf_b200 import frame, synthetic as syn

dev = torch.device("cuda")
model = bench.make_model(dev)
frames, intr, bg = bench.make_frames(512, 1)
bg_t = torch.from_numpy(bg).to(dev)[None]
kw = model.opt.render_kwargs()
f = frames[0]
ro, rd = syn.get_rays(f["pose"], intr, 512, 512)
d = dict(ro=torch.from_numpy(ro).to(dev)[None], rd=torch.from_numpy(rd).to(dev)[None], auds=torch.from_numpy(f["auds"]).to(dev),
         pose6=torch.from_numpy(f["pose6"]).to(dev), eye=torch.from_numpy(f["eye"]).to(dev))
from radnerf_b200 import abi
abi.lib()
model._fused = frame.FusedState(model)
model._fused.use_graph = False
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    for r in range(reps):
        model.render(d["ro"], d["rd"], d["auds"], bg_t, d["pose6"], eye=d["eye"], index=0, path="fused", **kw)
torch.cuda.synchronize()
print(json.dumps({"max_steps": kw["max_steps"], "schedule_n_alive_n_step_n_samples": frame.frame_stats(model)}))
