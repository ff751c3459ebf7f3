"""How far can ray sharding scale?  Renders rank 0's share of a 512x512 frame for world = 1, 2, 4, 8 on ONE GPU (no
collective), CUDA-graph replay, device-timed: the per-rank compute floor of the strong-scaling curve."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
sys.argv = sys.argv[:1]
import bench
from radnerf_b200 import frame, synthetic as syn
from radnerf_b200.sharding import FrameSharder
dev = torch.device("cuda")
hw = int(os.environ.get("HW", 512))
frames, intr, bg = bench.make_frames(hw, 8)
for world in (1, 2, 4, 8):
    model = bench.make_model(dev)
    kw = model.opt.render_kwargs()
    sh = FrameSharder(hw, hw, world, 0, dev)
    bg_l = sh.shard(torch.from_numpy(bg).to(dev))[None]
    devf = []
    for f in frames:
        ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
        devf.append(dict(ro=sh.shard(torch.from_numpy(ro).to(dev))[None], rd=sh.shard(torch.from_numpy(rd).to(dev))[None],
                         auds=torch.from_numpy(f["auds"]).to(dev), pose6=torch.from_numpy(f["pose6"]).to(dev), eye=torch.from_numpy(f["eye"]).to(dev)))
    def render(i):
        f = devf[i % len(devf)]
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            return model.render(f["ro"], f["rd"], f["auds"], bg_l, f["pose6"], eye=f["eye"], index=0, path="fused", **kw)
    for i in range(10): render(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): render(i)
    e1.record(); torch.cuda.synchronize()
    print(json.dumps({"hw": hw, "world": world, "rays_per_rank": hw * hw // world, "ms_per_frame_rank0": e0.elapsed_time(e1) / 200,
                      "schedule": frame.frame_stats(model)}))
    if os.environ.get("KERNELS"):
        from radnerf_b200 import roofline
        hbm, tfl, _ = roofline.peaks()
        for e in frame.roofline_entries(model, devf[0], bg_l, kw, hbm, tfl):
            print("   ", e["kernel"][:28], {k: (round(float(v), 4) if isinstance(v, (int, float)) or hasattr(v, "item") else v) for k, v in e.items() if k in ("ms", "ms_per_frame", "units", "launches_per_frame")})
