"""Is a small (ray-sharded-size) frame bound by the host's submission rate, by the per-frame dependency chain, or by the GPU?

    python tools/lane_probe.py [hw ...]      # one GPU; frames of hw x hw rays stand in for one rank's share of a 512x512 frame

For every frame size and lane count: resident frames/s over 400 frames, host time per submit() call and per collect() call."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
import bench
from radnerf_b200.stream import FrameStreamer, pack_inputs

dev = torch.device("cuda", 0)
sizes = [int(a) for a in sys.argv[1:]] or [512, 256, 184, 128]
out = []
for hw in sizes:
    model = bench.make_model(dev)
    frames, intr, bg = bench.make_frames(hw, 16)
    bg_t = torch.from_numpy(bg).to(dev)
    kw = model.opt.render_kwargs()
    packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]).to(dev) for f in frames]
    for lanes in (1, 2, 4, 8):
        model.enc_a = None
        st = FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, use_eye=True, deliver=False, depth=lanes, **kw)
        for i in range(3 * lanes + 8):
            if st.in_flight() == st.depth:
                st.collect()
            st.submit(packed[i % 16])
        while st.in_flight():
            st.collect()
        st.sync()
        torch.cuda.synchronize()
        n = 400
        t_sub = t_col = 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(n):
            if st.in_flight() == st.depth:
                a = time.perf_counter()
                st.collect()
                t_col += time.perf_counter() - a
            a = time.perf_counter()
            st.submit(packed[i % 16])
            t_sub += time.perf_counter() - a
        while st.in_flight():
            st.collect()
        st.sync()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        r = {"hw": hw, "rays": hw * hw, "lanes": lanes, "frames_per_s": n / ms * 1e3, "ms_per_frame": ms / n, "host_submit_us": t_sub / n * 1e6,
             "host_collect_wait_us": t_col / n * 1e6, "wall_ms_per_frame": wall / n * 1e3}
        print(json.dumps(r), flush=True)
        out.append(r)
        st.close()
        del st
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "lane_probe.json"), "w"), indent=1)
