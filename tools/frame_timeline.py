"""Kernel-by-kernel timeline of ONE fused frame replayed from its CUDA graph (torch.profiler / CUPTI): duration of every kernel and the
gap in front of it -- where the latency of a small (ray-sharded) frame goes.

    python tools/frame_timeline.py [hw=184] [frames=20]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from radnerf_b200.stream import FrameStreamer, pack_inputs

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 184
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
model = bench.make_model(dev)
frames, intr, bg = bench.make_frames(hw, 4)
bg_t = torch.from_numpy(bg).to(dev)
kw = model.opt.render_kwargs()
packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]).to(dev) for f in frames]
st = FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, use_eye=True, deliver=False, depth=1, **kw)
for i in range(12):
    st.submit(packed[i % 4]); st.collect()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(n):
        st.submit(packed[i % 4]); st.collect()
    torch.cuda.synchronize()
ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
# split into frames at the ray-generation kernel
rows, cur = [], []
for e in ev:
    if "get_rays" in e.name and cur:
        rows.append(cur); cur = []
    cur.append(e)
rows.append(cur)
rows = [r for r in rows if len(r) == len(rows[len(rows) // 2])]
k = len(rows[0])
lines = ["one %dx%d frame, one frame in flight, graph replay: %d device activities per frame (mean over %d frames)" % (hw, hw, k, len(rows)),
         "%-72s %9s %9s" % ("activity", "gap us", "dur us")]
tot_gap = tot_dur = 0.0
for j in range(k):
    gap = sum(max(0.0, r[j].time_range.start - max(x.time_range.end for x in r[:j])) if j else 0.0 for r in rows) / len(rows)
    dur = sum(r[j].time_range.end - r[j].time_range.start for r in rows) / len(rows)
    tot_gap += gap; tot_dur += dur
    lines.append("%-72s %9.2f %9.2f" % (rows[0][j].name.replace("rn::(anonymous namespace)::", "")[:72], gap, dur))
span = sum(max(x.time_range.end for x in r) - r[0].time_range.start for r in rows) / len(rows)
lines.append("frame span %.1f us; sum of durations %.1f us (activities on side streams overlap), sum of gaps on the critical order %.1f us" % (span, tot_dur, tot_gap))
out = "\n".join(lines)
print(out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "frame_timeline_%d.txt" % hw), "w").write(out + "\n")
