"""Per-stage device timing of the fused frame (CUDA events on the launching stream). Usage: python tools/frame_breakdown.py [hw]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
sys.argv = sys.argv[:1] + sys.argv[1:]
import bench
from radnerf_b200 import frame, abi, synthetic as syn

hw = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda")
model = bench.make_model(dev)
frames, intr, bg = bench.make_frames(hw, 8)
bg_t = torch.from_numpy(bg).to(dev)[None]
kw = model.opt.render_kwargs()
devf = []
for f in frames:
    ro, rd = syn.get_rays(f["pose"], intr, hw, hw)
    devf.append(dict(ro=torch.from_numpy(ro).to(dev)[None], rd=torch.from_numpy(rd).to(dev)[None], auds=torch.from_numpy(f["auds"]).to(dev),
                     pose6=torch.from_numpy(f["pose6"]).to(dev), eye=torch.from_numpy(f["eye"]).to(dev)))

def render(i):
    f = devf[i % len(devf)]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        return model.render(f["ro"], f["rd"], f["auds"], bg_t, f["pose6"], eye=f["eye"], index=0, path="fused", **kw)

for i in range(5): render(i)
torch.cuda.synchronize()
def time_frames(n=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for i in range(n): render(i)
    e1.record(); t_host = time.perf_counter() - t0; torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_host / n * 1e3
import time
print("CUDA-graph replay : %.3f ms device/frame, host issue %.3f ms/frame" % time_frames())
model._fused.use_graph = False
for i in range(3): render(i)
torch.cuda.synchronize()
print("direct launches   : %.3f ms device/frame, host issue %.3f ms/frame" % time_frames())
print("schedule (n_alive, n_step, n_samples):", frame.frame_stats(model))
# stage timing by monkeypatching the library calls with event brackets
L = abi.lib()
stages = {}
def wrap(name):
    fn = getattr(L, name)
    def w(*a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); rc = fn(*a); e.record()
        stages.setdefault(name, []).append((s, e))
        return rc
    return w
class Proxy:
    def __getattr__(self, n):
        if n in ("rn_frame_conditioning", "rn_frame_head", "rn_frame_torso", "rn_frame_finalize"):
            return wrap(n)
        return getattr(L, n)
abi._lib_real = L
frame.abi = type("A", (), {k: getattr(abi, k) for k in dir(abi)})
frame.abi.lib = staticmethod(lambda: Proxy())
for i in range(20): render(i)
torch.cuda.synchronize()
for k, v in stages.items():
    print("%-24s %.3f ms" % (k, sum(s.elapsed_time(e) for s, e in v) / len(v)))

# ---- per-phase cycle counters inside head_eval (thread 0 of every tile group)
Lr = abi._lib_real
Lr.rn_debug_set_head_prof.argtypes = [C.c_void_p]
prof = torch.zeros(8, dtype=torch.int64, device=dev)
Lr.rn_debug_set_head_prof(prof.data_ptr())
frame.abi = abi
render(0); torch.cuda.synchronize()
Lr.rn_debug_set_head_prof(None)
p = prof.cpu().numpy()
tot = p[4]
print("head_eval phase cycles (sum over %d group-runs): enc3 %.1f%% enc2 %.1f%% mma-wait %.1f%% epilogue(4-chunk) %.1f%% other %.1f%%; cycles/tile/group = %.0f" % (
    p[5], 100*p[0]/tot, 100*p[1]/tot, 100*p[2]/tot, 100*p[3]/tot, 100*(tot-p[0]-p[1]-p[2]-p[3])/tot, tot / (815354/128)))
print("head_eval CTA set-up (kernel entry -> first tile; barriers, TMA issue, TMEM alloc, level tables): %.0f cycles = %.2f us per group-run" % (
    p[6] / max(1, p[5]), p[6] / max(1, p[5]) / 1965.0))

# ---- conditioning kernel alone: 20 launches inside one CUDA graph -> pure device time
st = model._fused
f0 = devf[0]
cd = frame.conditioning_desc(model, st, f0["auds"].contiguous(), f0["eye"].reshape(-1).contiguous(), f0["pose6"].reshape(-1).contiguous())
saved = st.enc_a_state.clone()
Lr.rn_frame_conditioning(C.byref(cd), abi.cur_stream()); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20): Lr.rn_frame_conditioning(C.byref(cd), abi.cur_stream())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
st.enc_a_state.copy_(saved)
print("audio_frame_kernel alone: %.1f us" % (e0.elapsed_time(e1) / 20 * 1e3))
aprof = torch.zeros(16, dtype=torch.int64, device=dev)
Lr.rn_debug_set_audio_prof.argtypes = [C.c_void_p]
Lr.rn_debug_set_audio_prof(aprof.data_ptr())
Lr.rn_frame_conditioning(C.byref(cd), abi.cur_stream()); torch.cuda.synchronize()
Lr.rn_debug_set_audio_prof(None)
st.enc_a_state.copy_(saved)
ap = aprof.cpu().numpy()
print("audio kernel stage cycles [load, conv1, conv2-4, fc, att-convs, att-fc+softmax+sum+smooth, head hoists, torso hoists]:", [int(ap[i+1]-ap[i]) for i in range(8)])
from radnerf_b200 import roofline
hbm, tf, _ = roofline.peaks()
for e in frame.roofline_entries(model, devf[0], bg_t, kw, hbm, tf):
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items()})
