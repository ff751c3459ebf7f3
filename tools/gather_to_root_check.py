"""torchrun --nproc-per-node N tools/gather_to_root_check.py : the flag-based gather-to-root (csrc/peer_gather.cu: rn_scatter_rows_to_root +
rn_stage_frame_at_root, arrival counter / consumed flags instead of a barrier).

1. raw exchange: 400 synthetic frames through every frame buffer, each rank's rows carrying (frame, rank)-dependent values, ranks
   deliberately skewed (one rank sleeps on its stream every few frames): the root must stage exactly frame i every time -- a torn frame
   (write-after-read across frames in flight, the hazard the `consumed` flag closes) or a missing row shows up as a mismatch;
2. streamed frames: FrameStreamer with the ray-sharded frame, `lanes` frames in flight, fp32 and uint8 delivery, against the same frames
   rendered UNSHARDED on the root (bit-identical images expected: sharding only re-orders rays);
3. timing of the exchange alone (device events, max over ranks) next to the barrier-based all-to-all scatter and NCCL all-gather."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
import torch.distributed as dist
from radnerf_b200.sharding import FrameSharder

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H = W = int(os.environ.get("HW", 512))
lanes = int(os.environ.get("LANES", 4))
res = {"world": world, "lanes": lanes}

# ---- 1. raw exchange under skew
sh = FrameSharder(H, W, world, rank, dev)
assert sh.enable_peer_gather(n_buffers=lanes) and sh.ctrl is not None, getattr(sh, "peer_error", "no control block")
base = torch.rand(H * W, 3, generator=torch.Generator().manual_seed(0)).to(dev)
mine = sh.shard(base)
stage = [torch.zeros(H * W, 3, device=dev) for _ in range(lanes)]
streams = [torch.cuda.Stream() for _ in range(lanes)]
bad = 0
checks = []
for i in range(400):
    k = i % lanes
    with torch.cuda.stream(streams[k]):
        if (i // 7) % world == rank and i % 7 == 0:
            torch.cuda._sleep(2_000_000)           # ~1 ms of skew on one rank at a time
        loc = mine * float(i + 1)
        sh.gather_to_root(loc, slot=k, stage_to=stage[k] if rank == 0 else None)
        if rank == 0:
            checks.append((i, (stage[k] - base * float(i + 1)).abs().max()))
torch.cuda.synchronize()
if rank == 0:
    bad = sum(1 for _, d in checks if float(d) != 0.0)
res["raw_frames"] = 400
res["raw_mismatches"] = bad

# ---- 2. streamed frames vs the unsharded render on the root
import bench
from radnerf_b200.stream import FrameStreamer, pack_inputs
hw = 256
model = bench.make_model(dev)
frames, intr, bg = bench.make_frames(hw, 12)
bg_t = torch.from_numpy(bg).to(dev)
kw = model.opt.render_kwargs()
packed = [pack_inputs(f["pose"], f["auds"], f["pose6"], f["eye"]) for f in frames]
n_frames = 48
stream_ok = {}
for output in ("float32", "uint8"):
    ref = None
    if rank == 0:
        model.enc_a = None
        st1 = FrameStreamer(model, hw, hw, intr, bg_t, frames[0]["auds"].shape, use_eye=True, depth=2, output=output, **kw)
        ref = [img.clone() for img in st1.render_all([packed[i % 12] for i in range(n_frames)])]
        st1.close()
    sh2 = FrameSharder(hw, hw, world, rank, dev)
    assert sh2.enable_peer_gather(n_buffers=lanes)
    model.enc_a = None
    stN = FrameStreamer(model, hw, hw, intr, sh2.shard(bg_t), frames[0]["auds"].shape, use_eye=True, sharder=sh2, deliver=(rank == 0), depth=lanes,
                        output=output, **kw)
    got = []
    for i in range(n_frames):
        if stN.in_flight() == stN.depth:
            img = stN.collect()
            if rank == 0:
                got.append(img.clone())
        stN.submit(packed[i % 12])
    while stN.in_flight():
        img = stN.collect()
        if rank == 0:
            got.append(img.clone())
    stN.sync()
    torch.cuda.synchronize()
    if rank == 0:
        worst = max(float((a.float() - b.float()).abs().max()) for a, b in zip(got, ref))
        stream_ok[output] = {"frames": len(got), "max_abs_diff_vs_unsharded": worst, "fast_path": any(x is not None for x in stN.fast)}
    stN.close()
res["streamed"] = stream_ok


# ---- 3. the exchange alone
def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) * 1e3


loc = mine.clone()
res["us_gather_to_root_staged"] = timeit(lambda: sh.gather_to_root(loc, slot=0, stage_to=stage[0] if rank == 0 else None))
res["us_gather_to_root"] = timeit(lambda: sh.gather_to_root(loc, slot=0))
res["us_all_to_all_scatter_plus_barrier"] = timeit(lambda: sh.gather(loc, slot=0))
peer = sh.peer
sh.peer = None
res["us_nccl_all_gather_unpermute"] = timeit(lambda: sh.gather(loc))
sh.peer = peer
flag = torch.tensor([1 if (bad == 0 and all(v["max_abs_diff_vs_unsharded"] == 0.0 for v in stream_ok.values())) else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["ok"] = bool(flag.item())
if rank == 0:
    print(json.dumps(res))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "gather_to_root_%dgpu.json" % world), "w"), indent=1)
dist.destroy_process_group()
