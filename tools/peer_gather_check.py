"""torchrun --nproc-per-node N tools/peer_gather_check.py : the peer-store gather (csrc/peer_gather.cu + symmetric memory)
must assemble exactly the frame NCCL all_gather + un-permute assembles; prints both timings (device events, max over ranks)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import torch
import torch.distributed as dist
from radnerf_b200.sharding import FrameSharder
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H = W = int(os.environ.get("HW", 512))
sh = FrameSharder(H, W, world, rank, dev)
full = torch.rand(H * W, 3, generator=torch.Generator().manual_seed(0)).to(dev)
ok = True
ref = [sh.gather(sh.shard(full * (i + 1))).clone() for i in range(4)]
enabled = sh.enable_peer_gather()
res = {"world": world, "peer_gather_enabled": enabled, "error": getattr(sh, "peer_error", None)}
if enabled:
    for i in range(4):
        got = sh.gather(sh.shard(full * (i + 1)))
        ok = ok and torch.equal(got, ref[i]) and torch.equal(got, full * (i + 1))
    def timeit(fn, n=200):
        for _ in range(20): fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    loc = sh.shard(full)
    t_peer = timeit(lambda: sh.gather(loc))
    peer = sh.peer; sh.peer = None
    t_nccl = timeit(lambda: sh.gather(loc))
    sh.peer = peer
    res.update(identical=bool(ok), us_peer=1e3 * t_peer, us_nccl=1e3 * t_nccl)
flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
res["identical_on_all_ranks"] = bool(flag.item())
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
