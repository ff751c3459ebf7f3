"""Forward table lookup with the coarse levels staged in shared memory (RADNERF_GRID_STAGE_KB, gridencoder_impl.cuh STAGE) against the
plain kernel: device time of `rn_grid_encode_forward` through the GridEncoder module for a sweep of per-CTA shared-memory budgets.

    python tools/grid_stage_probe.py         # writes gpurun_out/grid_forward_staged.json

Inputs: the head model's two tables (3-D 16 levels x 2, 2^16 rows per level cap, fp16 under autocast; 2-D the same) on (a) uniform
random positions and (b) positions along marched rays of the synthetic head (what a frame really looks up: samples of a ray are
consecutive rows, neighbouring rays neighbouring threads)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch
from gridencoder import GridEncoder

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B = 1 << 20
res = {"B": B, "rows": []}


def timed(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def ray_positions(D):
    """points along rays through a box, 16 consecutive samples per ray, neighbouring rays from neighbouring pixels"""
    n_rays = B // 16
    side = int(np.sqrt(n_rays))
    u, v = np.meshgrid(np.linspace(-0.35, 0.35, side), np.linspace(-0.35, 0.35, n_rays // side), indexing="xy")
    t = np.linspace(-0.4, 0.4, 16)
    p = np.stack([u[..., None] + 0.05 * t, v[..., None] - 0.03 * t, np.broadcast_to(t, u.shape + (16,))], -1).reshape(-1, 3)
    p = p[:, :D].astype(np.float32)
    pad = np.zeros((B - p.shape[0], D), np.float32)
    return torch.from_numpy(np.concatenate([p, pad])).to(dev)


for D in (3, 2):
    enc = GridEncoder(input_dim=D, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048).to(dev)
    with torch.no_grad():
        enc.embeddings.uniform_(-1e-1, 1e-1)
    sizes = (enc.offsets[1:] - enc.offsets[:-1]).tolist()
    for name, x in (("uniform", torch.rand(B, D, device=dev) * 2 - 1), ("rays", ray_positions(D))):
        for dtype in (torch.float16, torch.float32):
            row = {"D": D, "inputs": name, "dtype": str(dtype).split(".")[-1], "level_rows": sizes[:10], "us": {}}
            for kb in (0, 24, 48, 80, 110, 160, 220):
                if kb:
                    os.environ["RADNERF_GRID_STAGE_KB"] = str(kb)
                else:
                    os.environ.pop("RADNERF_GRID_STAGE_KB", None)

                def fwd():
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=dtype == torch.float16):
                        return enc(x, bound=1)
                row["us"][str(kb)] = round(timed(fwd), 2)
            res["rows"].append(row)
            print(json.dumps(row), flush=True)
os.environ.pop("RADNERF_GRID_STAGE_KB", None)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "grid_forward_staged.json"), "w"), indent=1)
