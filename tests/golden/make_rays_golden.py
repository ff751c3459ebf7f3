"""Golden vectors for SURVEY 8(f) rank 1 (ray / pose / background-coordinate generation), produced by the REFERENCE's own
functions on the CPU of this container: nerf/utils.py get_rays (:248-333), get_bg_coords (:239-245), convert_poses (:230-237),
imported unmodified through baseline/stock.py.

    python tests/golden/make_rays_golden.py      ->  tests/golden/rays.npz   (needs /root/reference or baseline/_ref)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from rays_case import CASES, case_pose, case_intrinsics   # noqa: E402


def main():
    from baseline import install_ref, stock
    install_ref.ensure()
    U = stock.load("ours").utils     # the stock nerf/utils.py (its functions here depend on torch only)
    out = {}
    for name, (H, W, yaw, pitch) in CASES.items():
        pose = torch.from_numpy(case_pose(yaw, pitch))[None]
        intr = case_intrinsics(H, W)
        r = U.get_rays(pose, intr, H, W, -1)
        out[name + "_rays_o"] = r["rays_o"][0].contiguous().numpy()
        out[name + "_rays_d"] = r["rays_d"][0].contiguous().numpy()
        out[name + "_bg_coords"] = U.get_bg_coords(H, W, "cpu")[0].numpy()
        out[name + "_pose6"] = U.convert_poses(pose).numpy()
        out[name + "_pose"] = pose[0].numpy()
    np.savez_compressed(os.path.join(HERE, "rays.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
