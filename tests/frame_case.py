"""Inputs of the whole-frame parity case (row a18), shared by tests/golden/make_frame_golden.py (the reference's renderer)
and tests/test_frame_parity.py (ours)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rad-nerf_b200"))
HW = 48


def install_occupancy(net):
    """the bench's synthetic head / torso occupancy, written into a reference or mirror model alike"""
    from radnerf_b200 import synthetic as syn
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    with torch.no_grad():
        net.density_grid.copy_(torch.from_numpy(grid).to(net.density_grid.device))
        net.mean_density = float(np.clip(grid, 0, None).mean())
        net.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(net.mean_density, net.density_thresh))).to(net.density_bitfield.device))
        tg = syn.torso_density_grid(128)
        net.density_grid_torso.copy_(torch.from_numpy(tg).to(net.density_grid_torso.device))
        net.mean_density_torso = float(tg.mean())


def frame_inputs():
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.posemath import convert_poses
    pose = syn.orbit_pose(yaw_deg=6.0, pitch_deg=2.0)
    ro, rd = syn.get_rays(pose, syn.intrinsics_for(HW, HW), HW, HW)
    g = torch.Generator().manual_seed(11)
    return dict(rays_o=torch.from_numpy(ro)[None], rays_d=torch.from_numpy(rd)[None], auds=torch.randn(8, 44, 16, generator=g) * 3.0,
                bg_coords=torch.from_numpy(syn.get_bg_coords(HW, HW))[None], poses=convert_poses(torch.from_numpy(pose)[None]),
                eye=torch.tensor([[0.25]]))
