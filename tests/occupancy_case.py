"""Inputs of the occupancy-maintenance parity case (row a19), shared by tests/golden/make_occupancy_golden.py (which runs the
reference's code on them) and tests/test_gpu_occupancy.py (which runs ours)."""
import numpy as np
import torch


def analytic_sigma(x):
    """a smooth head-like blob plus a thin off-centre shell: deterministic stand-in for the density network, [N,3] -> [N]"""
    x = x.float()
    q = (x[:, 0] / 0.34) ** 2 + (x[:, 1] / 0.24) ** 2 + (x[:, 2] / 0.37) ** 2
    shell = torch.exp(-((x - x.new_tensor([0.45, 0.1, -0.3])).norm(dim=-1) - 0.2) ** 2 * 400.0)
    return 40.0 * torch.exp(-3.0 * q) + 15.0 * shell


def case_inputs():
    """three training cameras on the orbit of the synthetic sequence, intrinsics of a 450x450 frame"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rad-nerf_b200"))
    from radnerf_b200 import synthetic as syn
    poses = np.stack([syn.orbit_pose(yaw_deg=y, pitch_deg=2.0) for y in (-10.0, 0.0, 10.0)]).astype(np.float32)
    return dict(poses=poses, intrinsics=np.asarray(syn.intrinsics_for(450, 450), np.float32))


def torso_case():
    """state the torso-phase update needs (main.py:210-212 hands the model the dataset's audio features, eye areas and poses):
    a bank of audio windows, 16 orbit poses, a previous 2-D alpha grid, and the seed of Python's `random`"""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rad-nerf_b200"))
    from radnerf_b200 import synthetic as syn
    poses = np.stack([syn.orbit_pose(yaw_deg=float(y), pitch_deg=2.0) for y in np.linspace(-10, 10, 16)]).astype(np.float32)
    return dict(aud_features=torch.from_numpy(syn.audio_feature_bank(64, 44, 16, seed=3)), eye_area=torch.full((64, 1), 0.25),
                poses=torch.from_numpy(poses), grid0=torch.from_numpy(syn.torso_density_grid(128)) * 0.5, seed=5)
