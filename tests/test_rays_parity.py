"""ray / pose / background-coordinate generation against the reference's own functions
(nerf/utils.py get_rays :248-333, get_bg_coords :239-245, convert_poses :230-237; golden vectors produced by
tests/golden/make_rays_golden.py from the unmodified reference on the CPU)."""
import os

import numpy as np
import pytest
import torch

from rays_case import CASES, case_intrinsics

GOLD = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rays.npz")))


@pytest.mark.parametrize("name", list(CASES))
def test_host_input_generators_match_the_reference(name):
    """the numpy generators bench/tests build inputs with (radnerf_b200.synthetic, posemath) are the reference's formulas"""
    from radnerf_b200 import synthetic as syn
    from radnerf_b200.posemath import convert_poses
    H, W, _, _ = CASES[name]
    pose = GOLD[name + "_pose"]
    ro, rd = syn.get_rays(pose, case_intrinsics(H, W), H, W)
    assert np.abs(ro - GOLD[name + "_rays_o"]).max() == 0
    assert np.abs(rd - GOLD[name + "_rays_d"]).max() <= 1e-6
    assert np.abs(syn.get_bg_coords(H, W) - GOLD[name + "_bg_coords"]).max() <= 1e-6
    assert np.abs(convert_poses(torch.from_numpy(pose)[None]).numpy() - GOLD[name + "_pose6"]).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("world", [1, 2])
def test_rn_get_rays_matches_the_reference_get_rays(name, world):
    """device ray generation (csrc/rays.cu), full frame and the row-tile shards of a 2-GPU frame, <= 1e-6"""
    from radnerf_b200.rays import RayGenerator
    from radnerf_b200.sharding import FrameSharder
    H, W, _, _ = CASES[name]
    dev = torch.device("cuda")
    pose = torch.from_numpy(GOLD[name + "_pose"]).to(dev)
    ro_ref, rd_ref = torch.from_numpy(GOLD[name + "_rays_o"]).to(dev), torch.from_numpy(GOLD[name + "_rays_d"]).to(dev)
    if world == 2 and (H % 16):
        pytest.skip("frame does not split into 2 x 8-row tiles")
    for rank in range(world):
        sh = FrameSharder(H, W, world, rank, dev) if world > 1 else None
        ro, rd = RayGenerator(H, W, case_intrinsics(H, W), dev, sh)(pose)
        sel = slice(None) if sh is None else sh.ids
        assert (ro - ro_ref[sel]).abs().max().item() == 0.0
        assert (rd - rd_ref[sel]).abs().max().item() <= 1e-6
        assert ro.shape[0] == H * W // world


@pytest.mark.gpu
def test_device_convert_poses_matches_the_reference():
    """the conditioning kernel derives the torso's 6-vector pose from the 4x4 cam2world itself (rn_conditioning_desc.pose44)"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from radnerf_b200 import frame
    dev = torch.device("cuda")
    model = bench.make_model(dev, seed=2)
    f = bench.make_frames(32, 1)[0][0]
    auds, eye = torch.from_numpy(f["auds"]).to(dev), torch.from_numpy(f["eye"]).to(dev)
    for name in CASES:
        pose = torch.from_numpy(GOLD[name + "_pose"]).to(dev)
        want = torch.from_numpy(GOLD[name + "_pose6"]).to(dev)
        out6 = torch.zeros(6, device=dev)
        model.enc_a = None
        frame.launch_conditioning(model, 0, auds, eye, poses=out6, pose44=pose)       # device-side conversion, result written to out6
        consts_dev = frame.lane_state(model, 0).torso_consts.clone()
        model.enc_a = None
        frame.launch_conditioning(model, 0, auds, eye, poses=want.clone())            # host-converted 6-vector
        consts_host = frame.lane_state(model, 0).torso_consts.clone()
        torch.cuda.synchronize()
        assert (out6 - want.view(-1)).abs().max().item() <= 1e-6
        assert (consts_dev - consts_host).abs().max().item() <= 2e-3 * max(1.0, consts_host.abs().max().item())
