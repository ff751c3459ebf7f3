"""Golden vectors for the TORSO branch of the occupancy maintenance (nerf/renderer.py:455-501): the
reference's own NeRFNetwork(torso=True).update_extra_state() on the CPU -- random audio window and pose, torso alpha queried at
every cell of the 128^2 grid through forward_torso (frequency encoders, deformation MLP, 2-D torso grid, torso MLP), stored
x/y-transposed, 5x5 max-pool, max(decayed old, fresh).  Weights from tests/network_case.fill_parameters, encoders bound to the
oracle's CPU operators, in-cell jitter fixed at the cell centre (torch.rand_like -> 0.5), Python's `random` seeded.

    python tests/golden/make_torso_occupancy_golden.py   ->  tests/golden/occupancy_torso.npz   (needs /root/reference)
"""
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from network_case import fill_parameters   # noqa: E402
from occupancy_case import torso_case   # noqa: E402


def main():
    from oracle import cpu_backend
    ops = cpu_backend.CPUOps()
    rm = types.ModuleType("raymarching")
    for n in ("morton3D", "morton3D_dilation", "packbits", "near_far_from_aabb", "march_rays", "composite_rays"):
        setattr(rm, n, getattr(ops.rm, n))
    sys.modules["raymarching"] = rm
    enc = types.ModuleType("encoding")
    enc.get_encoder = cpu_backend.get_encoder
    sys.modules["encoding"] = enc
    act = types.ModuleType("activation")
    act.trunc_exp = torch.exp
    sys.modules["activation"] = act
    for name in ("trimesh", "tensorboardX", "matplotlib", "matplotlib.pyplot", "mcubes", "imageio", "lpips"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")
        m.ExponentialMovingAverage = object
        sys.modules["torch_ema"] = m
    sys.path.append("/root/reference")
    from nerf.network import NeRFNetwork
    from radnerf_b200.model import Options

    o = Options(torso=True, smooth_lips=False, fp16=False, exp_eye=True)
    net = NeRFNetwork(types.SimpleNamespace(**{**vars(o), "test_train": False}))
    fill_parameters(net)
    scales = np.load(os.path.join(HERE, "grid_g3_f32.npz"))["scales"]
    for e in (net.encoder, net.encoder_ambient, net.torso_encoder):
        e.device_scales = scales
    c = torso_case()
    net.aud_features, net.eye_area, net.poses = c["aud_features"], c["eye_area"], c["poses"]
    net.density_grid_torso.copy_(c["grid0"])
    net.local_step = 2
    net.step_counter[:2, 0] = torch.tensor([700, 901], dtype=torch.int32)
    orig = torch.rand_like
    torch.rand_like = lambda t, **kw: torch.full_like(t, 0.5)
    random.seed(c["seed"])
    try:
        net.update_extra_state()
        g1, md1, mc = net.density_grid_torso.clone(), net.mean_density_torso, net.mean_count
        net.update_extra_state()
    finally:
        torch.rand_like = orig
    np.savez_compressed(os.path.join(HERE, "occupancy_torso.npz"), grid_after_1=g1.numpy(), grid_after_2=net.density_grid_torso.numpy(),
                        mean_density_1=np.float64(md1), mean_density_2=np.float64(net.mean_density_torso), mean_count=np.int64(mc))
    print("mean density", md1, net.mean_density_torso, "mean_count", mc, "grid range", float(g1.min()), float(g1.max()))


if __name__ == "__main__":
    main()
