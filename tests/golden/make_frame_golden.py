"""Golden frame for SURVEY 8(a) row a18, rendered by the REFERENCE's own NeRFNetwork.render -> NeRFRenderer.run_cuda
(nerf/renderer.py:158-316, 504-537) on the CPU of this container: `raymarching` / `encoding` / `activation` are bound to the
oracle's CPU operators (pinned to the reference's CUDA kernels by tests/test_oracle_golden.py), weights come from
tests/network_case.fill_parameters, occupancy from the synthetic head / torso grids of the bench, fp32, no smoothing.

    python tests/golden/make_frame_golden.py      ->  tests/golden/frame.npz   (needs /root/reference)
"""
import os
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
from frame_case import frame_inputs, install_occupancy, HW   # noqa: E402


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

    opt = Options(torso=True, smooth_lips=False, fp16=False)
    net = NeRFNetwork(types.SimpleNamespace(**{**vars(opt), "test_train": False})).eval()
    fill_parameters(net)
    scales = np.load(os.path.join(HERE, "grid_g3_f32.npz"))["scales"]
    for e in (net.encoder, net.encoder_ambient, net.torso_encoder):
        e.device_scales = scales
    install_occupancy(net)
    f = frame_inputs()
    with torch.no_grad():
        out = net.render(f["rays_o"], f["rays_d"], f["auds"], f["bg_coords"], f["poses"], eye=f["eye"], index=[0], staged=False,
                         bg_color=None, perturb=False, force_all_rays=True, dt_gamma=opt.dt_gamma, max_steps=opt.max_steps)
    res = dict(image=out["image"].numpy().reshape(HW * HW, 3), depth=out["depth"].numpy().reshape(-1),
               torso_alpha=out["torso_alpha"].numpy().reshape(-1), torso_color=out["torso_color"].numpy().reshape(-1, 3))
    np.savez_compressed(os.path.join(HERE, "frame.npz"), **res)
    print({k: (v.shape, float(v.min()), float(v.max())) for k, v in res.items()})
    print("non-background pixels:", int((np.abs(res["image"] - 1).max(-1) > 1e-3).sum()), "of", HW * HW)


if __name__ == "__main__":
    main()
