"""Golden vectors for SURVEY 8(a) rows a16 / a17, produced by the REFERENCE's own NeRFNetwork (nerf/network.py:91-325).

The reference class is imported from /root/reference and constructed on the CPU: `encoding` / `raymarching` are bound to
the oracle's CPU operators (grid / frequency / SH encoders pinned bit-exact resp. <= 1e-5 to the reference's CUDA kernels by
tests/test_oracle_golden.py), GUI / IO packages that nerf/utils.py imports are stubbed.  Parameters are overwritten by
tests/network_case.fill_parameters (a function of the parameter NAME only, so our mirror gets the same weights), then
encode_audio / forward / density / forward_torso run in fp32.

    python tests/golden/make_network_golden.py      ->  tests/golden/network.npz   (needs /root/reference)
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
from network_case import fill_parameters, inputs   # noqa: E402


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
    # level scales as CUDA's exp2f produces them (recorded from the device in the grid goldens; same L / H / resolution)
    scales = np.load(os.path.join(HERE, "grid_g3_f32.npz"))["scales"]
    for e in (net.encoder, net.encoder_ambient, net.torso_encoder):
        assert e.num_levels == len(scales) and e.base_resolution == 16
        e.device_scales = scales
    c = inputs()
    out = {}
    with torch.no_grad():
        enc_a = net.encode_audio(c["auds"])
        out["enc_a"] = enc_a.numpy()
        ind = net.individual_codes[0]
        sigma, color, ambient = net(c["x"], c["d"], enc_a, ind, c["eye"])
        out.update(sigma=sigma.numpy(), color=color.numpy(), ambient=ambient.numpy())
        den = net.density(c["x"], enc_a, c["eye"])
        out.update(density_sigma=den["sigma"].numpy(), density_geo=den["geo_feat"].numpy())
        alpha, rgb, deform = net.forward_torso(c["xy"], c["poses"], enc_a, net.individual_codes_torso[0])
        out.update(torso_alpha=alpha.numpy(), torso_color=rgb.numpy(), torso_deform=deform.numpy())
    np.savez_compressed(os.path.join(HERE, "network.npz"), **out)
    print({k: (v.shape, float(np.abs(v).max())) for k, v in out.items()})


if __name__ == "__main__":
    main()
