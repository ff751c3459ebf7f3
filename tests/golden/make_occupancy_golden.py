"""Golden vectors for SURVEY 8(a) row a19 (occupancy maintenance), produced by the REFERENCE's own code.

Runs NeRFRenderer.mark_untrained_grid and NeRFRenderer.update_extra_state (nerf/renderer.py:318-381, 383-501) from
/root/reference on the CPU of this container: the reference module is imported with the name `raymarching` bound to the
oracle's CPU operators (morton3D / morton3D_dilation / packbits -- each pinned bit-exact to the reference's CUDA kernels by
tests/test_oracle_golden.py) and with empty stubs for the GUI / IO packages nerf/utils.py imports but this path never
touches.  A deterministic analytic density stands in for the network and torch.rand_like is pinned to 0.5, which removes
the in-cell jitter, so the result is a pure function of the inputs.

    python tests/golden/make_occupancy_golden.py      ->  tests/golden/occupancy.npz   (needs /root/reference)
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
from occupancy_case import analytic_sigma, case_inputs   # noqa: E402  (shared with the GPU test)


def main():
    from oracle.cpu_backend import CPUOps
    rm = types.ModuleType("raymarching")
    ops = CPUOps().rm
    for n in ("morton3D", "morton3D_dilation", "packbits", "near_far_from_aabb", "march_rays", "composite_rays"):
        setattr(rm, n, getattr(ops, n))
    sys.modules["raymarching"] = rm
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
    from nerf.renderer import NeRFRenderer

    c = case_inputs()
    opt = types.SimpleNamespace(bound=1.0, min_near=0.05, density_thresh=10.0, density_thresh_torso=0.01, exp_eye=True,
                                test_train=False, smooth_lips=False, torso=False, cuda_ray=True, ind_num=4, ind_dim=0,
                                ind_dim_torso=0, train_camera=False)

    class Ref(NeRFRenderer):
        def encode_audio(self, a):
            return None

        def density(self, x, enc_a, e=None):
            return {"sigma": analytic_sigma(x)}

    r = Ref(opt)
    r.att = 0
    r.aud_features = torch.zeros(4, 1, 16)
    r.eye_area = torch.full((4, 1), 0.25)
    r.mark_untrained_grid(c["poses"], c["intrinsics"])
    untrained = (r.density_grid < 0).numpy().reshape(-1)
    orig = torch.rand_like
    torch.rand_like = lambda t, **kw: torch.full_like(t, 0.5)
    try:
        r.local_step = 3
        r.step_counter[:3, 0] = torch.tensor([100, 200, 301], dtype=torch.int32)
        r.update_extra_state()
        grid1 = r.density_grid.clone()
        r.update_extra_state()          # second call: the EMA branch (previous * decay vs fresh)
    finally:
        torch.rand_like = orig
    out = dict(untrained_bits=np.packbits(untrained, bitorder="little"),
               grid_after_1=grid1.numpy().astype(np.float16),      # fp16 is enough to localise a mismatch; exact checks use the bits below
               mean_density_1=np.float64(grid1.clamp(min=0).mean().item()),
               mean_density_2=np.float64(r.mean_density), bitfield_2=r.density_bitfield.numpy(),
               grid_after_2_sample=r.density_grid.numpy().reshape(-1)[::4099].copy(), mean_count=np.int64(200))
    np.savez_compressed(os.path.join(HERE, "occupancy.npz"), **out)
    print({k: (v.shape, v.dtype) if hasattr(v, "shape") else v for k, v in out.items()})
    print("untrained cells:", int(untrained.sum()), "occupied bits:", int(np.unpackbits(out["bitfield_2"]).sum()), "mean_count", r.mean_count)


if __name__ == "__main__":
    main()
