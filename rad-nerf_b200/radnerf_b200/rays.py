"""Device-side ray generation from a camera pose (the reference builds rays per frame with ~15 torch ops in get_rays,
nerf/utils.py:248-333; SURVEY 8(f) ranks this first among the callers to absorb).  One kernel: pixel id -> (origin, dir)."""
import torch

from . import abi



class RayGenerator:
    def __init__(self, H, W, intrinsics, device, sharder=None):
        self.H, self.W = H, W
        self.fx, self.fy, self.cx, self.cy = [float(v) for v in intrinsics]
        self.ids = None if sharder is None or sharder.ids is None else sharder.ids.to(torch.int32).contiguous()
        self.n = H * W if self.ids is None else self.ids.numel()
        self.device = device

    def __call__(self, pose, out=None):
        """pose [4,4] fp32 on the device -> rays_o [n,3], rays_d [n,3] (written into `out = (rays_o, rays_d)` if given)"""
        pose = pose.contiguous()
        if out is None:
            ro = torch.empty(self.n, 3, device=self.device)
            rd = torch.empty(self.n, 3, device=self.device)
        else:
            ro, rd = out
            assert ro.is_contiguous() and rd.is_contiguous() and ro.numel() == 3 * self.n and rd.numel() == 3 * self.n
        abi.check(abi.lib().rn_get_rays(abi.ptr(pose), self.fx, self.fy, self.cx, self.cy, self.H, self.W, abi.ptr(self.ids),
                                        self.n, abi.ptr(ro), abi.ptr(rd), abi.cur_stream()))
        return ro, rd
