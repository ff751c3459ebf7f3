"""CPU operator bundle over the oracle (liboracle.so) for radnerf_b200.model.NeRFNetwork.  TEST INFRASTRUCTURE ONLY.

Gives the model mirror a complete CPU execution of the inference frame: the reference's kernels restated in C
(oracle.c, OpenMP over the host cores) for encode / march / composite, torch-CPU nn.Linear/Conv1d for the small MLPs
and the audio nets.  Used by bench.py for the `cpu_baseline` object and the `--impl reference` arm (the reference has no
CPU implementation of its own -- its extensions are CUDA-only -- so this port IS the CPU baseline), and by tests.
Inference only (no autograd through the oracle ops).
"""
import numpy as np
import torch
import torch.nn as nn

from . import oracle as O


def _np(t):
    return t.detach().cpu().numpy()


class _RM:
    """the raymarching functions the inference frame uses, on CPU tensors"""

    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = O.near_far_from_aabb(_np(rays_o), _np(rays_d), _np(aabb), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                   perturb=False, dt_gamma=0, max_steps=1024):
        noises = np.random.rand(n_alive).astype(np.float32) if perturb else None
        x, d, dl = O.march_rays(n_alive, n_step, _np(rays_alive), _np(rays_t), _np(rays_o), _np(rays_d), bound,
                                _np(density_bitfield), C, H, _np(near), _np(far), align, noises, dt_gamma, max_steps)
        return torch.from_numpy(x), torch.from_numpy(d), torch.from_numpy(dl)

    @staticmethod
    def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        # numpy views share memory with the (contiguous, CPU) torch tensors -> updated in place
        O.composite_rays(n_alive, n_step, rays_alive.numpy(), rays_t.numpy(), _np(sigmas.float()), _np(rgbs.float()),
                         _np(deltas), weights_sum.numpy(), depth.numpy(), image.numpy(), T_thresh)
        return tuple()

    @staticmethod
    def morton3D(coords):
        return torch.from_numpy(O.morton3D(_np(coords)))

    @staticmethod
    def packbits(grid, thresh, bitfield=None):
        b = torch.from_numpy(O.packbits(_np(grid), thresh))
        if bitfield is not None:
            bitfield.copy_(b)
            return bitfield
        return b

    @staticmethod
    def morton3D_dilation(grid):
        return torch.from_numpy(O.morton3D_dilation(_np(grid)))


class GridEncoderCPU(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        offsets, pls = O.grid_offsets(input_dim, num_levels, level_dim, base_resolution, log2_hashmap_size, desired_resolution,
                                      per_level_scale, align_corners)
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution = pls, base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype_id = {'hash': 0, 'tiled': 1}[gridtype]
        self.interp_id = {'linear': 0, 'smoothstep': 1}[interpolation]
        self.align_corners = align_corners
        self.register_buffer('offsets', torch.from_numpy(offsets))
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim).uniform_(-1e-4, 1e-4))
        # per-level scales as the DEVICE computes them (CUDA exp2f differs from libm by an ulp on some levels);
        # None = libm.  Set from a golden file's `scales` when outputs must agree with the CUDA path to fp32 rounding.
        self.device_scales = None

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        out, _ = O.grid_encode_forward(_np(inputs.reshape(-1, self.input_dim).float()), _np(self.embeddings), _np(self.offsets),
                                       self.per_level_scale, self.base_resolution, False, self.gridtype_id,
                                       self.align_corners, self.interp_id, scales=self.device_scales)
        return torch.from_numpy(out).view(prefix + [self.output_dim])


class FreqEncoderCPU(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def forward(self, inputs, **kw):
        prefix = list(inputs.shape[:-1])
        out = O.freq_encode_forward(_np(inputs.reshape(-1, self.input_dim).float()), self.degree)
        return torch.from_numpy(out).view(prefix + [self.output_dim])


class SHEncoderCPU(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2

    def forward(self, inputs, size=1):
        prefix = list(inputs.shape[:-1])
        out, _ = O.sh_encode_forward(_np((inputs / size).reshape(-1, 3).float()), self.degree)
        return torch.from_numpy(out).view(prefix + [self.output_dim])


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16,
                log2_hashmap_size=19, desired_resolution=2048, align_corners=False, **kwargs):
    if encoding == 'frequency':
        enc = FreqEncoderCPU(input_dim=input_dim, degree=multires)
    elif encoding == 'spherical_harmonics':
        enc = SHEncoderCPU(input_dim=input_dim, degree=degree)
    elif encoding in ('hashgrid', 'tiledgrid'):
        enc = GridEncoderCPU(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                             log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                             gridtype='hash' if encoding == 'hashgrid' else 'tiled', align_corners=align_corners,
                             interpolation=kwargs.get('interpolation', 'linear'))
    else:
        raise NotImplementedError(encoding)
    return enc, enc.output_dim


class CPUOps:
    def __init__(self):
        self.rm = _RM
        self.get_encoder = get_encoder
        self.trunc_exp = torch.exp
