"""Operator bundle over the REFERENCE's own compiled CUDA extensions (oracle/_ref/*.so).  TEST / BENCH INFRASTRUCTURE.

bench.py uses it to time "the reference's kernels, called the way the reference's Python calls them" on the same GPU
next to our numbers, and GPU tests use it as a live oracle when the .so files are present.  The host-side behaviour of
the reference wrappers that costs time is reproduced: fresh zero-filled sample buffers per march call
(raymarching.py:385-393), the per-call fp32->fp16 cast of the table under autocast (grid.py:43-44), the [L,B,C] kernel
output followed by a permute+reshape copy (grid.py:47,57).  RefOps() is inference only; RefOps(train=True) adds the autograd
halves in the order of the reference's own Function classes (grid.py:65-89, raymarching.py:195-345, activation.py) so that
bench.py can time a whole training step on the reference's kernels.
"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from . import ops_frame  # noqa: F401  (registers the op-by-op inference loop with radnerf_b200.model)

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_mods = {}


def available():
    return all(os.path.exists(os.path.join(_REF_DIR, n + ".so")) for n in
               ("_gridencoder", "_raymarching_face", "_freqencoder", "_shencoder"))


def backend(name):
    if name not in _mods:
        if _REF_DIR not in sys.path:
            sys.path.insert(0, _REF_DIR)
        _mods[name] = importlib.import_module(name)
    return _mods[name]


class _RM:
    @staticmethod
    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        rays_o = rays_o.float().contiguous().view(-1, 3)
        rays_d = rays_d.float().contiguous().view(-1, 3)
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        backend("_raymarching_face").near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
        return nears, fars

    @staticmethod
    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                   perturb=False, dt_gamma=0, max_steps=1024):
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        M = n_alive * n_step
        if align > 0:
            M += align - (M % align)
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=rays_o.device)
        if perturb:
            noises = torch.rand(n_alive, dtype=rays_o.dtype, device=rays_o.device)
        else:
            noises = torch.zeros(n_alive, dtype=rays_o.dtype, device=rays_o.device)
        backend("_raymarching_face").march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps,
                                                C, H, density_bitfield, near, far, xyzs, dirs, deltas, noises)
        return xyzs, dirs, deltas

    @staticmethod
    def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        backend("_raymarching_face").composite_rays(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas.float(), rgbs.float(),
                                                    deltas, weights_sum, depth, image)
        return tuple()

    @staticmethod
    def morton3D(coords):
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        backend("_raymarching_face").morton3D(coords.int(), N, indices)
        return indices

    @staticmethod
    def packbits(grid, thresh, bitfield=None):
        grid = grid.float().contiguous()
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        backend("_raymarching_face").packbits(grid, N, thresh, bitfield)
        return bitfield

    @staticmethod
    def morton3D_dilation(grid):
        grid = grid.float().contiguous()
        out = torch.empty_like(grid)
        backend("_raymarching_face").morton3D_dilation(grid, grid.shape[0], int(round(grid.shape[1] ** (1 / 3))), out)
        return out


class GridEncoderRef(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False, interpolation='linear'):
        super().__init__()
        from . import oracle as O
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

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim).contiguous()
        B, D = inputs.shape
        L, C = self.num_levels, self.level_dim
        emb = self.embeddings
        if torch.is_autocast_enabled() and C % 2 == 0:
            emb = emb.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=emb.dtype)
        backend("_gridencoder").grid_encode_forward(inputs, emb, self.offsets, outputs, B, D, C, L,
                                                    float(np.log2(self.per_level_scale)), self.base_resolution, None,
                                                    self.gridtype_id, self.align_corners, self.interp_id)
        outputs = outputs.permute(1, 0, 2).reshape(B, L * C)
        return outputs.view(prefix + [self.output_dim])


class FreqEncoderRef(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim + input_dim * 2 * degree

    def forward(self, inputs, **kw):
        prefix = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, self.input_dim).float().contiguous()
        B = inputs.shape[0]
        out = torch.empty(B, self.output_dim, dtype=inputs.dtype, device=inputs.device)
        backend("_freqencoder").freq_encode_forward(inputs, B, self.input_dim, self.degree, self.output_dim, out)
        return out.reshape(prefix + [self.output_dim])


class SHEncoderRef(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix = list(inputs.shape[:-1])
        inputs = inputs.reshape(-1, 3).float().contiguous()
        B = inputs.shape[0]
        out = torch.empty(B, self.output_dim, dtype=inputs.dtype, device=inputs.device)
        backend("_shencoder").sh_encode_forward(inputs, out, B, 3, self.degree, None)
        return out.reshape(prefix + [self.output_dim])


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16,
                log2_hashmap_size=19, desired_resolution=2048, align_corners=False, **kwargs):
    if encoding == 'frequency':
        enc = FreqEncoderRef(input_dim=input_dim, degree=multires)
    elif encoding == 'spherical_harmonics':
        enc = SHEncoderRef(input_dim=input_dim, degree=degree)
    elif encoding in ('hashgrid', 'tiledgrid'):
        enc = GridEncoderRef(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                             log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution,
                             gridtype='hash' if encoding == 'hashgrid' else 'tiled', align_corners=align_corners,
                             interpolation=kwargs.get('interpolation', 'linear'))
    else:
        raise NotImplementedError(encoding)
    return enc, enc.output_dim


# ------------------------------------------------------------------------------------------------ training (autograd) halves
from torch.amp import custom_bwd as _custom_bwd, custom_fwd as _custom_fwd   # noqa: E402

_fwd_f32 = _custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_fwd = _custom_fwd(device_type="cuda")
_bwd = _custom_bwd(device_type="cuda")


class _GridFn(torch.autograd.Function):
    """grid.py:27-89: half table under autocast, [L,B,C] kernel output + permute copy; backward permutes the gradient back,
    accumulates into a zeros_like(embeddings) (fp16 atomics under autocast) and returns it in that dtype"""

    @staticmethod
    @_fwd
    def forward(ctx, inputs, emb, offsets, S, H, calc_grad_inputs, gridtype, align_corners, interp):
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, emb.shape[1]
        if torch.is_autocast_enabled() and C % 2 == 0:
            emb = emb.to(torch.half)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=emb.dtype)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=emb.dtype) if calc_grad_inputs else None
        backend("_gridencoder").grid_encode_forward(inputs, emb, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp)
        ctx.save_for_backward(inputs, emb, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, align_corners, interp)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    @_bwd
    def backward(ctx, grad):
        inputs, emb, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, align_corners, interp = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_emb = torch.zeros_like(emb)
        grad_inputs = torch.zeros_like(inputs, dtype=emb.dtype) if dy_dx is not None else None
        backend("_gridencoder").grid_encode_backward(grad, inputs, emb, offsets, grad_emb, B, D, C, L, S, H, dy_dx, grad_inputs,
                                                     gridtype, align_corners, interp)
        if grad_inputs is not None:
            grad_inputs = grad_inputs.to(inputs.dtype)
        return grad_inputs, grad_emb, None, None, None, None, None, None, None


class GridEncoderRefTrain(GridEncoderRef):
    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        out = _GridFn.apply(inputs, self.embeddings, self.offsets, float(np.log2(self.per_level_scale)), self.base_resolution,
                            inputs.requires_grad, self.gridtype_id, self.align_corners, self.interp_id)
        return out.view(prefix + [self.output_dim])


class _MarchTrainFn(torch.autograd.Function):
    """raymarching.py:195-262 (forward only: camera poses are not optimised)"""

    @staticmethod
    @_fwd_f32
    def forward(ctx, rays_o, rays_d, bound, bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False, align=-1,
                force_all_rays=False, dt_gamma=0, max_steps=1024):
        rays_o, rays_d = rays_o.contiguous().view(-1, 3), rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=rays_o.device)
        rays = torch.empty(N, 3, dtype=torch.int32, device=rays_o.device)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=rays_o.device)
        noises = torch.rand(N, dtype=rays_o.dtype, device=rays_o.device) if perturb else torch.zeros(N, dtype=rays_o.dtype, device=rays_o.device)
        backend("_raymarching_face").march_rays_train(rays_o, rays_d, bitfield.contiguous(), bound, dt_gamma, max_steps, N, C, H, M,
                                                      nears, fars, xyzs, dirs, deltas, rays, step_counter, noises)
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()
            if align > 0:
                m += align - m % align
            xyzs, dirs, deltas = xyzs[:m], dirs[:m], deltas[:m]
            torch.cuda.empty_cache()
        return xyzs, dirs, deltas, rays


class _CompositeTrainFn(torch.autograd.Function):
    """raymarching.py:283-342"""

    @staticmethod
    @_fwd_f32
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, ambient = sigmas.contiguous(), rgbs.contiguous(), ambient.contiguous()
        M, N = sigmas.shape[0], rays.shape[0]
        weights_sum, ambient_sum, depth = (torch.empty(N, dtype=sigmas.dtype, device=sigmas.device) for _ in range(3))
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)
        backend("_raymarching_face").composite_rays_train_forward(sigmas, rgbs, ambient, deltas, rays, M, N, T_thresh, weights_sum,
                                                                  ambient_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image)
        ctx.dims = (M, N, T_thresh)
        return weights_sum, ambient_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, g_ws, g_as, g_depth, g_img):
        sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims
        g_s, g_r, g_a = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(ambient)
        backend("_raymarching_face").composite_rays_train_backward(g_ws.contiguous(), g_as.contiguous(), g_img.contiguous(), sigmas, rgbs,
                                                                   ambient, deltas, rays, weights_sum, ambient_sum, image, M, N, T_thresh,
                                                                   g_s, g_r, g_a)
        return g_s, g_r, g_a, None, None, None


class _TruncExpFn(torch.autograd.Function):
    """activation.py:5-17"""

    @staticmethod
    @_fwd_f32
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    @_bwd
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class _RMTrain(_RM):
    march_rays_train = staticmethod(_MarchTrainFn.apply)
    composite_rays_train = staticmethod(_CompositeTrainFn.apply)


def get_encoder_train(encoding, **kw):
    enc, dim = get_encoder(encoding, **kw)
    if isinstance(enc, GridEncoderRef):
        enc.__class__ = GridEncoderRefTrain
    return enc, dim


class _TruncExp:
    """activation.py:3-17, forward only: exp in fp32"""

    def __call__(self, x):
        return torch.exp(x.float())


class RefOps:
    def __init__(self, train=False):
        self.rm = _RMTrain if train else _RM
        self.get_encoder = get_encoder_train if train else get_encoder
        self.trunc_exp = _TruncExpFn.apply if train else _TruncExp()
