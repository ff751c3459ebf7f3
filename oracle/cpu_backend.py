"""CPU operator bundle over the oracle (liboracle.so) for radnerf_b200.model.NeRFNetwork.  TEST INFRASTRUCTURE ONLY.

Gives the model mirror a complete CPU execution of the inference frame: the reference's kernels restated in C
(oracle.c, OpenMP over the host cores) for encode / march / composite, torch-CPU nn.Linear/Conv1d for the small MLPs
and the audio nets.  Used by bench.py for the `cpu_baseline` object and the `--impl reference` arm (the reference has no
CPU implementation of its own -- its extensions are CUDA-only -- so this port IS the CPU baseline), and by tests.
CPUOps() is the inference bundle; CPUOps(train=True) adds autograd through the oracle ops (grid encoder forward/backward with
dy_dx, march_rays_train, composite_rays_train forward/backward, trunc_exp's clamped backward) so that a whole training step --
the reference's NeRFNetwork / NeRFRenderer.run_cuda / Trainer.train_step, or our mirror -- runs and differentiates on the CPU
(tests/golden/make_train_golden.py, tests/test_train_parity.py).
"""
import numpy as np
import torch
import torch.nn as nn

from . import oracle as O
from . import ops_frame  # noqa: F401  (registers the op-by-op inference loop with radnerf_b200.model)


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
        flat = inputs.reshape(-1, self.input_dim).float()
        if torch.is_grad_enabled() and (flat.requires_grad or self.embeddings.requires_grad):
            out = _GridFnCPU.apply(flat, self.embeddings, self)
        else:
            out, _ = O.grid_encode_forward(_np(flat), _np(self.embeddings), _np(self.offsets), self.per_level_scale,
                                           self.base_resolution, False, self.gridtype_id, self.align_corners, self.interp_id,
                                           scales=self.device_scales)
            out = torch.from_numpy(out)
        return out.view(prefix + [self.output_dim])


class _GridFnCPU(torch.autograd.Function):
    """_grid_encode (gridencoder/grid.py:27-89) on the oracle kernels, fp32: dy_dx only when the input needs a gradient"""

    @staticmethod
    def forward(ctx, x, emb, enc):
        want = bool(x.requires_grad)
        out, dy = O.grid_encode_forward(_np(x), _np(emb), _np(enc.offsets), enc.per_level_scale, enc.base_resolution, want,
                                        enc.gridtype_id, enc.align_corners, enc.interp_id, scales=enc.device_scales)
        ctx.save_for_backward(x, emb)
        ctx.enc, ctx.dy = enc, dy
        return torch.from_numpy(out)

    @staticmethod
    def backward(ctx, grad):
        x, emb = ctx.saved_tensors
        enc = ctx.enc
        ge, gi = O.grid_encode_backward(_np(grad.contiguous().float()), _np(x), _np(enc.offsets), enc.per_level_scale,
                                        enc.base_resolution, emb.shape[0], emb.shape[1], dy_dx=ctx.dy, gridtype=enc.gridtype_id,
                                        align_corners=enc.align_corners, interpolation=enc.interp_id, scales=enc.device_scales)
        return (None if gi is None else torch.from_numpy(gi.astype(np.float32))), torch.from_numpy(ge.astype(np.float32)), None


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


# ------------------------------------------------------------------------------------------------ training (autograd) halves
TRAIN_NOISE = None    # per-ray start offsets for perturb=True, set by the caller ([N] float32 in [0,1)); None draws np.random


class _CompositeTrainCPU(torch.autograd.Function):
    """_composite_rays_train (raymarching/raymarching.py:283-342) on the oracle kernels"""

    @staticmethod
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, ambient = sigmas.float().contiguous(), rgbs.float().contiguous(), ambient.float().contiguous()
        ws, am, dp, im = O.composite_rays_train_forward(_np(sigmas), _np(rgbs), _np(ambient), _np(deltas), _np(rays), T_thresh)
        ws, am, dp, im = (torch.from_numpy(a) for a in (ws, am, dp, im))
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, ws, im)
        ctx.T_thresh = T_thresh
        return ws, am, dp, im

    @staticmethod
    def backward(ctx, g_ws, g_am, g_dp, g_im):
        sigmas, rgbs, deltas, rays, ws, im = ctx.saved_tensors
        gs, gr, ga = O.composite_rays_train_backward(_np(g_ws.contiguous()), _np(g_am.contiguous()), _np(g_im.contiguous()), _np(sigmas),
                                                     _np(rgbs), _np(deltas), _np(rays), _np(ws), _np(im), ctx.T_thresh)
        return torch.from_numpy(gs), torch.from_numpy(gr), torch.from_numpy(ga), None, None, None


class _TruncExpCPU(torch.autograd.Function):
    """activation.py:5-17"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        return g * torch.exp(ctx.saved_tensors[0].clamp(-15, 15))


class _RMTrain(_RM):
    @staticmethod
    def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                         align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """_march_rays_train.forward (raymarching/raymarching.py:195-262); rays are visited in order"""
        o, d = _np(rays_o).reshape(-1, 3), _np(rays_d).reshape(-1, 3)
        N = o.shape[0]
        M = N * max_steps
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count
        if perturb:
            noises = np.random.rand(N).astype(np.float32) if TRAIN_NOISE is None else np.asarray(TRAIN_NOISE, np.float32)[:N]
        else:
            noises = np.zeros(N, np.float32)
        x, dd, dl, rays, counter = O.march_rays_train(o, d, bound, _np(density_bitfield), C, H, _np(nears), _np(fars), noises, M, dt_gamma,
                                                      max_steps)
        if step_counter is not None:
            step_counter.copy_(torch.from_numpy(counter))
        if force_all_rays or mean_count <= 0:
            m = int(counter[0])
            if align > 0:
                m += align - m % align
            x, dd, dl = x[:m], dd[:m], dl[:m]
        return torch.from_numpy(x), torch.from_numpy(dd), torch.from_numpy(dl), torch.from_numpy(rays)

    composite_rays_train = staticmethod(_CompositeTrainCPU.apply)


class CPUOps:
    def __init__(self, train=False):
        self.rm = _RMTrain if train else _RM
        self.get_encoder = get_encoder
        self.trunc_exp = _TruncExpCPU.apply if train else torch.exp
