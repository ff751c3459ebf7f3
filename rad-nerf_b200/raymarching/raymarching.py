"""raymarching -- occupancy-grid utilities, ray marching and compositing on libradnerf_b200 (sm_100a).

Drop-in for /root/reference/raymarching/raymarching.py: the same ten module-level functions with the same positional
signatures, defaults, return shapes/dtypes, allocation rules (incl. the `+align` padding that adds a full block when
already aligned, raymarching.py:227-229,251-253,382-383) and AMP contract (inputs up-cast to fp32).  Each call crosses
the C ABI (include/radnerf_b200.h) with raw device pointers and torch's current stream.
"""
import numpy as np

import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi as _L

__all__ = ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits', 'morton3D_dilation',
           'march_rays_train', 'composite_rays_train', 'march_rays', 'composite_rays']

_fwd32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def _f32c(t):
    """contiguous fp32 view/copy (the reference's kernels assume both)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ----------------------------------------
# utils
# ----------------------------------------

class _near_far_from_aabb(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        ''' rays_o/rays_d: float [N, 3]; aabb: float [6] (xmin, ymin, zmin, xmax, ymax, zmax) -> nears, fars: float [N] '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        aabb = _f32c(aabb.to(rays_o.device))
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        fars = torch.empty(N, dtype=rays_o.dtype, device=rays_o.device)
        _L.check(_L.lib().rn_near_far_from_aabb(_L.ptr(rays_o), _L.ptr(rays_d), _L.ptr(aabb), N, float(min_near),
                                                _L.ptr(nears), _L.ptr(fars), _L.cur_stream()))
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        ''' spherical coordinate on the background sphere; rays_o assumed inside Sphere(radius) -> coords [N, 2] in [-1, 1] '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=rays_o.dtype, device=rays_o.device)
        _L.check(_L.lib().rn_sph_from_ray(_L.ptr(rays_o), _L.ptr(rays_d), float(radius), N, _L.ptr(coords),
                                          _L.cur_stream()))
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        ''' coords: [N, 3] int32 in [0, 128) -> indices [N] int32 in [0, 128^3) '''
        if not coords.is_cuda: coords = coords.cuda()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        coords = coords.int().contiguous()
        _L.check(_L.lib().rn_morton3D(_L.ptr(coords), N, _L.ptr(indices), _L.cur_stream()))
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        ''' indices: [N] int32 in [0, 128^3) -> coords [N, 3] int32 in [0, 128) '''
        if not indices.is_cuda: indices = indices.cuda()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        indices = indices.int().contiguous()
        _L.check(_L.lib().rn_morton3D_invert(_L.ptr(indices), N, _L.ptr(coords), _L.cur_stream()))
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        ''' grid: float [C, H*H*H]; thresh: float -> bitfield uint8 [C*H*H*H/8] (bit i of byte n = cell 8n+i) '''
        if not grid.is_cuda: grid = grid.cuda()
        grid = grid.contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        N = C * H3 // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        _L.check(_L.lib().rn_packbits(_L.ptr(grid), N, float(thresh), _L.ptr(bitfield), _L.cur_stream()))
        return bitfield


packbits = _packbits.apply


class _morton3D_dilation(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, grid):
        ''' 6-neighbour max pooling in Morton order; grid: float [C, H*H*H] -> same shape '''
        if not grid.is_cuda: grid = grid.cuda()
        grid = grid.contiguous()
        C = grid.shape[0]
        H3 = grid.shape[1]
        H = int(np.cbrt(H3))
        if H ** 3 != H3:  # np.cbrt may land just below the integer
            H = int(round(np.cbrt(H3)))
        grid_dilation = torch.empty_like(grid)
        _L.check(_L.lib().rn_morton3D_dilation(_L.ptr(grid), C, H, _L.ptr(grid_dilation), _L.cur_stream()))
        return grid_dilation


morton3D_dilation = _morton3D_dilation.apply


# ----------------------------------------
# train functions
# ----------------------------------------

class _march_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        ''' march rays to generate points (forward only)
        Returns:
            xyzs: float [M, 3]; dirs: float [M, 3]; deltas: float [M, 2] (delta_t, t after the step)
            rays: int32 [N, 3] (ray id, point offset, point count)
        '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()
        if not density_bitfield.is_cuda: density_bitfield = density_bitfield.cuda()

        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        density_bitfield = density_bitfield.contiguous()
        nears = nears.contiguous()
        fars = fars.contiguous()

        N = rays_o.shape[0]
        M = N * max_steps  # upper bound on the number of points

        # running estimate from previous steps; rays are dropped if it turns out too small
        if not force_all_rays and mean_count > 0:
            if align > 0:
                mean_count += align - mean_count % align
            M = mean_count

        xyzs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        dirs = torch.zeros(M, 3, dtype=rays_o.dtype, device=rays_o.device)
        deltas = torch.zeros(M, 2, dtype=rays_o.dtype, device=rays_o.device)
        rays = torch.empty(N, 3, dtype=torch.int32, device=rays_o.device)  # id, offset, num_steps

        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=rays_o.device)  # point counter, ray counter

        if perturb:
            noises = torch.rand(N, dtype=rays_o.dtype, device=rays_o.device)
        else:
            noises = torch.zeros(N, dtype=rays_o.dtype, device=rays_o.device)

        _L.check(_L.lib().rn_march_rays_train(
            _L.ptr(rays_o), _L.ptr(rays_d), _L.ptr(density_bitfield), float(bound), float(dt_gamma), int(max_steps), N,
            int(C), int(H), M, _L.ptr(nears), _L.ptr(fars), _L.ptr(xyzs), _L.ptr(dirs), _L.ptr(deltas), _L.ptr(rays),
            _L.ptr(step_counter), _L.ptr(noises), _L.cur_stream()))

        # only used at the first (few) epochs.
        if force_all_rays or mean_count <= 0:
            m = step_counter[0].item()  # D2H copy
            if align > 0:
                m += align - m % align
            xyzs = xyzs[:m]
            dirs = dirs[:m]
            deltas = deltas[:m]

            torch.cuda.empty_cache()

        ctx.save_for_backward(rays, deltas)

        return xyzs, dirs, deltas, rays

    # to support optimizing camera poses.
    @staticmethod
    @_bwd
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        rays, deltas = ctx.saved_tensors
        N = rays.shape[0]
        M = grad_xyzs.shape[0]
        grad_xyzs = _f32c(grad_xyzs)
        grad_dirs = _f32c(grad_dirs)
        grad_rays_o = torch.zeros(N, 3, device=rays.device)
        grad_rays_d = torch.zeros(N, 3, device=rays.device)
        _L.check(_L.lib().rn_march_rays_train_backward(_L.ptr(grad_xyzs), _L.ptr(grad_dirs), _L.ptr(rays), _L.ptr(deltas),
                                                       N, M, _L.ptr(grad_rays_o), _L.ptr(grad_rays_d), _L.cur_stream()))
        return grad_rays_o, grad_rays_d, None, None, None, None, None, None, None, None, None, None, None, None, None


march_rays_train = _march_rays_train.apply


class _composite_rays_train(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        ''' composite rays' rgbs, according to the ray marching formula.
        Args: sigmas [M], rgbs [M, 3], ambient [M], deltas [M, 2], rays int32 [N, 3]
        Returns: weights_sum [N], ambient_sum [N], depth [N], image [N, 3]
        '''
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        ambient = ambient.contiguous()
        deltas = deltas.contiguous()
        rays = rays.contiguous()

        M = sigmas.shape[0]
        N = rays.shape[0]

        weights_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        ambient_sum = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        depth = torch.empty(N, dtype=sigmas.dtype, device=sigmas.device)
        image = torch.empty(N, 3, dtype=sigmas.dtype, device=sigmas.device)

        _L.check(_L.lib().rn_composite_rays_train_forward(
            _L.ptr(sigmas), _L.ptr(rgbs), _L.ptr(ambient), _L.ptr(deltas), _L.ptr(rays), M, N, float(T_thresh),
            _L.ptr(weights_sum), _L.ptr(ambient_sum), _L.ptr(depth), _L.ptr(image), _L.cur_stream()))

        ctx.save_for_backward(sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, depth, image)
        ctx.dims = [M, N, T_thresh]

        return weights_sum, ambient_sum, depth, image

    @staticmethod
    @_bwd
    def backward(ctx, grad_weights_sum, grad_ambient_sum, grad_depth, grad_image):
        # NOTE: grad_depth is not used (as in the reference): it is not propagated to sigmas.
        grad_weights_sum = _f32c(grad_weights_sum)
        grad_ambient_sum = _f32c(grad_ambient_sum)
        grad_image = _f32c(grad_image)

        sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh = ctx.dims

        grad_sigmas = torch.zeros_like(sigmas)
        grad_rgbs = torch.zeros_like(rgbs)
        grad_ambient = torch.zeros_like(ambient)

        _L.check(_L.lib().rn_composite_rays_train_backward(
            _L.ptr(grad_weights_sum), _L.ptr(grad_ambient_sum), _L.ptr(grad_image), _L.ptr(sigmas), _L.ptr(rgbs),
            _L.ptr(ambient), _L.ptr(deltas), _L.ptr(rays), _L.ptr(weights_sum), _L.ptr(ambient_sum), _L.ptr(image), M, N,
            float(T_thresh), _L.ptr(grad_sigmas), _L.ptr(grad_rgbs), _L.ptr(grad_ambient), _L.cur_stream()))

        return grad_sigmas, grad_rgbs, grad_ambient, None, None, None


composite_rays_train = _composite_rays_train.apply


# ----------------------------------------
# infer functions
# ----------------------------------------

class _march_rays(Function):
    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        ''' march rays to generate points (forward only, for inference)
        Returns xyzs [M, 3], dirs [M, 3], deltas [M, 2] with M = n_alive * n_step padded by `align`; slot n*n_step+k
        holds sample k of alive ray n, zero rows mean "no sample".
        '''
        if not rays_o.is_cuda: rays_o = rays_o.cuda()
        if not rays_d.is_cuda: rays_d = rays_d.cuda()

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

        _L.check(_L.lib().rn_march_rays(
            int(n_alive), int(n_step), _L.ptr(rays_alive), _L.ptr(rays_t), _L.ptr(rays_o), _L.ptr(rays_d), float(bound),
            float(dt_gamma), int(max_steps), int(C), int(H), _L.ptr(density_bitfield), _L.ptr(near), _L.ptr(far),
            _L.ptr(xyzs), _L.ptr(dirs), _L.ptr(deltas), _L.ptr(noises), _L.cur_stream()))

        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_fwd32  # need to cast sigmas & rgbs to float
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        ''' composite rays' rgbs, according to the ray marching formula. (for inference)
        In-place outputs: weights_sum [N], depth [N], image [N, 3]; rays_alive[n] = -1 marks termination.
        '''
        sigmas = sigmas.contiguous()
        rgbs = rgbs.contiguous()
        _L.check(_L.lib().rn_composite_rays(
            int(n_alive), int(n_step), float(T_thresh), _L.ptr(rays_alive), _L.ptr(rays_t), _L.ptr(sigmas), _L.ptr(rgbs),
            _L.ptr(deltas), _L.ptr(weights_sum), _L.ptr(depth), _L.ptr(image), _L.cur_stream()))
        return tuple()


composite_rays = _composite_rays.apply
