"""Occupancy-grid utilities, ray marching and compositing on libradnerf_b200 (sm_100a).

The ten module-level functions of the reference's raymarching/raymarching.py, with its positional signatures, defaults, return
shapes / dtypes and in-place semantics:

    near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2)                                   -> nears [N], fars [N]
    sph_from_ray(rays_o, rays_d, radius)                                                     -> coords [N, 2]
    morton3D(coords) / morton3D_invert(indices)                                              -> int32
    packbits(grid, thresh, bitfield=None)                                                    -> uint8 [C*H^3/8]
    morton3D_dilation(grid)                                                                  -> grid
    march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024) -> xyzs, dirs, deltas, rays
    composite_rays_train(sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4)                 -> weights_sum, ambient_sum, depth, image
    march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
               perturb=False, dt_gamma=0, max_steps=1024)                                    -> xyzs, dirs, deltas
    composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2) -> ()

Host-side rules kept from the reference because callers depend on them: float arguments are up-cast to fp32 even under
autocast; sample buffers are zero-filled (an all-zero row means "no sample"); the `align` padding always ADDS a block,
also when the count is already a multiple (raymarching.py:227-229, 251-253, 382-383); the training marcher sizes its buffers
from `mean_count` once that is known and trims them with one device->host read of the counter before.  Every call crosses
the C ABI (include/radnerf_b200.h) with raw device pointers on torch's current stream."""
import torch
from torch.amp import custom_bwd, custom_fwd

from radnerf_b200 import abi

__all__ = ['near_far_from_aabb', 'sph_from_ray', 'morton3D', 'morton3D_invert', 'packbits', 'morton3D_dilation',
           'march_rays_train', 'composite_rays_train', 'march_rays', 'composite_rays']

_as_f32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_grad = custom_bwd(device_type="cuda")


def _dev(t):
    return t if t.is_cuda else t.cuda()


def _rays(t):
    """[..., 3] -> contiguous [N, 3] on the device"""
    return _dev(t).contiguous().view(-1, 3)


def _padded(count, align):
    return count + (align - count % align) if align > 0 else count


def _start_offsets(n, perturb, like):
    """per-ray jitter of the first step in units of dt: U[0,1) when perturbing, else zeros"""
    return torch.rand(n, dtype=like.dtype, device=like.device) if perturb else torch.zeros(n, dtype=like.dtype, device=like.device)


def _sample_buffers(m, like):
    z = lambda w: torch.zeros(m, w, dtype=like.dtype, device=like.device)   # noqa: E731
    return z(3), z(3), z(2)


# ------------------------------------------------------------------------------------------------ geometry / occupancy
class NearFarFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        o, d = _rays(rays_o), _rays(rays_d)
        box = aabb.to(o.device, torch.float32).contiguous()
        near, far = o.new_empty(o.shape[0]), o.new_empty(o.shape[0])
        abi.call("rn_near_far_from_aabb", o, d, box, o.shape[0], float(min_near), near, far)
        return near, far


class SphFromRayFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, rays_o, rays_d, radius):
        o, d = _rays(rays_o), _rays(rays_d)
        uv = o.new_empty(o.shape[0], 2)
        abi.call("rn_sph_from_ray", o, d, float(radius), o.shape[0], uv)
        return uv


class MortonFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, coords):
        c = _dev(coords).int().contiguous()
        code = torch.empty(c.shape[0], dtype=torch.int32, device=c.device)
        abi.call("rn_morton3D", c, c.shape[0], code)
        return code


class MortonInvertFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, indices):
        code = _dev(indices).int().contiguous()
        c = torch.empty(code.shape[0], 3, dtype=torch.int32, device=code.device)
        abi.call("rn_morton3D_invert", code, code.shape[0], c)
        return c


class PackbitsFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, grid, thresh, bitfield=None):
        g = _dev(grid).contiguous()
        n_bytes = g.shape[0] * g.shape[1] // 8
        bits = bitfield if bitfield is not None else torch.empty(n_bytes, dtype=torch.uint8, device=g.device)
        abi.call("rn_packbits", g, n_bytes, float(thresh), bits)      # bit i of byte n <-> cell 8n + i
        if bitfield is not None:
            # the kernel wrote through the raw pointer: tell autograd (and everything that caches on `_version`, e.g. the fused
            # renderer's occupied-cell box) that the caller's tensor changed
            ctx.mark_dirty(bitfield)
            torch.autograd.graph.increment_version(bitfield)
        return bits


class DilationFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, grid):
        g = _dev(grid).contiguous()
        cells = g.shape[1]
        side = round(cells ** (1.0 / 3.0))
        if side ** 3 != cells:
            raise RuntimeError(f"morton3D_dilation: {cells} cells per cascade is not a cube")
        out = torch.empty_like(g)
        abi.call("rn_morton3D_dilation", g, g.shape[0], side, out)
        return out


near_far_from_aabb = NearFarFn.apply
sph_from_ray = SphFromRayFn.apply
morton3D = MortonFn.apply
morton3D_invert = MortonInvertFn.apply
packbits = PackbitsFn.apply
morton3D_dilation = DilationFn.apply


# ------------------------------------------------------------------------------------------------ training
class sample_budget:
    """`with sample_budget(capacity, budget):` -- inside, march_rays_train allocates `capacity` sample slots whatever `mean_count` is
    and drops a ray when it would end past min(capacity, budget[0]), `budget` being an int32 tensor ON THE DEVICE
    (rn_march_rays_train_budget).  Same rays kept / dropped as the reference with mean_count = budget[0], but the shapes of a step
    no longer depend on the running estimate: radnerf_b200.train.GraphedTrainStep replays one captured graph across occupancy
    updates and only rewrites `budget`.  Not part of the reference's API; the default behaviour is untouched."""
    active = None

    def __init__(self, capacity, budget):
        if budget.dtype != torch.int32 or not budget.is_cuda:
            raise TypeError("sample_budget: budget must be an int32 CUDA tensor")
        self.capacity, self.budget = int(capacity), budget

    def __enter__(self):
        self.previous, sample_budget.active = sample_budget.active, self
        return self

    def __exit__(self, *exc):
        sample_budget.active = self.previous


class MarchTrainFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1, perturb=False,
                align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        o, d = _rays(rays_o), _rays(rays_d)
        bits = _dev(density_bitfield).contiguous()
        n = o.shape[0]
        # capacity: the running estimate of earlier steps when there is one (rays are dropped if it is too small),
        # otherwise the worst case, trimmed after the launch with one counter read-back
        estimated = (not force_all_rays) and mean_count > 0
        fixed = sample_budget.active if estimated else None
        capacity = (fixed.capacity if fixed is not None else _padded(mean_count, align)) if estimated else n * max_steps
        xyzs, dirs, deltas = _sample_buffers(capacity, o)
        rays = torch.empty(n, 3, dtype=torch.int32, device=o.device)              # (ray id, first sample, sample count)
        counter = step_counter if step_counter is not None else torch.zeros(2, dtype=torch.int32, device=o.device)
        abi.call("rn_march_rays_train_budget", o, d, bits, float(bound), float(dt_gamma), int(max_steps), n, int(C), int(H), capacity,
                 None if fixed is None else fixed.budget, nears.contiguous(), fars.contiguous(), xyzs, dirs, deltas, rays, counter,
                 _start_offsets(n, perturb, o))
        if not estimated:
            used = _padded(int(counter[0].item()), align)
            xyzs, dirs, deltas = xyzs[:used], dirs[:used], deltas[:used]
            torch.cuda.empty_cache()
        ctx.save_for_backward(rays, deltas)
        return xyzs, dirs, deltas, rays

    @staticmethod
    @_grad
    def backward(ctx, d_xyzs, d_dirs, d_deltas, d_rays):   # only needed when camera poses are optimised
        rays, deltas = ctx.saved_tensors
        n = rays.shape[0]
        d_o, d_d = torch.zeros(n, 3, device=rays.device), torch.zeros(n, 3, device=rays.device)
        abi.call("rn_march_rays_train_backward", d_xyzs.float().contiguous(), d_dirs.float().contiguous(), rays, deltas, n,
                 d_xyzs.shape[0], d_o, d_d)
        return (d_o, d_d) + (None,) * 13


class CompositeTrainFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
        sigmas, rgbs, ambient, deltas, rays = (t.contiguous() for t in (sigmas, rgbs, ambient, deltas, rays))
        m, n = sigmas.shape[0], rays.shape[0]
        weights_sum, ambient_sum, depth = sigmas.new_empty(n), sigmas.new_empty(n), sigmas.new_empty(n)
        image = sigmas.new_empty(n, 3)
        abi.call("rn_composite_rays_train_forward", sigmas, rgbs, ambient, deltas, rays, m, n, float(T_thresh), weights_sum,
                 ambient_sum, depth, image)
        ctx.save_for_backward(sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image)
        ctx.T_thresh = float(T_thresh)
        return weights_sum, ambient_sum, depth, image

    @staticmethod
    @_grad
    def backward(ctx, d_weights_sum, d_ambient_sum, d_depth, d_image):   # depth carries no gradient, as in the reference
        sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image = ctx.saved_tensors
        d_sigmas, d_rgbs, d_ambient = torch.zeros_like(sigmas), torch.zeros_like(rgbs), torch.zeros_like(ambient)
        abi.call("rn_composite_rays_train_backward", d_weights_sum.float().contiguous(), d_ambient_sum.float().contiguous(),
                 d_image.float().contiguous(), sigmas, rgbs, ambient, deltas, rays, weights_sum, ambient_sum, image, sigmas.shape[0],
                 rays.shape[0], ctx.T_thresh, d_sigmas, d_rgbs, d_ambient)
        return d_sigmas, d_rgbs, d_ambient, None, None, None


march_rays_train = MarchTrainFn.apply
composite_rays_train = CompositeTrainFn.apply


# ------------------------------------------------------------------------------------------------ inference
class MarchFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far, align=-1,
                perturb=False, dt_gamma=0, max_steps=1024):
        o, d = _rays(rays_o), _rays(rays_d)
        xyzs, dirs, deltas = _sample_buffers(_padded(n_alive * n_step, align), o)      # slot n*n_step+k = sample k of alive ray n
        abi.call("rn_march_rays", int(n_alive), int(n_step), rays_alive, rays_t, o, d, float(bound), float(dt_gamma), int(max_steps),
                 int(C), int(H), density_bitfield, near, far, xyzs, dirs, deltas, _start_offsets(n_alive, perturb, o))
        return xyzs, dirs, deltas


class CompositeFn(torch.autograd.Function):
    @staticmethod
    @_as_f32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
        # accumulates into weights_sum / depth / image in place; rays_alive[n] = -1 marks a terminated ray
        abi.call("rn_composite_rays", int(n_alive), int(n_step), float(T_thresh), rays_alive, rays_t, sigmas.contiguous(),
                 rgbs.contiguous(), deltas, weights_sum, depth, image)
        return tuple()


march_rays = MarchFn.apply
composite_rays = CompositeFn.apply
