"""numpy front-end of the CPU oracle (oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs import this module.
Nothing under rad-nerf_b200/ does: the product has no CPU path.

Every function takes/returns numpy arrays and restates one reference kernel (see oracle.c for file:line).
Parity status: pinned against golden vectors produced by the reference's own CUDA kernels on a B200
(tests/golden/, generator oracle/make_golden.py, checked by tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"] + (["-B"] if force else []), check=True,
                       capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.o_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return int(lib().o_num_threads())


def set_num_threads(n):
    lib().o_set_num_threads(C.c_int(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


u32, f32c, i32c = C.c_uint32, C.c_float, C.c_int

# ------------------------------------------------------------------------------------------------ grid


def grid_offsets(input_dim, num_levels, level_dim, base_resolution, log2_hashmap_size, desired_resolution=None,
                 per_level_scale=2.0, align_corners=False):
    """Level table of GridEncoder.__init__ (gridencoder/grid.py:100-131). Returns (offsets int32[L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        params = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        params = int(np.ceil(params / 8) * 8)
        offsets.append(offset)
        offset += params
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def grid_level_geometry(S, H, L):
    sc = np.empty(L, np.float32)
    rs = np.empty(L, np.uint32)
    lib().o_grid_level_geometry(f32c(S), u32(H), u32(L), _p(sc), _p(rs))
    return sc, rs


def grid_encode_forward(inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False,
                        gridtype=0, align_corners=False, interpolation=0, scales=None):
    """_grid_encode.forward (gridencoder/grid.py:27-63) incl. the permute to [B, L*C].  embeddings float32 or float16
    decides the arithmetic.  Returns (outputs [B, L*C], dy_dx [B, L*D*C] or None)."""
    inputs = _f32(inputs)
    B, D = inputs.shape
    half = embeddings.dtype == np.float16
    emb = np.ascontiguousarray(embeddings)
    offsets = _i32(offsets)
    L, Cc = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    out = np.empty((L, B, Cc), emb.dtype)
    dy = np.empty((B, L * D * Cc), emb.dtype) if calc_grad_inputs else None
    sc = None if scales is None else _f32(scales)
    lib().o_grid_encode_forward(_p(inputs), _p(emb), _p(offsets), _p(out), u32(B), u32(D), u32(Cc), u32(L), f32c(S),
                                u32(base_resolution), _p(dy), u32(gridtype), i32c(int(align_corners)),
                                u32(interpolation), i32c(int(half)), _p(sc))
    return np.ascontiguousarray(out.transpose(1, 0, 2)).reshape(B, L * Cc), dy


def grid_encode_backward(grad, inputs, offsets, per_level_scale, base_resolution, n_rows, level_dim, dy_dx=None,
                         gridtype=0, align_corners=False, interpolation=0, scales=None):
    """_grid_encode.backward (gridencoder/grid.py:65-89).  grad [B, L*C] (float32 or float16).
    Returns (grad_embeddings float64 [rows, C], grad_inputs float64 [B, D] or None)."""
    inputs = _f32(inputs)
    B, D = inputs.shape
    offsets = _i32(offsets)
    L, Cc = offsets.shape[0] - 1, level_dim
    half = grad.dtype == np.float16
    g = np.ascontiguousarray(grad.reshape(B, L, Cc).transpose(1, 0, 2))  # grid.py:75
    S = np.float32(np.log2(per_level_scale))
    gg = np.zeros((n_rows, Cc), np.float64)
    sc = None if scales is None else _f32(scales)
    lib().o_grid_encode_backward(_p(g), _p(inputs), _p(offsets), _p(gg), u32(B), u32(D), u32(Cc), u32(L), f32c(S),
                                 u32(base_resolution), u32(gridtype), i32c(int(align_corners)), u32(interpolation),
                                 i32c(int(half)), _p(sc))
    gi = None
    if dy_dx is not None:
        gi = np.zeros((B, D), np.float64)
        dy = np.ascontiguousarray(dy_dx)
        lib().o_grid_input_backward(_p(g), _p(dy), _p(gi), u32(B), u32(D), u32(Cc), u32(L), i32c(int(half)))
    return gg, gi


def grad_total_variation(inputs, embeddings, offsets, per_level_scale, base_resolution, weight, gridtype=0,
                         align_corners=False, scales=None):
    inputs = _f32(inputs)
    B, D = inputs.shape
    emb = _f32(embeddings)
    offsets = _i32(offsets)
    L, Cc = offsets.shape[0] - 1, emb.shape[1]
    S = np.float32(np.log2(per_level_scale))
    g = np.zeros(emb.shape, np.float64)
    sc = None if scales is None else _f32(scales)
    lib().o_grad_total_variation(_p(inputs), _p(emb), _p(g), _p(offsets), f32c(weight), u32(B), u32(D), u32(Cc), u32(L),
                                 f32c(S), u32(base_resolution), u32(gridtype), i32c(int(align_corners)), _p(sc))
    return g


# ------------------------------------------------------------------------------------------------ raymarching utils


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().o_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32c(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().o_sph_from_ray(_p(rays_o), _p(rays_d), f32c(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords)
    N = coords.shape[0]
    out = np.empty(N, np.int32)
    lib().o_morton3D(_p(coords), u32(N), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    N = indices.shape[0]
    out = np.empty((N, 3), np.int32)
    lib().o_morton3D_invert(_p(indices), u32(N), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().o_packbits(_p(grid), u32(N), f32c(thresh), _p(out))
    return out


def morton3D_dilation(grid):
    grid = _f32(grid)
    Cc, H3 = grid.shape
    H = int(round(H3 ** (1 / 3)))
    out = np.empty_like(grid)
    lib().o_morton3D_dilation(_p(grid), u32(Cc), u32(H), _p(out))
    return out


# ------------------------------------------------------------------------------------------------ training


def march_rays_train(rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, noises, M, dt_gamma=0.0, max_steps=1024,
                     counter=None):
    """kernel_march_rays_train with rays visited in order.  Returns xyzs[M,3], dirs[M,3], deltas[M,2], rays[N,3], counter[2]."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32) if counter is None else counter
    nears, fars, noises = _f32(nears), _f32(fars), _f32(noises)
    lib().o_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32c(bound), f32c(dt_gamma), u32(max_steps), u32(N),
                             u32(Cc), u32(H), u32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays),
                             _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def march_rays_train_backward(grad_xyzs, grad_dirs, rays, deltas):
    grad_xyzs, grad_dirs, deltas, rays = _f32(grad_xyzs), _f32(grad_dirs), _f32(deltas), _i32(rays)
    N, M = rays.shape[0], grad_xyzs.shape[0]
    go, gd = np.zeros((N, 3), np.float32), np.zeros((N, 3), np.float32)
    lib().o_march_rays_train_backward(_p(grad_xyzs), _p(grad_dirs), _p(rays), _p(deltas), u32(N), u32(M), _p(go), _p(gd))
    return go, gd


def composite_rays_train_forward(sigmas, rgbs, ambient, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, ambient, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(ambient), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, am, dp, im = (np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros(N, np.float32),
                      np.zeros((N, 3), np.float32))
    lib().o_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(ambient), _p(deltas), _p(rays), u32(M), u32(N),
                                         f32c(T_thresh), _p(ws), _p(am), _p(dp), _p(im))
    return ws, am, dp, im


def composite_rays_train_backward(g_ws, g_amb, g_img, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4):
    g_ws, g_amb, g_img = _f32(g_ws), _f32(g_amb), _f32(g_img)
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    weights_sum, image = _f32(weights_sum), _f32(image)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gr, ga = np.zeros(M, np.float32), np.zeros((M, 3), np.float32), np.zeros(M, np.float32)
    lib().o_composite_rays_train_backward(_p(g_ws), _p(g_amb), _p(g_img), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                          _p(weights_sum), _p(image), u32(M), u32(N), f32c(T_thresh), _p(gs), _p(gr),
                                          _p(ga))
    return gs, gr, ga


# ------------------------------------------------------------------------------------------------ inference


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, Cc, H, nears, fars, align=-1,
               noises=None, dt_gamma=0.0, max_steps=1024):
    """_march_rays.forward (raymarching/raymarching.py:348-397) + kernel_march_rays."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    noises = np.zeros(n_alive, np.float32) if noises is None else _f32(noises)
    rays_alive, rays_t = _i32(rays_alive), _f32(rays_t)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    nears, fars = _f32(nears), _f32(fars)
    lib().o_march_rays(u32(n_alive), u32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d), f32c(bound),
                       f32c(dt_gamma), u32(max_steps), u32(Cc), u32(H), _p(bitfield), _p(nears), _p(fars), _p(xyzs),
                       _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """kernel_composite_rays; all state arrays are updated IN PLACE (must be contiguous, right dtype)."""
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                  (image, np.float32)):
        assert a.dtype == dt and a.flags.c_contiguous
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    lib().o_composite_rays(u32(n_alive), u32(n_step), f32c(T_thresh), _p(rays_alive), _p(rays_t), _p(sigmas), _p(rgbs),
                           _p(deltas), _p(weights_sum), _p(depth), _p(image))


# ------------------------------------------------------------------------------------------------ freq / SH


def freq_encode_forward(inputs, degree):
    inputs = _f32(inputs)
    B, D = inputs.shape
    Cc = D + 2 * D * degree
    out = np.empty((B, Cc), np.float32)
    lib().o_freq_encode_forward(_p(inputs), u32(B), u32(D), u32(degree), u32(Cc), _p(out))
    return out


def freq_encode_backward(grad, outputs, D, degree):
    grad, outputs = _f32(grad), _f32(outputs)
    B, Cc = grad.shape
    gi = np.empty((B, D), np.float32)
    lib().o_freq_encode_backward(_p(grad), _p(outputs), u32(B), u32(D), u32(degree), u32(Cc), _p(gi))
    return gi


def sh_encode_forward(inputs, degree, calc_grad_inputs=False):
    inputs = _f32(inputs)
    B, D = inputs.shape
    out = np.empty((B, degree * degree), np.float32)
    dy = np.empty((B, D * degree * degree), np.float32) if calc_grad_inputs else None
    lib().o_sh_encode_forward(_p(inputs), _p(out), u32(B), u32(D), u32(degree), _p(dy))
    return out, dy


def sh_encode_backward(grad, dy_dx, D, degree):
    grad, dy_dx = _f32(grad), _f32(dy_dx)
    B = grad.shape[0]
    gi = np.zeros((B, D), np.float32)
    lib().o_sh_encode_backward(_p(grad), u32(B), u32(D), u32(degree), _p(dy_dx), _p(gi))
    return gi


def adam_step(p, g, m, v, lr, betas, eps, weight_decay, step, inv_scale=1.0):
    """one torch.optim.Adam step on flat fp32 arrays, IN PLACE on p, m, v (step is 1-based); see oracle.c o_adam_step"""
    for a in (p, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    g = _f32(g)
    lib().o_adam_step(_p(p), _p(g), _p(m), _p(v), C.c_uint64(p.size), C.c_double(lr), C.c_double(betas[0]),
                      C.c_double(betas[1]), C.c_double(eps), C.c_double(weight_decay), C.c_double(step), C.c_float(inv_scale))


def ema_update(shadow, p, decay):
    """torch_ema update IN PLACE on shadow"""
    assert shadow.dtype == np.float32 and shadow.flags.c_contiguous
    lib().o_ema_update(_p(shadow), _p(_f32(p)), C.c_uint64(shadow.size), C.c_double(decay))
