"""TEST INFRASTRUCTURE -- the reference's op-by-op INFERENCE loop (nerf/renderer.py:229-262), restated for any operator bundle.

`NeRFRenderer.run_cuda` renders a frame with a host loop: while rays are alive, march them `n_step = clamp(N // n_alive, 1, 8)` samples
forward (`march_rays`), evaluate the network on the samples, composite (`composite_rays`), drop the terminated rays -- one host
synchronisation per iteration for the new length of the alive list.  The product replaces this loop with the device-driven fused frame
(radnerf_b200/frame.py, csrc/frame_ctl.cu) and never runs it; the parity tests and the CPU port do, through the hook
`radnerf_b200.model.register_ops_frame`: importing this module installs it.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs import oracle/ (tests/test_abi.py enforces that the product does not).
"""
import torch


def ops_frame(model, rays_o, rays_d, nears, fars, enc_a, ind_code, eye, dt_gamma, perturb, max_steps, T_thresh):
    """-> weights_sum [N], depth [N], image [N,3] accumulated over the loop; also leaves `model.last_frame_stats` =
    [(n_alive, n_step, sample slots)] per iteration (what the fused frame's device-side controller must reproduce)"""
    rm = model.ops.rm
    N, device = rays_o.shape[0], rays_o.device
    weights_sum = torch.zeros(N, dtype=torch.float32, device=device)          # renderer.py:243-245
    depth = torch.zeros(N, dtype=torch.float32, device=device)
    image = torch.zeros(N, 3, dtype=torch.float32, device=device)
    rays_alive = torch.arange(N, dtype=torch.int32, device=device)            # renderer.py:247-249
    rays_t = nears.clone()
    step = 0
    model.last_frame_stats = []
    while step < max_steps:                                                   # renderer.py:251-262
        n_alive = rays_alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, model.bound, model.density_bitfield,
                                           model.cascade, model.grid_size, nears, fars, 128, perturb if step == 0 else False, dt_gamma, max_steps)
        sigmas, rgbs, ambient = model(xyzs, dirs, enc_a, ind_code, eye)
        sigmas = model.density_scale * sigmas
        rm.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
        rays_alive = rays_alive[rays_alive >= 0]      # host sync: the new length is needed on the CPU
        model.last_frame_stats.append((n_alive, n_step, xyzs.shape[0]))
        step += n_step
    return weights_sum, depth, image


def install():
    from radnerf_b200 import model as _model
    _model.register_ops_frame(ops_frame)


install()
