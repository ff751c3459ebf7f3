"""BASELINE configs[0] / BASELINE.md 2b: the reference's own CPU-runnable case, as a CPU baseline and as a parity + timing case.

    march N = 2^16 synthetic rays (<= 16 samples each) through the synthetic occupancy bitfield
    -> GridEncoder(3-D, 16 levels x 2 features, log2 hashmap 16, 16 -> 2048, tiled) forward + backward on the M samples, fp32
    -> composite_rays_train forward + backward

CPU leg: the oracle's C restatement of the reference kernels (oracle/oracle.c, OpenMP over all host cores; the reference has no
CPU implementation of its own), best of `reps` after one warm-up, core count printed.  GPU leg (when a CUDA device is present):
the same inputs through the C ABI of libradnerf_b200.so, checked against the CPU leg (march counters and sample floats
identical, encode forward bit-identical, table gradient and compositing <= 1e-5) and timed with CUDA events.

    python tools/config1.py [--rays 65536] [--reps 5] [--out gpurun_out/config1.json]

`run_cpu` / `compare` are also what tests/test_host_cpu.py::test_config1_case runs at a small size."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np

GRID = dict(D=3, L=16, C=2, H=16, log2_T=16, res=2048, gridtype=1)     # 'tiled'
BOUND, CASCADE, GRID_SIZE, MAX_STEPS, DT_GAMMA, MIN_NEAR = 1.0, 1, 128, 16, 1.0 / 256, 0.05


def inputs(n_rays, seed=0):
    """seeded, as BASELINE.md 2b lists them"""
    from radnerf_b200 import synthetic as syn
    rng = np.random.default_rng(seed)
    hw = int(np.ceil(np.sqrt(n_rays)))
    ro, rd = syn.get_rays(syn.orbit_pose(yaw_deg=5.0, pitch_deg=2.0), syn.intrinsics_for(hw, hw), hw, hw)
    pick = rng.permutation(hw * hw)[:n_rays]
    grid = syn.head_density_grid(GRID_SIZE, semi_axes=(0.34, 0.24, 0.37))
    from oracle import oracle as O
    offsets, pls = O.grid_offsets(GRID["D"], GRID["L"], GRID["C"], GRID["H"], GRID["log2_T"], GRID["res"])
    return dict(rays_o=np.ascontiguousarray(ro[pick]), rays_d=np.ascontiguousarray(rd[pick]),
                bitfield=syn.packbits_np(grid, min(float(np.clip(grid, 0, None).mean()), 10.0)),
                noises=rng.random(n_rays, dtype=np.float32), offsets=offsets, per_level_scale=pls,
                table=rng.uniform(-1e-4, 1e-4, (int(offsets[-1]), GRID["C"])).astype(np.float32), seed=seed)


def slot_order(rays):
    """slot of every sample in canonical (ray id, sample index) order -- the marcher's slot order is arbitrary"""
    r = rays[np.argsort(rays[:, 0], kind="stable")]
    return np.concatenate([np.arange(o, o + k) for o, k in zip(r[:, 1], r[:, 2])]) if len(r) else np.zeros(0, np.int64)


def sample_inputs(c, rays, m, seed=1):
    """what the network would hand the compositor / the encoder's backward: sigma ~ 5 U(0,1), rgb ~ U(0,1), d(features) ~ N(0,1),
    drawn per (ray id, sample index) and placed into this run's slots, so that two runs with different slot orders see the same
    values on the same samples"""
    rng = np.random.default_rng(seed)
    canon = dict(sigmas=(5.0 * rng.random(m, dtype=np.float32)), rgbs=rng.random((m, 3), dtype=np.float32),
                 ambient=rng.random(m, dtype=np.float32), d_feat=rng.standard_normal((m, GRID["L"] * GRID["C"])).astype(np.float32))
    idx = slot_order(rays)
    out = {}
    for k, v in canon.items():
        out[k] = np.zeros_like(v)
        out[k][idx] = v
    return out


def _best(fn, reps):
    fn()
    best, out = float("inf"), None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t)
    return best, out


def run_cpu(c, reps=5, scales=None):
    from oracle import oracle as O
    n = c["rays_o"].shape[0]
    aabb = np.array([-BOUND] * 3 + [BOUND] * 3, np.float32)
    t_nf, (nears, fars) = _best(lambda: O.near_far_from_aabb(c["rays_o"], c["rays_d"], aabb, MIN_NEAR), reps)
    t_march, (xyzs, dirs, deltas, rays, counter) = _best(
        lambda: O.march_rays_train(c["rays_o"], c["rays_d"], BOUND, c["bitfield"], CASCADE, GRID_SIZE, nears, fars, c["noises"],
                                   n * MAX_STEPS, DT_GAMMA, MAX_STEPS), reps)
    m = int(counter[0])
    x01 = ((xyzs[:m] + BOUND) / (2 * BOUND)).astype(np.float32)
    s = sample_inputs(c, rays, m)
    t_fwd, (feat, _) = _best(lambda: O.grid_encode_forward(x01, c["table"], c["offsets"], c["per_level_scale"], GRID["H"], False,
                                                          GRID["gridtype"], False, 0, scales=scales), reps)
    t_bwd, (g_table, _) = _best(lambda: O.grid_encode_backward(s["d_feat"], x01, c["offsets"], c["per_level_scale"], GRID["H"],
                                                              c["table"].shape[0], GRID["C"], gridtype=GRID["gridtype"], scales=scales), reps)
    t_cf, (ws, am, dp, im) = _best(lambda: O.composite_rays_train_forward(s["sigmas"], s["rgbs"], s["ambient"], deltas[:m], rays), reps)
    g_ws, g_am, g_im = np.ones(n, np.float32), np.full(n, 0.1, np.float32), np.ones((n, 3), np.float32)
    t_cb, (gs, gr, ga) = _best(lambda: O.composite_rays_train_backward(g_ws, g_am, g_im, s["sigmas"], s["rgbs"], deltas[:m], rays, ws, im), reps)
    res = dict(nears=nears, fars=fars, xyzs=xyzs[:m], dirs=dirs[:m], deltas=deltas[:m], rays=rays, counter=counter, feat=feat,
               g_table=g_table, weights_sum=ws, ambient_sum=am, depth=dp, image=im, g_sigmas=gs, g_rgbs=gr, g_ambient=ga)
    timing = {"cores": O.num_threads(), "rays": n, "samples": m,
              "march_Mrays_per_s": n / t_march / 1e6, "encode_fwd_Msamples_per_s": m / t_fwd / 1e6,
              "encode_bwd_Msamples_per_s": m / t_bwd / 1e6, "composite_fwd_Mrays_per_s": n / t_cf / 1e6,
              "composite_bwd_Mrays_per_s": n / t_cb / 1e6,
              "ms": {"near_far": t_nf * 1e3, "march": t_march * 1e3, "encode_fwd": t_fwd * 1e3, "encode_bwd": t_bwd * 1e3,
                     "composite_fwd": t_cf * 1e3, "composite_bwd": t_cb * 1e3}}
    return res, timing


def run_gpu(c, reps=20):
    """the same pipeline through the drop-in operators (C ABI), on the current CUDA device"""
    import torch
    import raymarching
    from gridencoder import GridEncoder
    dev = "cuda"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
    n = c["rays_o"].shape[0]
    ro, rd, bits = T(c["rays_o"]), T(c["rays_d"]), T(c["bitfield"])
    aabb = torch.tensor([-BOUND] * 3 + [BOUND] * 3, device=dev)
    import raymarching.raymarching as rmod
    noise = T(c["noises"])
    saved = rmod._start_offsets
    rmod._start_offsets = lambda k, perturb, like: noise[:k] if perturb else torch.zeros(k, dtype=like.dtype, device=like.device)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / 1e3, out
    try:
        t_nf, (nears, fars) = timed(lambda: raymarching.near_far_from_aabb(ro, rd, aabb, MIN_NEAR))
        counter = torch.zeros(2, dtype=torch.int32, device=dev)

        def march():
            counter.zero_()
            # mean_count = the worst case: buffers of N * max_steps rows, no ray can be dropped, no counter read-back inside the timing
            return raymarching.march_rays_train(ro, rd, BOUND, bits, CASCADE, GRID_SIZE, nears, fars, counter, n * MAX_STEPS, True, -1, False, DT_GAMMA, MAX_STEPS)
        t_march, (xyzs, dirs, deltas, rays) = timed(march)
        m = int(counter[0])
        enc = GridEncoder(input_dim=3, num_levels=GRID["L"], level_dim=GRID["C"], base_resolution=GRID["H"], log2_hashmap_size=GRID["log2_T"],
                          desired_resolution=GRID["res"], gridtype="tiled").to(dev)
        with torch.no_grad():
            enc.embeddings.copy_(T(c["table"]))
        s = {k: T(v) for k, v in sample_inputs(c, rays.cpu().numpy(), m).items()}
        x = xyzs[:m].contiguous()
        t_fwd, feat = timed(lambda: enc(x, bound=BOUND).detach())

        def bwd():
            enc.embeddings.grad = None
            enc(x, bound=BOUND).backward(s["d_feat"])
            return enc.embeddings.grad
        t_fb, g_table = timed(bwd)
        sig, rgb, amb = (s[k].clone().requires_grad_(True) for k in ("sigmas", "rgbs", "ambient"))
        t_cf, (ws, am, dp, im) = timed(lambda: tuple(t.detach() for t in raymarching.composite_rays_train(sig, rgb, amb, deltas[:m], rays)))

        def cbwd():
            for t in (sig, rgb, amb):
                t.grad = None
            w, a, _, i = raymarching.composite_rays_train(sig, rgb, amb, deltas[:m], rays)
            torch.autograd.backward([w, a, i], [torch.ones_like(w), torch.full_like(a, 0.1), torch.ones_like(i)])
            return sig.grad, rgb.grad, amb.grad
        t_cfb, (gs, gr, ga) = timed(cbwd)
    finally:
        rmod._start_offsets = saved
    N_ = lambda t: t.detach().float().cpu().numpy()   # noqa: E731
    res = dict(nears=N_(nears), fars=N_(fars), xyzs=N_(xyzs[:m]), dirs=N_(dirs[:m]), deltas=N_(deltas[:m]), rays=rays.cpu().numpy(),
               counter=counter.cpu().numpy(), feat=N_(feat), g_table=N_(g_table), weights_sum=N_(ws), ambient_sum=N_(am), depth=N_(dp),
               image=N_(im), g_sigmas=N_(gs), g_rgbs=N_(gr), g_ambient=N_(ga))
    t_bwd, t_cb = max(t_fb - t_fwd, 1e-9), max(t_cfb - t_cf, 1e-9)
    timing = {"rays": n, "samples": m, "march_Mrays_per_s": n / t_march / 1e6, "encode_fwd_Msamples_per_s": m / t_fwd / 1e6,
              "encode_bwd_Msamples_per_s": m / t_bwd / 1e6, "composite_fwd_Mrays_per_s": n / t_cf / 1e6,
              "composite_bwd_Mrays_per_s": n / t_cb / 1e6,
              "ms": {"near_far": t_nf * 1e3, "march": t_march * 1e3, "encode_fwd": t_fwd * 1e3, "encode_fwd_plus_bwd": t_fb * 1e3,
                     "composite_fwd": t_cf * 1e3, "composite_fwd_plus_bwd": t_cfb * 1e3},
              "note": "through the Python drop-in operators (allocation + autograd included); backward = (fwd+bwd) - fwd"}
    return res, timing


def canonical(res):
    """the marcher's slot order is arbitrary: order everything by ray id / (ray id, sample index)"""
    rays = res["rays"]
    order = np.argsort(rays[:, 0], kind="stable")
    r = rays[order]
    idx = slot_order(rays)
    out = {k: res[k][idx] for k in ("xyzs", "dirs", "deltas", "feat", "g_sigmas", "g_rgbs", "g_ambient")}
    out["counts"] = r[:, 2]
    for k in ("weights_sum", "ambient_sum", "depth", "image"):
        out[k] = res[k]          # per-ray outputs are addressed by RAY ID (raymarching.cu:622, 690-697), not by slot: already canonical
    for k in ("nears", "fars", "g_table", "counter"):
        out[k] = res[k]
    return out


def compare(a, b):
    """deviations between two runs of the case (canonical order); what must be exact is reported as booleans"""
    a, b = canonical(a), canonical(b)
    e = {"counter_equal": bool(np.array_equal(a["counter"], b["counter"])), "counts_equal": bool(np.array_equal(a["counts"], b["counts"]))}
    for k in ("nears", "fars", "xyzs", "dirs", "deltas", "feat"):
        e[k + "_bit_identical"] = bool(a[k].shape == b[k].shape and np.array_equal(a[k], b[k]))
    for k in ("g_table", "weights_sum", "ambient_sum", "depth", "image", "g_sigmas", "g_rgbs", "g_ambient"):
        scale = max(1.0, float(np.abs(b[k]).max()))
        e[k + "_max_rel"] = float(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)).max()) / scale if a[k].shape == b[k].shape else float("inf")
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config1.json"))
    args = ap.parse_args()
    from oracle import oracle as O
    O.build()
    O.set_num_threads(os.cpu_count())
    c = inputs(args.rays)
    line = {"workload": "BASELINE configs[0]: march 2^16 rays -> GridEncoder 3-D L16 C2 T=2^16 tiled fwd+bwd (fp32) -> composite_rays_train fwd+bwd",
            "seed": c["seed"]}
    scales = None
    try:
        import torch
        if torch.cuda.is_available():
            from radnerf_b200 import abi
            sc = torch.empty(GRID["L"], device="cuda")
            rs = torch.empty(GRID["L"], dtype=torch.int32, device="cuda")
            abi.check(abi.lib().rn_grid_level_geometry(float(np.log2(c["per_level_scale"])), GRID["H"], GRID["L"], abi.ptr(sc), abi.ptr(rs), None))
            torch.cuda.synchronize()
            scales = sc.cpu().numpy()          # the device's exp2f level scales, so that the encoders can agree bit for bit
    except ImportError:
        pass
    cpu_res, line["cpu"] = run_cpu(c, args.reps, scales=scales)
    line["cpu"]["kind"] = "oracle C restatement of the reference kernels, OpenMP"
    if scales is not None:
        gpu_res, line["gpu"] = run_gpu(c)
        line["parity_gpu_vs_cpu"] = compare(gpu_res, cpu_res)
        line["speedup"] = {k: line["gpu"][k] / line["cpu"][k] for k in line["cpu"] if k.endswith("_per_s")}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(line, open(args.out, "w"), indent=1)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
