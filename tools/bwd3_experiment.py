"""3-D table backward: which structure wins?  (VERDICT r01 task 3; csrc/grid_bwd3.cu)

Times the variants of rn_grid_backward3 on the samples a real training step scatters (march_rays_train over 2^16 random
pixels of the synthetic head scene, BASELINE configs[3]) plus, per variant, one launch per level (level_mask) to see where
the time goes.  Every variant is checked against the generic kernel (rn_grid_encode_backward) first.

    python tools/bwd3_experiment.py [out.json]          # CUDA events, 20 launches each
    PROFILE_VARIANT=0 python tools/bwd3_experiment.py   # one launch of that variant between cudaProfilerStart/Stop (for ncu)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "rad-nerf_b200"))
import numpy as np
import torch

from radnerf_b200 import abi, synthetic as syn
from radnerf_b200.model import NeRFNetwork, Options


def training_samples(dev, n_rays=65536, dense=False):
    import raymarching as rm
    torch.manual_seed(0)
    m = NeRFNetwork(Options(torso=False, fp16=True, exp_eye=True))
    grid = syn.head_density_grid(128, semi_axes=(0.34, 0.24, 0.37))
    m.density_grid.copy_(torch.from_numpy(grid))
    m.mean_density = float(np.clip(grid, 0, None).mean())
    m.density_bitfield.copy_(torch.from_numpy(syn.packbits_np(grid, min(m.mean_density, m.density_thresh))))
    m = m.to(dev)
    if dense:   # what the bench's steady-state training step sees: the occupancy grid after update_extra_state on a random-init
        # network (sigma ~ 1 everywhere, threshold = the mean): about half of ALL cells occupied, ~0.7 M samples per batch
        m.aud_features = torch.from_numpy(syn.audio_feature_bank(600, 44, 16, seed=0))
        m.eye_area = torch.full((600, 1), 0.25)
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
    b = syn.batch_to(syn.training_batch(512, 512, n_rays, frame_index=0), dev)
    ro, rd = b["rays_o"][0].contiguous(), b["rays_d"][0].contiguous()
    nears, fars = rm.near_far_from_aabb(ro, rd, m.aabb_train, m.min_near)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    xyzs, dirs, deltas, rays = rm.march_rays_train(ro, rd, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter, -1,
                                                   True, 128, False, m.opt.dt_gamma, m.opt.max_steps)
    n = int(counter[0].item())
    return m, xyzs[:n].contiguous()


def main():
    dev = torch.device("cuda", 0)
    dense = os.environ.get("DENSE", "0") == "1"
    m, xyz = training_samples(dev, dense=dense)
    enc = m.encoder
    B = xyz.shape[0]
    x01 = ((xyz + m.bound) / (2 * m.bound)).contiguous()
    g = torch.Generator(device="cpu").manual_seed(3)
    grad16 = (torch.randn(B, 32, generator=g) * 0.01).half().to(dev)
    offsets = enc.offsets
    L, S, H = 16, float(np.log2(enc.per_level_scale)), int(enc.base_resolution)
    rows = int(offsets[-1])
    lib = abi.lib()
    sizes = (offsets[1:] - offsets[:-1]).cpu().numpy()

    def generic(out):   # the round-1 structure: one v2 atomic per corner (variant 0 of the same entry point)
        abi.check(lib.rn_grid_backward3(abi.ptr(grad16), abi.ptr(x01), abi.ptr(offsets), abi.ptr(out), B, L, S, H, 1, 1, 0, 0xffff, 0, 0, 0,
                                        abi.cur_stream()))

    # privatisable: leading dense levels (row count == (res+1)^3 rounded) that fit in 150 KB
    priv_levels, priv_rows = 0, 0
    for l in range(L):
        if priv_rows + int(sizes[l]) <= 150 * 1024 // 8 and sizes[l] < 65536:
            priv_levels, priv_rows = l + 1, priv_rows + int(sizes[l])
        else:
            break

    def variant(v, out, mask=0xffff, agg=6):
        abi.check(lib.rn_grid_backward3(abi.ptr(grad16), abi.ptr(x01), abi.ptr(offsets), abi.ptr(out), B, L, S, H, 1, 1, v, mask, agg,
                                        priv_levels, priv_rows, abi.cur_stream()))

    prof = os.environ.get("PROFILE_VARIANT")
    if prof is not None:
        out = torch.zeros(rows, 2, device=dev)
        for _ in range(3):
            variant(int(prof), out) if int(prof) >= 0 else generic(out)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        variant(int(prof), out) if int(prof) >= 0 else generic(out)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("profiled variant", prof, "B", B)
        return

    ref = torch.zeros(rows, 2, device=dev)
    generic(ref)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    report = {"distribution": "dense occupancy after update_extra_state (bench steady state)" if dense else "head-shaped occupancy", "samples": B, "rows": rows, "priv_levels": priv_levels, "priv_rows": priv_rows, "ref_absmax": scale, "variants": {}}

    def timeit(fn, reps=20):
        out = torch.zeros(rows, 2, device=dev)
        for _ in range(3):
            fn(out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(out)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    report["generic_ms"] = timeit(generic)

    def production(out):
        abi.call("rn_grid_encode_backward", grad16, x01, None, offsets, out, B, 3, 2, L, S, H, None, None, 1, 0, 0, 1, 1, 0)
    report["production_ms"] = timeit(production)
    report["production_gbs_algorithmic"] = B * 1100 / report["production_ms"] / 1e6
    names = {0: "plain (8 x v2 per level)", 1: "z-merge + x-pair v4", 3: "z-merge + x-pair + segmented warp aggregation",
             5: "z-merge + x-pair + smem privatisation", 7: "all three"}
    for v, name in names.items():
        out = torch.zeros(rows, 2, device=dev)
        variant(v, out)
        torch.cuda.synchronize()
        err = (out - ref).abs().max().item() / scale
        entry = {"name": name, "rel_err_vs_generic": err, "ms": timeit(lambda o: variant(v, o))}
        if v in (0, 1, 3):
            entry["ms_per_level"] = [timeit(lambda o, l=l: variant(v, o, mask=1 << l), reps=10) for l in range(L)]
        if v == 3:
            entry["ms_by_agg_levels"] = {str(a): timeit(lambda o, a=a: variant(v, o, agg=a)) for a in (2, 4, 6, 8, 10)}
        report["variants"][str(v)] = entry
        print(v, name, json.dumps(entry))
    # distinct rows touched per level (how hot the addresses are)
    touched = []
    for l in range(L):
        out = torch.zeros(rows, 2, device=dev)
        variant(0, out, mask=1 << l)
        touched.append(int((out.abs().sum(1) > 0).sum().item()))
    report["rows_touched_per_level"] = touched
    report["level_sizes"] = [int(s) for s in sizes]
    print(json.dumps(report))
    if len(sys.argv) > 1:
        json.dump(report, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
