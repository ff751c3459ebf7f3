"""Seeded input cases shared by oracle/make_golden.py (which runs the REFERENCE kernels on them on a B200 and stores
the outputs under tests/golden/) and by the tests (which run the CPU oracle / our CUDA path on the same inputs).

Inputs are regenerated from numpy's PCG64 (bit-stable across machines); only outputs are stored in the fixtures.
"""
import os
import sys

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (_ROOT, os.path.join(_ROOT, "rad-nerf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from radnerf_b200 import synthetic as syn  # noqa: E402  (input generators only)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def level_offsets(D, L, C, H, log2_hashmap_size, desired_resolution=None, per_level_scale=2.0, align_corners=False):
    """GridEncoder.__init__ level table (gridencoder/grid.py:100-131)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / H) / (L - 1))
    offsets, offset = [], 0
    for i in range(L):
        res = int(np.ceil(H * per_level_scale ** i))
        n = min(2 ** log2_hashmap_size, (res if align_corners else res + 1) ** D)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, np.int32), float(per_level_scale)


# name: D, L, C, H, log2T, desired_res, gridtype(0 hash,1 tiled), align, interp, half, B, dy, store_bwd_full
GRID_CASES = {
    "g3_f32":        dict(D=3, L=16, C=2, H=16, T=16, res=2048, gridtype=1, align=False, interp=0, half=False, B=1024, dy=False),
    "g3_f16":        dict(D=3, L=16, C=2, H=16, T=16, res=2048, gridtype=1, align=False, interp=0, half=True, B=1024, dy=False),
    "g2_f16_dy":     dict(D=2, L=16, C=2, H=16, T=16, res=2048, gridtype=1, align=False, interp=0, half=True, B=512, dy=True),
    "g2_f32_dy":     dict(D=2, L=16, C=2, H=16, T=16, res=2048, gridtype=1, align=False, interp=0, half=False, B=512, dy=True),
    "g3_hash_sm_f32": dict(D=3, L=8, C=4, H=8, T=12, res=256, gridtype=0, align=False, interp=1, half=False, B=512, dy=True),
    "g3_hash_f16":   dict(D=3, L=8, C=2, H=8, T=12, res=256, gridtype=0, align=False, interp=0, half=True, B=512, dy=True),
    "g2_align_c1":   dict(D=2, L=6, C=1, H=4, T=10, res=None, gridtype=1, align=True, interp=0, half=False, B=512, dy=True),
    "g4_f32":        dict(D=4, L=4, C=2, H=4, T=10, res=None, gridtype=0, align=False, interp=0, half=False, B=256, dy=False),
    "g3_c8_f16":     dict(D=3, L=4, C=8, H=8, T=11, res=None, gridtype=1, align=False, interp=1, half=True, B=256, dy=True),
}


def grid_case(name):
    c = dict(GRID_CASES[name])
    seed = sum(ord(ch) for ch in name)
    rng = np.random.default_rng(1000 + seed)
    offsets, pls = level_offsets(c["D"], c["L"], c["C"], c["H"], c["T"], c["res"], 2.0, c["align"])
    B, D = c["B"], c["D"]
    x = rng.random((B, D), dtype=np.float32)
    x[0] = -0.25  # out of range on every axis
    x[1, 0] = 1.5  # out of range on one axis
    x[2] = 0.0
    x[3] = 1.0
    x[4] = 0.5
    table = rng.uniform(-1.0, 1.0, (int(offsets[-1]), c["C"])).astype(np.float32)
    grad = (rng.standard_normal((B, c["L"] * c["C"])) * 0.1).astype(np.float32)
    if c["half"]:
        table = table.astype(np.float16)
        grad = grad.astype(np.float16)
    c.update(name=name, offsets=offsets, per_level_scale=pls, inputs=x, table=table, grad=grad,
             bwd_rows=np.sort(rng.choice(int(offsets[-1]), size=min(2048, int(offsets[-1])), replace=False)))
    return c


# --------------------------------------------------------------------------------------------- scene for the marchers
def scene(hw=48, yaw=7.0, cascades=1, bound=1.0):
    pose = syn.orbit_pose(yaw_deg=yaw, pitch_deg=3.0)
    intr = syn.intrinsics_for(hw, hw)
    rays_o, rays_d = syn.get_rays(pose, intr, hw, hw)
    H = 128 if cascades == 1 else 64
    if cascades == 1:
        grid = syn.head_density_grid(H)
    else:  # several cascades: a head in cascade 0 and random blobs in the coarser ones
        rng = np.random.default_rng(77)
        grid = np.concatenate([syn.head_density_grid(H)] +
                              [(rng.random((1, H ** 3)) > 0.97).astype(np.float32) * 20 for _ in range(cascades - 1)], 0)
    bitfield = syn.packbits_np(grid, 10.0)
    aabb = np.array([-bound, -bound / 2, -bound, bound, bound / 2, bound], np.float32)
    return dict(rays_o=rays_o, rays_d=rays_d, bitfield=bitfield, aabb=aabb, H=H, C=cascades, bound=bound, grid=grid)


MARCH_CASES = {
    # name: scene args, marcher args
    "m_head": dict(scene=dict(hw=48, yaw=7.0, cascades=1, bound=1.0), dt_gamma=1.0 / 256, max_steps=16, min_near=0.05),
    "m_casc": dict(scene=dict(hw=32, yaw=-20.0, cascades=3, bound=4.0), dt_gamma=0.0, max_steps=48, min_near=0.2),
    "m_cone": dict(scene=dict(hw=32, yaw=33.0, cascades=2, bound=2.0), dt_gamma=1.0 / 128, max_steps=64, min_near=0.05),
}


def march_case(name):
    c = dict(MARCH_CASES[name])
    sc = scene(**c["scene"])
    seed = sum(ord(ch) for ch in name)
    rng = np.random.default_rng(2000 + seed)
    N = sc["rays_o"].shape[0]
    c.update(sc)
    c["name"] = name
    c["N"] = N
    c["noises"] = rng.random(N, dtype=np.float32)
    c["M"] = N * c["max_steps"]
    # per-sample network outputs for the compositors (drawn for the maximum possible M, sliced by the consumer)
    Mmax = c["M"] + 256
    sig = rng.random(Mmax, dtype=np.float32) * 5.0
    opaque = rng.random(Mmax) < 0.08
    sig[opaque] *= 60.0  # some very dense samples so the T < T_thresh early exit triggers
    c["sigmas"] = sig
    c["rgbs"] = rng.random((Mmax, 3), dtype=np.float32)
    c["ambient"] = rng.random(Mmax, dtype=np.float32)
    c["g_ws"] = rng.standard_normal(N).astype(np.float32)
    c["g_amb"] = rng.standard_normal(N).astype(np.float32)
    c["g_img"] = rng.standard_normal((N, 3)).astype(np.float32)
    # inference: a shuffled subset of the rays is alive
    n_alive = (N * 2) // 3
    c["n_alive"] = n_alive
    c["n_step"] = 3
    c["rays_alive"] = rng.permutation(N)[:n_alive].astype(np.int32)
    c["infer_noises"] = rng.random(n_alive, dtype=np.float32)
    c["ws0"] = (rng.random(N, dtype=np.float32) * 0.3)
    c["depth0"] = rng.random(N, dtype=np.float32)
    c["image0"] = rng.random((N, 3), dtype=np.float32) * 0.3
    return c


def util_case():
    rng = np.random.default_rng(31337)
    c = {}
    c["coords"] = rng.integers(0, 128, (4096, 3)).astype(np.int32)
    c["coords"][:4] = [[1, 0, 0], [0, 1, 0], [0, 0, 1], [127, 127, 127]]
    c["indices"] = rng.integers(0, 2 ** 21, 4096).astype(np.int32)
    g = rng.random((1, 32 ** 3), dtype=np.float32) * 20
    g[0, rng.integers(0, 32 ** 3, 500)] = -1.0
    g[0, :8] = [10.0, 10.000001, 9.999999, 0, -1, 20, 10, 11]
    c["grid32"] = g
    c["grid16x2"] = rng.random((2, 16 ** 3), dtype=np.float32) * 3
    c["thresh"] = 10.0
    c["sph_o"] = (rng.standard_normal((256, 3)) * 0.2).astype(np.float32)
    d = rng.standard_normal((256, 3))
    c["sph_d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    c["radius"] = 2.5
    return c


def enc_case():
    rng = np.random.default_rng(4242)
    c = {}
    c["freq2"] = (rng.random((256, 2), dtype=np.float32) * 1.6 - 0.8)
    c["freq6"] = np.concatenate([rng.random((64, 3), dtype=np.float32) * 3.2 - 1.6,
                                 rng.random((64, 3), dtype=np.float32) * 7 - 3.5], 1)
    c["gfreq2"] = rng.standard_normal((256, 42)).astype(np.float32)
    c["gfreq6"] = rng.standard_normal((64, 54)).astype(np.float32)
    d = rng.standard_normal((256, 3))
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    d[:8] *= rng.random((8, 1)) + 0.2  # a few non-unit inputs
    c["dirs"] = d.astype(np.float32)
    c["gsh4"] = rng.standard_normal((256, 16)).astype(np.float32)
    c["gsh8"] = rng.standard_normal((256, 64)).astype(np.float32)
    return c


def canonical_rays(rays, xyzs, dirs, deltas, M):
    """Order-independent view of a march_rays_train result: per-ray counts (by ray id), 'kept' flags, and the kept
    rays' samples concatenated in ray-id order."""
    rays = np.asarray(rays)
    order = np.argsort(rays[:, 0], kind="stable")
    r = rays[order]
    counts = r[:, 2].astype(np.int32)
    kept = (counts > 0) & (r[:, 1].astype(np.int64) + counts <= M)
    xs, ds, dl = [], [], []
    for off, cnt, k in zip(r[:, 1], counts, kept):
        if k:
            xs.append(xyzs[off:off + cnt]); ds.append(dirs[off:off + cnt]); dl.append(deltas[off:off + cnt])
    cat = lambda a, w: np.concatenate(a, 0) if a else np.zeros((0, w), np.float32)
    return r[:, 0].astype(np.int32), counts, kept, cat(xs, 3), cat(ds, 3), cat(dl, 2)
