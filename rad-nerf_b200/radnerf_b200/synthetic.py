"""Synthetic, seeded inputs for tests and benchmarks (there are no datasets or checkpoints offline).

Input GENERATORS only -- numpy/torch-CPU code that builds the tensors the hot path consumes: camera rays from an
orbit pose (the formula of the reference's get_rays, nerf/utils.py:248-333), background coordinates
(get_bg_coords, nerf/utils.py:239-245), audio-feature windows (get_audio_features att=2, nerf/utils.py:56-72), an
analytic "head" occupancy grid in the reference's Morton layout (nerf/renderer.py:112-113), and a torso mask grid.
Nothing here is on the measured path.
"""
import math

import numpy as np


# --------------------------------------------------------------------------------------------- camera
def orbit_pose(yaw_deg=0.0, radius=3.35, pitch_deg=0.0):
    """cam2world [4,4] float32: camera on a circle of `radius` around the origin, looking at the origin from +y
    (the recorded obama pose has translation ~(0.07, 3.38, -0.23), nerf/utils.py:236)."""
    yaw, pitch = math.radians(yaw_deg), math.radians(pitch_deg)
    # camera position
    pos = np.array([radius * math.sin(yaw) * math.cos(pitch), radius * math.cos(yaw) * math.cos(pitch),
                    radius * math.sin(pitch)], dtype=np.float64)
    fwd = -pos / np.linalg.norm(pos)  # camera +z looks at the origin
    up = np.array([0.0, 0.0, 1.0])
    right = np.cross(up, fwd)
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    c2w = np.eye(4, dtype=np.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, down, fwd, pos
    return c2w.astype(np.float32)


def intrinsics_for(H, W, fovy_deg=21.24):
    """(fx, fy, cx, cy): focal from the vertical field of view (main.py:71-72 default fovy)."""
    f = H / (2 * math.tan(math.radians(fovy_deg) / 2))
    return np.array([f, f, W / 2, H / 2], dtype=np.float32)


def get_rays(pose, intrinsics, H, W):
    """Full-image rays, row-major pixel order.  Returns rays_o [H*W,3], rays_d [H*W,3] float32.
    Same arithmetic as the reference (pixel centre +0.5, normalise, rotate by the pose)."""
    fx, fy, cx, cy = [np.float32(v) for v in intrinsics]
    i = (np.arange(W, dtype=np.float32) + np.float32(0.5))[None, :].repeat(H, 0).reshape(-1)
    j = (np.arange(H, dtype=np.float32) + np.float32(0.5))[:, None].repeat(W, 1).reshape(-1)
    xs = (i - cx) / fx
    ys = (j - cy) / fy
    zs = np.ones_like(xs)
    d = np.stack([xs, ys, zs], -1).astype(np.float32)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True).astype(np.float32)
    R = pose[:3, :3].astype(np.float32)
    rays_d = (d @ R.T).astype(np.float32)
    rays_o = np.broadcast_to(pose[:3, 3].astype(np.float32), rays_d.shape).copy()
    return rays_o, rays_d


def get_bg_coords(H, W):
    """[H*W, 2] float32 in [-1,1]; note the reference's meshgrid is indexed (H, W) -> (x = row, y = col)."""
    X = np.arange(H, dtype=np.float32) / np.float32(H - 1) * 2 - 1
    Y = np.arange(W, dtype=np.float32) / np.float32(W - 1) * 2 - 1
    xs, ys = np.meshgrid(X, Y, indexing="ij")
    return np.stack([xs.reshape(-1), ys.reshape(-1)], -1).astype(np.float32)


# --------------------------------------------------------------------------------------------- occupancy
def _spread3(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3D_np(x, y, z):
    return _spread3(x) | (_spread3(y) << np.uint32(1)) | (_spread3(z) << np.uint32(2))


def head_density_grid(H=128, semi_axes=(0.28, 0.22, 0.30), inside=20.0, neck=True, bound=1.0):
    """density_grid [1, H^3] float32, Morton-indexed: `inside` within an ellipsoid (+ a neck cylinder), 0 elsewhere."""
    c = (-1.0 + (2.0 * np.arange(H) + 1.0) / H) * bound
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    a, b, cc = semi_axes
    occ = (X / a) ** 2 + (Y / b) ** 2 + (Z / cc) ** 2 <= 1.0
    if neck:
        occ |= ((X ** 2 + Y ** 2) <= 0.11 ** 2) & (Z > 0.2) & (Z < 0.55)
    ii, jj, kk = np.meshgrid(np.arange(H), np.arange(H), np.arange(H), indexing="ij")
    idx = morton3D_np(ii.reshape(-1), jj.reshape(-1), kk.reshape(-1)).astype(np.int64)
    grid = np.zeros((1, H ** 3), np.float32)
    grid[0, idx] = np.where(occ.reshape(-1), np.float32(inside), np.float32(0.0))
    return grid


def packbits_np(grid, thresh):
    """numpy restatement of the bitfield layout (bit i of byte n = cell 8n+i), for building inputs."""
    bits = (grid.reshape(-1) > np.float32(thresh)).astype(np.uint8)
    return np.packbits(bits, bitorder="little")


def torso_density_grid(H=128, value=0.5):
    """density_grid_torso [H*H] float32: a trapezoid in the lower third of the image (~32% of the pixels).
    Indexed [y * H + x] as the reference fills it (nerf/renderer.py:472) and sampled by grid_sample on bg_coords."""
    g = np.zeros((H, H), np.float32)
    for r in range(H):
        v = r / (H - 1)  # 0 top .. 1 bottom ; rows map to bg_coords x (image rows) via grid_sample's y
        if v > 0.58:
            half = 0.18 + 0.5 * (v - 0.58) / 0.42
            lo, hi = int((0.5 - half) * H), int((0.5 + half) * H)
            g[max(lo, 0):min(hi, H), r] = value
    return g.reshape(-1)


# --------------------------------------------------------------------------------------------- audio
def audio_feature_bank(n=600, dim=44, win=16, seed=0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, dim, win)) * 3.0).astype(np.float32)


def audio_window(bank, index, att=2):
    """get_audio_features (nerf/utils.py:42-74) for att modes 0 and 2: 8-frame window centred on `index`, zero padded."""
    if att == 0:
        return bank[[index]]
    left, right = index - 4, index + 4
    pad_l = pad_r = 0
    if left < 0:
        pad_l, left = -left, 0
    if right > bank.shape[0]:
        pad_r, right = right - bank.shape[0], bank.shape[0]
    a = bank[left:right]
    if pad_l:
        a = np.concatenate([np.zeros((pad_l,) + a.shape[1:], a.dtype), a], 0)
    if pad_r:
        a = np.concatenate([a, np.zeros((pad_r,) + a.shape[1:], a.dtype)], 0)
    return a


# --------------------------------------------------------------------------------------------- training batches
def training_batch(H, W, n_rays, frame_index=0, seed=0, att=2, audio_dim=44):
    """One synthetic training batch in the shape NeRFDataset.collate produces for training (nerf/provider.py:250-290,
    SURVEY 8(d) "Training"): n_rays random pixels of an H x W frame seen from the orbit camera, target colours ~ U(0,1),
    face mask = centre box, white background, audio window / eye / pose of frame `frame_index`.  numpy, host side."""
    from .posemath import convert_poses
    import torch
    rng = np.random.default_rng(1000 * seed + frame_index)
    pose = orbit_pose(yaw_deg=10.0 * np.sin(2 * np.pi * frame_index / 64), pitch_deg=2.0)
    ro, rd = get_rays(pose, intrinsics_for(H, W), H, W)
    pix = rng.choice(H * W, size=n_rays, replace=n_rays > H * W)
    ys, xs = pix // W, pix % W
    face = (np.abs(ys - H / 2) < H / 4) & (np.abs(xs - W / 2) < W / 4)
    bank = audio_feature_bank(600, audio_dim, 16, seed=0)
    return dict(rays_o=ro[pix][None], rays_d=rd[pix][None], bg_coords=get_bg_coords(H, W)[pix][None],
                poses=convert_poses(torch.from_numpy(pose)[None]).numpy(), auds=audio_window(bank, 8 + frame_index, att),
                eye=np.array([[0.25]], np.float32), index=[frame_index], rgb=rng.random((1, n_rays, 3), dtype=np.float32),
                face_mask=face[None], bg_color=np.ones((n_rays, 3), np.float32))


def batch_to(batch, device):
    import torch
    out = {}
    for k, v in batch.items():
        out[k] = torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v
    return out
