"""Head-pose helpers for the torso branch: the 6-vector (XYZ Euler angles, translation) the reference feeds to its pose
encoder (convert_poses, nerf/utils.py:229-237, built on pytorch3d's matrix_to_euler_angles for convention 'XYZ').

For R = Rx(a) Ry(b) Rz(c):  R[0,2] = sin b,  R[1,2] = -sin a cos b,  R[2,2] = cos a cos b,  R[0,1] = -cos b sin c,
R[0,0] = cos b cos c, hence the closed form below (identical values to the generic convention code for 'XYZ')."""
import torch


def matrix_to_euler_xyz(R):
    a = torch.atan2(-R[..., 1, 2], R[..., 2, 2])
    b = torch.asin(R[..., 0, 2])
    c = torch.atan2(-R[..., 0, 1], R[..., 0, 0])
    return torch.stack((a, b, c), -1)


def euler_xyz_to_matrix(e):
    a, b, c = e.unbind(-1)
    ca, sa, cb, sb, cc, sc = a.cos(), a.sin(), b.cos(), b.sin(), c.cos(), c.sin()
    one, zero = torch.ones_like(a), torch.zeros_like(a)
    Rx = torch.stack((one, zero, zero, zero, ca, -sa, zero, sa, ca), -1).reshape(a.shape + (3, 3))
    Ry = torch.stack((cb, zero, sb, zero, one, zero, -sb, zero, cb), -1).reshape(a.shape + (3, 3))
    Rz = torch.stack((cc, -sc, zero, sc, cc, zero, zero, zero, one), -1).reshape(a.shape + (3, 3))
    return Rx @ Ry @ Rz


@torch.autocast(device_type="cuda", enabled=False)
def convert_poses(poses):
    """poses [B,4,4] cam2world -> [B,6] = (euler XYZ of the rotation, translation), fp32"""
    poses = torch.as_tensor(poses)
    out = torch.empty(poses.shape[0], 6, dtype=torch.float32, device=poses.device)
    out[:, :3] = matrix_to_euler_xyz(poses[:, :3, :3].float())
    out[:, 3:] = poses[:, :3, 3]
    return out
