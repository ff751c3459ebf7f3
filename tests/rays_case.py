"""Cases of the ray-generation parity test (shared by tests/golden/make_rays_golden.py and the tests)."""
import numpy as np

# name -> (H, W, yaw_deg, pitch_deg); one non-square frame pins the (row, column) conventions of get_rays / get_bg_coords
CASES = {"sq48": (48, 48, 6.0, 2.0), "rect24x40": (24, 40, -9.0, 3.5), "sq64": (64, 64, 0.0, 0.0)}


def case_pose(yaw, pitch):
    from radnerf_b200 import synthetic as syn
    return syn.orbit_pose(yaw_deg=yaw, pitch_deg=pitch)


def case_intrinsics(H, W):
    from radnerf_b200 import synthetic as syn
    fx, fy, cx, cy = syn.intrinsics_for(H, W)
    return np.array([fx, fy * 1.01, cx + 0.25, cy - 0.5], dtype=np.float32)   # distinct fx/fy, off-centre principal point
