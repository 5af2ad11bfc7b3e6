"""Shared by make_golden_process_slam.py (REFERENCE NerfFusion.process_slam) and the tests: the SLAM packet fed to the
fusion side (keys of visual_frontend.py:1364-1382), seeded."""
import types

import numpy as np
import torch

N, H, W = 3, 16, 24
MASK_TYPES = ("ours", "raw", "ours_w_thresh", "no_depth")


class _Model:
    def numpy(self):
        return np.array([21.5, 22.25, 11.5, 7.75])


class _Res:
    def numpy(self):
        return np.array([W, H])


def make_packet(seed=5):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    poses = np.concatenate([rng.normal(size=(N, 3)), q], 1).astype(np.float32)          # cam0_T_world (t, q xyzw)
    calib = types.SimpleNamespace(aabb=np.array([[-2, -2, -2], [2, 2, 2]]), depth_scale=1.0 / 6553.5,
                                  camera_model=_Model(), resolution=_Res())
    imgs = rng.integers(0, 256, (N, 3, H, W)).astype(np.uint8)
    imgs[0, :, 0, :8] = np.arange(8, dtype=np.uint8)[None] * 3                           # values around the sRGB knee (0.04045 * 255 = 10.3)
    slam = {"viz_idx": torch.tensor([2, 5, 7]), "cam0_poses": torch.from_numpy(poses),
            "cam0_images": torch.from_numpy(imgs),
            "cam0_idepths_up": torch.from_numpy(rng.uniform(0.2, 2.0, (N, H, W)).astype(np.float32)),
            "cam0_depths_cov_up": torch.from_numpy(rng.uniform(0.01, 4.0, (N, H, W)).astype(np.float32)),
            "gt_depths": torch.from_numpy(rng.uniform(1000, 20000, (N, 1, H, W)).astype(np.float32)),
            "calibs": [calib], "is_last_frame": False}
    return slam


def make_data_packet(seed=9):
    """a dataset packet (datasets/nerf_dataset.py:155-162) as the GT-fitting path receives it (--fusion=nerf without --slam)"""
    rng = np.random.default_rng(seed)
    n = 2
    q = rng.normal(size=(n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    w2c = np.zeros((n, 4, 4)); w2c[:, 3, 3] = 1
    for k in range(n):
        x, y, z, w = q[k]
        w2c[k, :3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        w2c[k, :3, 3] = rng.normal(size=3)
    calib = types.SimpleNamespace(aabb=np.array([[-2.0, -1.0, -2.0], [2.0, 3.0, 2.0]]), depth_scale=1.0 / 6553.5,
                                  camera_model=_Model(), resolution=_Res())
    return {"k": np.array([4, 6]), "t_cams": np.array([4, 6]), "poses": w2c,
            "images": rng.integers(0, 256, (n, H, W, 4)).astype(np.uint8),
            "depths": rng.integers(0, 40000, (n, H, W, 1)).astype(np.int32),
            "calibs": np.array([calib, calib]), "is_last_frame": False}
