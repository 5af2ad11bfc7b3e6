"""Shared by make_golden_dataset.py (reads with the REFERENCE's NeRFDataset) and tests/test_cpu_datasets.py (reads with
nerf_slam_b200.datasets.NeRFDataset): which datasets are written, how they are loaded, how a packet is digested."""
import hashlib
import types

import numpy as np

# (name, width, height, frames written, loader arguments)
CASES = [
    ("small", 64, 48, 6, dict(initial_k=0, final_k=None, img_stride=1)),
    ("small_strided", 64, 48, 6, dict(initial_k=1, final_k=5, img_stride=2)),
    ("large_resized", 800, 704, 2, dict(initial_k=0, final_k=None, img_stride=1)),      # > 640*640 pixels: down-scaled on load
]


def loader_args(dataset_dir, initial_k, final_k, img_stride):
    return types.SimpleNamespace(dataset_dir=dataset_dir, initial_k=initial_k, final_k=final_k, img_stride=img_stride, stereo=False)


def write_case(name, w, h, n, out_dir):
    from nerf_slam_b200.datasets import write_transforms_dataset
    from nerf_slam_b200.synthetic import SyntheticRoom
    write_transforms_dataset(SyntheticRoom(w, h, n, seed=3, step=0.05), out_dir)


def _grid_means(a, g=8):
    a = np.asarray(a, np.float64)
    H, W = a.shape[:2]
    return [[float(a[i * H // g:(i + 1) * H // g, j * W // g:(j + 1) * W // g].mean()) for j in range(g)] for i in range(g)]


def digest_packet(p, exact):
    cal = p["calibs"][0]
    d = {"k": np.asarray(p["k"]).tolist(), "t_cams": np.asarray(p["t_cams"]).tolist(),
         "poses": np.asarray(p["poses"], np.float64).tolist(), "is_last_frame": bool(p["is_last_frame"]),
         "image_shape": list(p["images"].shape), "image_dtype": str(p["images"].dtype),
         "depth_shape": list(p["depths"].shape), "depth_dtype": str(p["depths"].dtype),
         "intrinsics": np.asarray(cal.camera_model.numpy(), np.float64).tolist(),
         "resolution": [int(cal.resolution.width), int(cal.resolution.height)],
         "aabb": np.asarray(cal.aabb, np.float64).tolist(), "depth_scale": float(cal.depth_scale),
         "rate_hz": float(cal.rate_hz)}
    if exact:                                   # lossless path: the decoded arrays must be identical
        d["image_sha1"] = hashlib.sha1(np.ascontiguousarray(p["images"]).tobytes()).hexdigest()
        d["depth_sha1"] = hashlib.sha1(np.ascontiguousarray(p["depths"]).tobytes()).hexdigest()
    else:                                       # resized path: block means (cv2.resize may differ by an LSB across CPUs)
        d["image_means"] = _grid_means(p["images"][0])
        d["depth_means"] = _grid_means(p["depths"][0])
    return d
