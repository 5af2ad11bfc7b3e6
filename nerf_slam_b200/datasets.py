"""Input side of the hot path: the reference's on-disk format and its packet schema.

  NeRFDataset(args, device)        datasets/nerf_dataset.py:14-171 — `transforms.json` + PNG directory -> packets
      {"k", "t_cams", "poses" (w2c 4x4), "images" u8 [1,H,W,4] RGBA, "depths" int32 [1,H,W,1], "calibs", "is_last_frame"}
  write_transforms_dataset(room, dir, n)   the procedural stream (synthetic.SyntheticRoom) in that format, so that
      the reference's own CLI (`examples/slam_demo.py --dataset_name=nerf --dataset_dir=...`) and this repo read the
      same frames.

Parity: tests/test_cpu_datasets.py compares this reader, frame for frame, with digests recorded from the REFERENCE's
`NeRFDataset` reading the same files (tests/golden/make_golden_dataset.py), including its down-scaling rule for
images above 640 x 640 pixels (nerf_dataset.py:45-62) and the nerf -> ngp pose convention (utils/utils.py:104-116).
"""
import json
import os
import types

import numpy as np

from .synthetic import CameraCalibration, PinholeCameraModel, Resolution


def nerf_matrix_to_ngp(nerf_matrix, scale=1.0, offset=0.5):
    """utils/utils.py:104-116: flip y/z axes, scale + offset the position, cycle the rows xyz <- yzx"""
    r = np.array(nerf_matrix, dtype=np.float64, copy=True)
    r[:3, 1] *= -1
    r[:3, 2] *= -1
    r[:3, 3] = r[:3, 3] * scale + offset
    return np.concatenate([r[1:2], r[2:3], r[0:1], r[3:4]], 0)


def ngp_matrix_to_nerf(ngp_matrix, scale=1.0, offset=0.5):
    """exact inverse of nerf_matrix_to_ngp (the reference's own inverse, utils/utils.py:119-131, overwrites a row it
    still needs)"""
    g = np.asarray(ngp_matrix, dtype=np.float64)
    r = np.concatenate([g[2:3], g[0:1], g[1:2], g[3:4]], 0).copy()
    r[:3, 3] = (r[:3, 3] - offset) / scale
    r[:3, 1] *= -1
    r[:3, 2] *= -1
    return r


class RadTanDistortionModel:
    """datasets/dataset.py:103-113"""

    def __init__(self, k1, k2, p1, p2):
        self.model = "RadTan"
        self.k1, self.k2, self.p1, self.p2 = k1, k2, p1, p2

    def get_distortion_as_vector(self):
        return np.array([self.k1, self.k2, self.p1, self.p2])


def write_transforms_dataset(room, out_dir, n_frames=None):
    """room: synthetic.SyntheticRoom -> out_dir/transforms.json, images/frame%06d.png (RGBA 8 bit),
    depths/frame%06d.png (16 bit, metric depth / integer_depth_scale)"""
    import cv2
    n = len(room) if n_frames is None else n_frames
    os.makedirs(os.path.join(out_dir, "images"), exist_ok=True)
    os.makedirs(os.path.join(out_dir, "depths"), exist_ok=True)
    cm = room.calib.camera_model
    meta = {"w": room.W, "h": room.H, "fl_x": cm.fx, "fl_y": cm.fy, "cx": cm.cx, "cy": cm.cy,
            "aabb": np.asarray(room.calib.aabb).tolist(), "integer_depth_scale": room.calib.depth_scale, "frames": []}
    for k in range(n):
        rgba, d16, w2c = room.render(k)
        name = f"frame{k:06d}.png"
        cv2.imwrite(os.path.join(out_dir, "images", name), cv2.cvtColor(rgba, cv2.COLOR_RGBA2BGRA))
        cv2.imwrite(os.path.join(out_dir, "depths", name), d16[..., 0].astype(np.uint16))
        meta["frames"].append({"file_path": f"images/{name}", "depth_path": f"depths/{name}",
                               "transform_matrix": ngp_matrix_to_nerf(np.linalg.inv(w2c)).tolist()})
    with open(os.path.join(out_dir, "transforms.json"), "w") as f:
        json.dump(meta, f)
    return out_dir


class NeRFDataset:
    """datasets/nerf_dataset.py:14-171 (+ the base class fields of datasets/dataset.py:9-26).
    args: dataset_dir, initial_k, final_k, img_stride, stereo (argparse namespace of examples/slam_demo.py)"""

    def __init__(self, args, device="cpu"):
        self.name, self.args, self.device = "Nerf", args, device
        self.dataset_dir = args.dataset_dir
        self.initial_k, self.final_k, self.img_stride = args.initial_k, args.final_k, args.img_stride
        self.stereo = getattr(args, "stereo", False)
        self.viz = False
        self.data_packets = None
        self.parse_metadata()

    def get_cam_calib(self):
        """:21-36"""
        j = self.json
        depth_scale = j["integer_depth_scale"] if "integer_depth_scale" in j else 1.0
        return CameraCalibration(np.eye(4, 4), PinholeCameraModel(j["fl_x"], j["fl_y"], j["cx"], j["cy"]),
                                 RadTanDistortionModel(0, 0, 0, 0), 10.0, Resolution(j["w"], j["h"]), j["aabb"], depth_scale)

    def parse_metadata(self):
        """:38-96"""
        with open(os.path.join(self.dataset_dir, "transforms.json"), "r") as f:
            self.json = json.load(f)
        self.calib = self.get_cam_calib()
        self.resize_images = self.calib.resolution.total() > 640 * 640
        if self.resize_images:
            # equal-area down-scaling to ~341 x 640 pixels, both sides multiples of 8 (:45-62)
            h0, w0 = self.calib.resolution.height, self.calib.resolution.width
            total = 341 * 640
            self.h1 = int(h0 * np.sqrt(total / (h0 * w0)))
            self.w1 = int(w0 * np.sqrt(total / (h0 * w0)))
            self.h1 -= self.h1 % 8
            self.w1 -= self.w1 % 8
            self.calib.camera_model.scale_intrinsics(self.w1 / w0, self.h1 / h0)
            self.calib.resolution = Resolution(self.w1, self.h1)
        frames = self.json["frames"][self.initial_k:self.final_k:self.img_stride]
        self.image_paths, self.depth_paths, self.w2c = [], [], []
        for i, frame in enumerate(frames):
            c2w = nerf_matrix_to_ngp(np.array(frame["transform_matrix"]))
            fp = frame["file_path"]
            image_path = os.path.join(self.dataset_dir, fp if fp.endswith((".png", ".jpg")) else fp + ".png")
            depth_path = os.path.join(self.dataset_dir, frame["depth_path"]) if "depth_path" in frame else None
            self.image_paths.append([i, image_path])
            self.depth_paths.append(depth_path)
            self.w2c.append(np.linalg.inv(c2w))
        # (the reference calls sorted() on the paths and drops the result, :86-91: json order is the stream order)
        self.args.world_T_imu_t0 = self.w2c[0]

    def _get_data_packet(self, k0, k1=None):
        """:98-160"""
        import cv2
        k1 = k0 + 1 if k1 is None else k1
        assert k1 >= k0
        W, H = self.calib.resolution.width, self.calib.resolution.height
        ts, poses, images, depths, calibs = [], [], [], [], []
        for k in np.arange(k0, k1):
            i, image_path = self.image_paths[k]
            depth_path = self.depth_paths[i]
            image = cv2.cvtColor(cv2.imread(image_path), cv2.COLOR_BGRA2RGBA)        # imread drops alpha; RGBA with A = 255
            if depth_path:
                depth = cv2.imread(depth_path, cv2.IMREAD_UNCHANGED)[..., None]
            else:
                depth = (-1 * np.ones_like(image[:, :, 0])).astype(np.uint16)        # invalid depth (2-D, as the reference)
            if self.resize_images:
                image = cv2.resize(image, (self.w1, self.h1))
                depth = cv2.resize(depth, (self.w1, self.h1))[:, :, np.newaxis]
            assert image.shape[:2] == (H, W) and image.shape[2] in (3, 4) and image.dtype == np.uint8
            assert depth.shape == (H, W, 1) and depth.dtype == np.uint16
            ts.append(i); poses.append(self.w2c[i]); images.append(image); depths.append(depth.astype(np.int32))
            calibs.append(self.calib)
        return {"k": np.arange(k0, k1), "t_cams": np.array(ts), "poses": np.array(poses), "images": np.array(images),
                "depths": np.array(depths), "calibs": np.array(calibs), "is_last_frame": (i >= len(self) - 1)}

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, k):
        return self._get_data_packet(k) if self.data_packets is None else self.data_packets[k]

    def stream(self):
        for k in range(len(self)):
            yield self[k]


def dataset_args(dataset_dir, initial_k=0, final_k=None, img_stride=1, stereo=False):
    """the fields of examples/slam_demo.py's argparse namespace that the dataset reads"""
    return types.SimpleNamespace(dataset_dir=dataset_dir, initial_k=initial_k, final_k=final_k, img_stride=img_stride, stereo=stereo)
