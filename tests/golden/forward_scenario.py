"""Per-frame control flow of the live front end (RaftVisualFrontend.forward, visual_frontend.py:240-365, with
initialize_buffers :160-238 and get_viz_out :1337-1391) driven identically for the reference's methods executed verbatim
(make_golden_forward.py) and for this repo's class on a CPU shim (tests/test_cpu_frontend_forward.py).
Stand-ins on both sides: the encoders (features that encode the frame number), the motion filter decision, and
__initialize / __update / rm_keyframe / terminate (their own behaviour is pinned by live_frontend_scenario.py)."""
import types

import numpy as np

H, W = 32, 48          # image size (1/8: 4 x 6)
CASES = [dict(seed=51, n_frames=40, buffer=64, last_has_motion=True),
         dict(seed=52, n_frames=36, buffer=64, last_has_motion=False),
         dict(seed=53, n_frames=60, buffer=12, last_has_motion=True)]      # reaches the buffer-full stop


class _Model:
    def numpy(self):
        return np.array([24.0, 25.0, 23.5, 15.5])


def plan(seed, n_frames, last_has_motion):
    rng = np.random.default_rng(seed)
    motion = rng.random(n_frames) < 0.7
    motion[0] = True
    motion[-1] = last_has_motion
    accept = rng.random(n_frames) < 0.6
    return motion, accept


def packet(k, n_frames):
    rng = np.random.default_rng(1000 + k)
    img = rng.integers(0, 256, (1, H, W, 4)).astype(np.uint8)
    img[0, 0, 0, 0] = k                                         # the stand-in encoders read the frame number here
    pose = np.eye(4); pose[:3, 3] = [0.1 * k, -0.05 * k, 0.02 * k]
    calib = types.SimpleNamespace(camera_model=_Model(), depth_scale=1.0 / 6553.5, aabb=[[-2, -2, -2], [2, 2, 2]])
    return {"k": np.arange(k, k + 1), "t_cams": np.array([100 + k]), "poses": np.array([pose]), "images": img,
            "depths": rng.integers(0, 30000, (1, H, W, 1)).astype(np.int32), "calibs": np.array([calib]),
            "is_last_frame": k >= n_frames - 1}


def summarize_viz(v):
    if v is None:
        return None
    out = {"keys": sorted(v.keys()), "is_last_frame": bool(v["is_last_frame"])}
    if "viz_idx" in v:
        out.update({"viz_idx": [int(i) for i in v["viz_idx"].tolist()], "kf_idx": int(v["kf_idx"]),
                    "kf_idx_to_f_idx": {int(a): int(b) for a, b in v["kf_idx_to_f_idx"].items()},
                    "shapes": {k: list(v[k].shape) for k in ("cam0_poses", "cam0_images", "cam0_idepths_up", "cam0_depths_cov_up", "gt_depths",
                                                             "cam0_intrinsics", "world_T_body_cov")},
                    "images_sum": int(v["cam0_images"].long().sum()), "gt_depth_sum": float(v["gt_depths"].double().sum()),
                    "intr": [round(float(x), 5) for x in v["cam0_intrinsics"].reshape(-1).tolist()],
                    "poses_sum": round(float(v["cam0_poses"].double().sum()), 5), "tstamps": None})
    return out
