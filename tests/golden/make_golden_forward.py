"""Record the per-frame control flow of the reference's live front end by executing its own `forward`,
`initialize_buffers` and `get_viz_out` verbatim (compiled from the source text via `ast`, see
make_golden_live_frontend.py) — build container only.

  python tests/golden/make_golden_forward.py        ->  tests/golden/ref_forward_traces.json.gz"""
import ast
import gzip
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = ["forward", "initialize_buffers", "get_viz_out"]


def reference_class():
    src = open(os.path.join(REF, "slam", "visual_frontends", "visual_frontend.py")).read()
    cls = [n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "RaftVisualFrontend"][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert sorted(n.name for n in keep) == sorted(METHODS)
    new = ast.Module(body=[ast.ClassDef(name="RaftVisualFrontend", bases=[], keywords=[], body=keep, decorator_list=[])], type_ignores=[])
    ast.fix_missing_locations(new)

    def coords_grid(ht, wd, device=None):
        y, x = torch.meshgrid(torch.arange(ht).float(), torch.arange(wd).float(), indexing="ij")
        return torch.stack([x, y], dim=-1)
    Empty = type("Empty", (), {})
    ns = {"torch": torch, "np": np, "ic": lambda *a, **k: None, "cv2": None, "Values": Empty, "NonlinearFactorGraph": Empty,
          "pops": types.SimpleNamespace(coords_grid=coords_grid),
          "gtsam_pose_to_torch": lambda pose, device=None, dtype=None: torch.as_tensor(pose, dtype=dtype)}
    exec(compile(new, "visual_frontend.py (reference, selected methods)", "exec"), ns)
    return ns["RaftVisualFrontend"]


def run_case(Ref, sc, case):
    motion, accept = sc.plan(case["seed"], case["n_frames"], case["last_has_motion"])
    log = []

    class Shim(Ref):
        def __init__(self):
            self.device, self.buffer, self.dsf, self.stereo, self.viz = "cpu", case["buffer"], 8, False, False
            self.args = types.SimpleNamespace(multi_gpu=False)
            self.kf_idx, self.last_kf_idx, self.last_k = 0, 0, None
            self.kf_idx_to_f_idx, self.f_idx_to_kf_idx = {}, {}
            self.is_initialized, self.keyframe_warmup, self.stop = False, 8, False
            self.cam0_t0_T_world = torch.tensor([0.1, 0.2, 0.3, 0, 0, 0, 1.0])
            self.world_T_body_t0 = [-0.1, -0.2, -0.3, 0, 0, 0, 1.0]
            self.g_prior_cov = torch.block_diag(0.01 ** 2 * torch.eye(3), 0.01 ** 2 * torch.eye(3))
            self.idepth_prior_cov = torch.tensor(0.1) ** 2

        def _normalize_imgs(self, images, droid_normalization=True):
            return images[:, :, :3, ...].float()

        def _k(self, imgs):
            return int(imgs[0, 0, 0, 0, 0])

        def _RaftVisualFrontend__feature_encoder(self, imgs):
            return torch.full((1, 128, self.ht, self.wd), float(self._k(imgs)), dtype=torch.half)

        def _RaftVisualFrontend__context_encoder(self, imgs):
            k = float(self._k(imgs))
            return torch.full((1, 128, self.ht, self.wd), k + 0.25, dtype=torch.half), torch.full((1, 128, self.ht, self.wd), k + 0.5, dtype=torch.half)

        def has_enough_motion(self, feats):
            return bool(motion[int(feats[0, 0, 0, 0])])

        def _RaftVisualFrontend__initialize(self):
            log.append(["initialize", self.kf_idx]); self.is_initialized = True; self.viz_idx[:self.kf_idx + 1] = True

        def _RaftVisualFrontend__update(self):
            ok = bool(accept[self.kf_idx_to_f_idx[self.kf_idx]])
            log.append(["update", self.kf_idx, ok]); self.viz_idx[max(self.kf_idx - 2, 0):self.kf_idx + 1] = True
            return ok

        def rm_keyframe(self, k):
            log.append(["rm_keyframe", k])

        def terminate(self, stream=None):
            log.append(["terminate", self.kf_idx]); self.stop = True

    fe = Shim()
    trace = []
    for k in range(case["n_frames"]):
        x0, factors, viz = fe.forward(sc.packet(k, case["n_frames"]))
        assert x0 is not None and factors is not None
        d = {"k": k, "kf_idx": int(fe.kf_idx), "last_k": None if fe.last_k is None else int(fe.last_k), "last_kf_idx": int(fe.last_kf_idx),
             "is_initialized": bool(fe.is_initialized), "stop": bool(fe.stop),
             "kf2f": {int(a): int(b) for a, b in fe.kf_idx_to_f_idx.items()}, "f2kf": {int(a): int(b) for a, b in fe.f_idx_to_kf_idx.items()},
             "viz": sc.summarize_viz(viz), "log": list(log),
             "feat_ids": [int(v) for v in fe.features_imgs[:, 0, 0, 0, 0].tolist()],
             "ctx_ids": [round(float(v), 2) for v in fe.contexts_imgs[:, 0, 0, 0, 0].tolist()],
             "tstamps": [float(v) for v in fe.cam0_timestamps.tolist()]}
        log.clear()
        trace.append(d)
        if fe.stop:
            break
    init = {"idepths_cov": float(fe.cam0_idepths_cov[-1, 0, 0]), "depths_cov": float(fe.cam0_depths_cov[-1, 0, 0]),
            "idepths": float(fe.cam0_idepths[-1, 0, 0]), "idepths_up": float(fe.cam0_idepths_up[-1, 0, 0]),
            "depths_cov_up": float(fe.cam0_depths_cov_up[-1, 0, 0]), "T_world": fe.cam0_T_world[-1].tolist(),
            "wTb": fe.world_T_body[-1].tolist(), "wTb_cov_diag": torch.diagonal(fe.world_T_body_cov[-1]).tolist(),
            "damping": float(fe.damping[-1, 0, 0]), "coords0_last": fe.coords0[-1, -1].tolist(),
            "shapes": {n: list(getattr(fe, n).shape) for n in ("cam0_images", "gt_depths", "cam0_idepths", "cam0_idepths_up", "features_imgs",
                                                                "contexts_imgs", "cst_contexts_imgs", "world_T_body_cov", "cam0_intrinsics")}}
    return {"case": case, "trace": trace, "init": init}


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, HERE)
    import forward_scenario as sc
    Ref = reference_class()
    out = [run_case(Ref, sc, c) for c in sc.CASES]
    for o in out:
        t = o["trace"]
        print(o["case"], "frames run", len(t), "keyframes", t[-1]["kf_idx"], "stop", t[-1]["stop"],
              "viz packets", sum(1 for x in t if x["viz"] and "viz_idx" in x["viz"]))
    path = os.path.join(HERE, "ref_forward_traces.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as gz, __import__("io").TextIOWrapper(gz) as f:   # mtime=0: reproducible bytes
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
