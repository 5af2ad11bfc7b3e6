"""Golden for TsdfFusion.update_history / get_history_packet (fusion/tsdf_fusion.py:486-543), executed VERBATIM (methods cut
out of the reference file with `ast`; they are plain torch / Python) on a scripted packet sequence.

    python tests/golden/make_golden_tsdf_history.py      (needs /root/reference; writes tests/golden/ref_tsdf_history.json)"""
import ast
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/fusion/tsdf_fusion.py"


def packets(H=6, W=8):
    """three SLAM packets with overlapping dirty windows + the last-frame packet; tensors seeded per (keyframe, version)"""
    out = []
    k2f = {}
    for kf_idx, (ids, frames) in enumerate([((0, 1, 2), (0, 3, 5)), ((1, 2, 3), (3, 5, 9)), ((3, 4), (9, 12))]):
        k2f.update(dict(zip(ids, frames)))
        g = torch.Generator().manual_seed(100 + kf_idx)
        n = len(ids)
        out.append({"is_last_frame": False, "kf_idx": kf_idx + 2, "viz_idx": torch.tensor(ids),
                    "cam0_poses": torch.randn(n, 7, generator=g), "cam0_depths_cov_up": torch.rand(n, H, W, generator=g),
                    "cam0_idepths_up": torch.rand(n, H, W, generator=g) + 0.5,
                    "cam0_images": torch.randint(0, 255, (n, 3, H, W), dtype=torch.uint8, generator=g),
                    "cam0_intrinsics": torch.rand(n, 4, generator=g), "gt_depths": torch.rand(n, 1, H, W, generator=g),      # [n,1,H,W] as the front end sends it (visual_frontend.py:181,1343)
                    "calibs": [types.SimpleNamespace(depth_scale=0.5)], "kf_idx_to_f_idx": dict(k2f)})
    out.append({"is_last_frame": True})
    return out


def summarize(fusion, returns):
    hist = {str(k): {"kf_idx": int(h["kf_idx"]), "viz_idx": int(h["viz_idx"]),
                     "sums": [round(float(torch.as_tensor(h[key]).double().sum()), 5) for key in
                              ("cam0_poses", "cam0_depths_cov_up", "cam0_idepths_up", "cam0_images", "cam0_intrinsics")]}
            for k, h in fusion.history.items()}
    pk = fusion.get_history_packet()
    packet = {k: {"shape": list(pk[k].shape), "sum": round(float(pk[k].double().sum()), 4)}
              for k in ("viz_idx", "cam0_poses", "cam0_depths_cov_up", "cam0_idepths_up", "cam0_images", "cam0_intrinsics")}
    packet["n_calibs"] = len(pk["calibs"])
    packet["gt_depths_sum"] = round(float(pk["gt_depths"].double().sum()), 4)
    return {"returns": returns, "order": [str(k) for k in fusion.history], "history": hist, "packet": packet}


def main():
    src = open(REF).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "TsdfFusion")
    body = "\n".join(ast.get_source_segment(src, f) for f in cls.body
                     if isinstance(f, ast.FunctionDef) and f.name in ("update_history", "get_history_packet"))
    code = "class TsdfFusion:\n" + "\n".join("    " + l if not l.startswith("    ") else l for l in body.split("\n"))
    ns = {"torch": torch, "np": np}
    import warnings
    warnings.filterwarnings("ignore")
    exec(compile(code, REF, "exec"), ns)
    ref = ns["TsdfFusion"].__new__(ns["TsdfFusion"])
    ref.history = {}
    returns = [bool(ref.update_history(p)) for p in packets()]
    with open(os.path.join(HERE, "ref_tsdf_history.json"), "w") as f:
        json.dump(summarize(ref, returns), f, indent=1, sort_keys=True)
    print("returns", returns, "history keys", list(ref.history))


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    main()
