"""Record what the REFERENCE's `NerfFusion.process_slam` + `send_data` (fusion/nerf_fusion.py:140-289) hand to
`ngp.nerf.training.update_training_images` for a seeded SLAM packet, for every mask type — build container only.

  python tests/golden/make_golden_process_slam.py        ->  tests/golden/ref_process_slam.npz

The class is imported from /root/reference; stubs only for what cannot be installed: `pyngp` (a recorder standing where
the trainer is), `lietorch.SE3` (pose (t, q_xyzw) -> 4x4 matrix, the one method used), icecream."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


class SE3:
    def __init__(self, data):
        self.data = data

    def matrix(self):
        t, q = self.data[:, :3].double(), self.data[:, 3:].double()
        x, y, z, w = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).view(-1, 3, 3)
        T = torch.eye(4, dtype=torch.float64).repeat(len(t), 1, 1)
        T[:, :3, :3] = R; T[:, :3, 3] = t
        return T.float()               # lietorch returns the dtype of its data (fp32 poses)


def main():
    lt = types.ModuleType("lietorch"); lt.SE3 = SE3
    ic = types.ModuleType("icecream"); ic.ic = lambda *a, **k: None
    png = types.ModuleType("pyngp")
    for name, m in (("lietorch", lt), ("icecream", ic), ("pyngp", png)):
        sys.modules.setdefault(name, m)
    sys.path.insert(0, REF); sys.path.insert(0, HERE)
    from fusion.nerf_fusion import NerfFusion
    import process_slam_scenario as sc
    out = {}
    for mt in sc.MASK_TYPES:
        calls = []
        training = types.SimpleNamespace(update_training_images=lambda *a: calls.append(a))
        nf = object.__new__(NerfFusion)                     # __init__ needs the real trainer; process_slam does not
        nf.mask_type, nf.device, nf.viz = mt, "cpu", False
        nf.ngp = types.SimpleNamespace(nerf=types.SimpleNamespace(training=training))
        training.optimize_extrinsics = True
        assert nf.process_slam([None, sc.make_packet()]) is False and len(calls) == 1
        ids, poses, images, depths, covs, res, pp, fl, dscale, cscale = calls[0]
        out[f"{mt}.ids"] = np.asarray(ids); out[f"{mt}.poses"] = np.stack(poses).astype(np.float64)
        out[f"{mt}.images"] = np.stack(images); out[f"{mt}.depths"] = np.stack(depths); out[f"{mt}.covs"] = np.stack(covs)
        out[f"{mt}.res"] = np.asarray(res); out[f"{mt}.pp"] = np.asarray(pp); out[f"{mt}.fl"] = np.asarray(fl)
        out[f"{mt}.scales"] = np.array([dscale, cscale], np.float64)
        print(mt, out[f"{mt}.images"].shape, out[f"{mt}.images"].dtype, out[f"{mt}.depths"].dtype, float(out[f"{mt}.depths"].min()))
    last = sc.make_packet(); last["is_last_frame"] = True
    nf.mask_type = "ours"
    out["last_frame_skipped"] = np.array([nf.process_slam([None, last]) is True and len(calls) == 1])
    # process_data (GT fitting, :121-138).  send_data calls `frame_ids.cpu()` on batch["k"]: the data path hands it a
    # numpy array, for which that fails in the reference as shipped; the recorder receives a tensor copy of "k".
    calls.clear()
    pkt = sc.make_data_packet()
    pkt["k"] = torch.from_numpy(pkt["k"])
    assert nf.process_data(pkt) is False and len(calls) == 1
    ids, poses, images, depths, covs, res, pp, fl, dscale, cscale = calls[0]
    out["data.ids"] = np.asarray(ids); out["data.poses"] = np.stack(poses).astype(np.float64)
    out["data.images"] = np.stack(images); out["data.depths"] = np.stack(depths); out["data.covs"] = np.stack(covs)
    out["data.res"] = np.asarray(res); out["data.pp"] = np.asarray(pp); out["data.fl"] = np.asarray(fl)
    out["data.scales"] = np.array([dscale, cscale], np.float64)
    print("data", out["data.images"].shape, out["data.images"].dtype, out["data.poses"][0, :, 3], out["data.scales"])
    np.savez_compressed(os.path.join(HERE, "ref_process_slam.npz"), **out)
    print("wrote ref_process_slam.npz", os.path.getsize(os.path.join(HERE, "ref_process_slam.npz")))


if __name__ == "__main__":
    main()
