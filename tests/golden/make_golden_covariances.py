"""Record the output of the REFERENCE's own covariance extraction (slam/visual_frontends/visual_frontend.py, the
`if compute_covariances:` block of RaftVisualFrontend.ba, A14) for a seeded BA window — build container only.

  python tests/golden/make_golden_covariances.py        ->  tests/golden/ref_covariances.npz

The block is read from the reference file at run time (between the marker lines below), dedented and executed with the
variables it expects (`linear_factor_graph.hessian()`, E, Q, ii, jj, kf0, kf1, N, HW, pose_keys, `self.*` buffers);
nothing of it is copied into this repository.  Only `gtsam.Symbol(key).index()` — a key -> frame-id lookup — is stubbed.
The inputs come from the oracle's reduced camera matrix of a seeded window (any SPD H and consistent E, Q would do)."""
import os
import sys
import textwrap
import types

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
START = "            H, v = linear_factor_graph.hessian()"
END = "                self.cam0_depths_cov[kx] = depth_cov"


def problem(seed=21):
    sys.path.insert(0, ROOT)
    from oracle import ba as oba
    from tests.util import make_targets, make_window
    rng = np.random.default_rng(seed)
    nframes, ht, wd = 6, 8, 12
    poses, disps, intr, ii, jj = make_window(rng, nframes, ht, wd, extra_edges=3)
    target, weight = make_targets(rng, poses, disps, intr, ii, jj)
    kf0, kf1 = 2, nframes                                 # frames 0,1 fixed: more depth maps (K = 6) than poses (P = 4)
    kx = np.unique(np.concatenate([np.arange(kf0, kf1), ii]))
    eta = rng.uniform(1e-3, 1e-1, (len(kx), ht, wd)).astype(np.float32)
    ext = np.array([0, 0, 0, 0, 0, 0, 1.0], np.float32)
    r = oba.reduced_camera_matrix(poses, disps, intr, ext, np.zeros_like(disps), target, weight, eta, ii, jj, kf0, kf1)
    H = r["H"] + 1e-3 * np.eye(r["H"].shape[0])           # SPD system as gtsam's hessian() would return it
    return dict(H=H, E=r["E"].astype(np.float32), Q=r["Q"].astype(np.float32), ii=ii, jj=jj, kf0=kf0, kf1=kf1,
                disps=disps, ht=ht, wd=wd)


def main():
    p = problem()
    lines = open(os.path.join(REF, "slam", "visual_frontends", "visual_frontend.py")).read().split("\n")
    a = lines.index(START); b = lines.index(END)
    code = textwrap.dedent("\n".join(lines[a:b + 1]))
    N = p["kf1"] - p["kf0"]
    nb = p["disps"].shape[0]
    self = types.SimpleNamespace(device="cpu", ht=p["ht"], wd=p["wd"], cam0_idepths=torch.from_numpy(p["disps"]).clone(),
                                 f_idx_to_kf_idx={10 + k: p["kf0"] + k for k in range(N)},
                                 world_T_body_cov=torch.zeros(nb, 6, 6), cam0_idepths_cov=torch.zeros(nb, p["ht"], p["wd"]),
                                 cam0_depths_cov=torch.zeros(nb, p["ht"], p["wd"]))
    gtsam = types.SimpleNamespace(Symbol=lambda key: types.SimpleNamespace(index=lambda: key))
    ns = dict(torch=torch, self=self, gtsam=gtsam, N=N, HW=p["ht"] * p["wd"], kf0=p["kf0"], kf1=p["kf1"],
              ii=torch.from_numpy(p["ii"]), jj=torch.from_numpy(p["jj"]), E=torch.from_numpy(p["E"]), Q=torch.from_numpy(p["Q"]),
              pose_keys=[10 + k for k in range(N)],
              linear_factor_graph=types.SimpleNamespace(hessian=lambda: (p["H"], np.zeros(6 * N))), print=print)
    exec(code, ns)
    kx = torch.unique(torch.from_numpy(p["ii"]))
    out = dict(H=p["H"], E=p["E"], Q=p["Q"], ii=p["ii"], jj=p["jj"], kf=np.array([p["kf0"], p["kf1"]]), disps=p["disps"],
               sigma_g=self.world_T_body_cov[p["kf0"]:p["kf1"]].numpy(), kx=kx.numpy(),
               z_cov=self.cam0_idepths_cov[kx].numpy(), depth_cov=self.cam0_depths_cov[kx].numpy())
    np.savez_compressed(os.path.join(HERE, "ref_covariances.npz"), **out)
    print("block lines", b - a + 1, "P", N, "K", len(kx), "z_cov mean", float(out["z_cov"].mean()), "sigma_g trace", float(np.trace(out["sigma_g"][0])))


if __name__ == "__main__":
    main()
