"""Record the output of the REFERENCE's own `projective_transform` (networks/geom/projective_ops.py:98-145, the live
path's reprojection, jacobian=False) for seeded inputs — build container only.

  python tests/golden/make_golden_reproject.py        ->  tests/golden/ref_reproject.npz

`lietorch` cannot be installed; the three group operations the function uses (indexing, inverse, composition / action on
homogeneous points with inverse depth) are provided by a stand-in written from the SE3 formulas that the reference
duplicates in src/droid_kernels.cu:66-120.  Everything else — inverse projection, the stereo special case, the
MIN_DEPTH rules, projection, the validity mask — is the reference's code."""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _qmul(a, b):
    ax, ay, az, aw = a.unbind(-1); bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _qrot(q, v):
    u = q[..., :3]; w = q[..., 3:4]
    uv = torch.cross(u, v, dim=-1)
    return v + 2 * (w * uv + torch.cross(u, uv, dim=-1))


class SE3:
    """data [..., 7] = (t, q_xyzw)"""

    def __init__(self, data):
        self.data = data

    device = property(lambda self: self.data.device)

    def __getitem__(self, idx):
        return SE3(self.data[idx])

    def inv(self):
        t, q = self.data[..., :3], self.data[..., 3:]
        qi = q * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=q.dtype)
        return SE3(torch.cat([-_qrot(qi, t), qi], -1))

    def __mul__(self, other):
        t, q = self.data[..., :3], self.data[..., 3:]
        if isinstance(other, SE3):
            return SE3(torch.cat([_qrot(q, other.data[..., :3]) + t, _qmul(q, other.data[..., 3:])], -1))
        X = other                                           # homogeneous points (X, Y, Z, inverse depth)
        q, t = q.expand(X.shape[:-1] + (4,)), t.expand(X.shape[:-1] + (3,))
        return torch.cat([_qrot(q, X[..., :3]) + t * X[..., 3:4], X[..., 3:4]], -1)


def inputs(seed=7, n=6, ht=12, wd=16):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(n, 4)) * 0.15 + np.array([0, 0, 0, 1.0]); q /= np.linalg.norm(q, axis=1, keepdims=True)
    poses = np.concatenate([rng.normal(size=(n, 3)) * 0.3, q], 1).astype(np.float32)
    disps = rng.uniform(0.05, 2.0, (n, ht, wd)).astype(np.float32)
    disps[1, :2] = 6.0                                      # very close points: behind / near the target camera
    intr = np.stack([np.array([20.0 + k, 21.0 + k, 7.5, 5.5], np.float32) for k in range(n)])
    ii = np.array([0, 1, 2, 3, 4, 5, 2, 1, 0]); jj = np.array([1, 0, 3, 2, 5, 4, 2, 4, 5])      # incl. a stereo edge (2,2)
    poses[5, :3] += np.array([0, 0, 1.2], np.float32)       # a camera moved forward past some points
    return poses, disps, intr, ii, jj


def main():
    lt = types.ModuleType("lietorch"); lt.SE3 = SE3; lt.Sim3 = type("Sim3", (), {})
    ic = types.ModuleType("icecream"); ic.ic = lambda *a, **k: None
    sys.modules.setdefault("lietorch", lt); sys.modules.setdefault("icecream", ic)
    sys.path.insert(0, REF)
    import warnings
    warnings.filterwarnings("ignore")
    from networks.geom import projective_ops as pops
    poses, disps, intr, ii, jj = inputs()
    x1, valid, _ = pops.projective_transform(SE3(torch.from_numpy(poses)[None]), torch.from_numpy(disps)[None],
                                             torch.from_numpy(intr)[None], torch.from_numpy(ii), torch.from_numpy(jj))
    out = dict(poses=poses, disps=disps, intr=intr, ii=ii, jj=jj, coords=x1[0].numpy(), valid=valid[0].numpy(),
               min_depth=np.array([pops.MIN_DEPTH]))
    np.savez_compressed(os.path.join(HERE, "ref_reproject.npz"), **out)
    print("coords", out["coords"].shape, "valid fraction", float(out["valid"].mean()), "min depth", pops.MIN_DEPTH)


if __name__ == "__main__":
    main()
