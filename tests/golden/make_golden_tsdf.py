"""Golden vectors for the Sigma / TSDF fusion rule, produced by EXECUTING the reference's own methods.

    python tests/golden/make_golden_tsdf.py        (needs /root/reference; writes tests/golden/ref_tsdf_integrate.npz)

`fusion/tsdf_fusion.py` cannot be imported (Open3D, lietorch, icecream are absent), so — like make_golden_live_frontend.py —
the source text of `TsdfFusion.build_volume`, `custom_volume_integrate` and `get_depth_masks` is cut out of the reference file
with `ast`, compiled unchanged into a class of the same name, and run on stand-ins:
  * `o3d.core.Tensor` / `o3d.t.geometry.Image`: thin wrappers around torch CPU tensors with exactly the operations those
    methods use (`to`, `T`, `@`, `round`, indexing with tensors, views through `reshape`, comparisons, arithmetic);
  * `self.volume`: a DENSE voxel grid standing in for Open3D's `VoxelBlockGrid`: every voxel is "active", voxel coordinates
    are float32 `origin + voxel_size * index` (Open3D hands out float32 metric coordinates), attributes are flat torch tensors;
  * lietorch's `SE3(poses).matrix()`: the closed form.
What the golden pins is therefore the reference's per-voxel arithmetic (fp64 projection, `round()`, masks, fp32 running
averages, weight saturation, the masking of uncertain pixels in `build_volume`), NOT Open3D's choice of active blocks."""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/fusion/tsdf_fusion.py"
WANTED = ("build_volume", "custom_volume_integrate", "get_depth_masks")


def _unwrap(x):
    if isinstance(x, T):
        return x.t
    if isinstance(x, tuple):
        return tuple(_unwrap(v) for v in x)
    return x


class T:
    """stand-in for o3d.core.Tensor"""

    def __init__(self, x):
        if isinstance(x, T):
            x = x.t
        self.t = x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))

    @staticmethod
    def from_dlpack(capsule):
        return T(torch.utils.dlpack.from_dlpack(capsule))

    shape = property(lambda self: tuple(self.t.shape))

    def to(self, *args):
        out = self.t
        for a in args:
            if isinstance(a, torch.dtype):
                out = out.to(a)
        return T(out)

    def T(self):
        return T(self.t.T)

    def round(self):
        return T(self.t.round())                       # half-to-even, like Open3D's Round

    def reshape(self, shape):
        return T(self.t.reshape(tuple(shape)))         # a VIEW for contiguous storage: writes reach the volume

    def __getitem__(self, idx):
        return T(self.t[_unwrap(idx)])

    def __setitem__(self, idx, val):
        self.t[_unwrap(idx)] = _unwrap(val)

    def _bin(self, other, op):
        return T(op(self.t, _unwrap(other)))

    __matmul__ = lambda s, o: s._bin(o, torch.matmul)
    __add__ = lambda s, o: s._bin(o, torch.add)
    __sub__ = lambda s, o: s._bin(o, torch.sub)
    __mul__ = lambda s, o: s._bin(o, torch.mul)
    __rmul__ = lambda s, o: s._bin(o, torch.mul)
    __truediv__ = lambda s, o: s._bin(o, torch.div)
    __and__ = lambda s, o: s._bin(o, torch.logical_and)
    __gt__ = lambda s, o: s._bin(o, torch.gt)
    __ge__ = lambda s, o: s._bin(o, torch.ge)
    __lt__ = lambda s, o: s._bin(o, torch.lt)
    __neg__ = lambda s: T(-s.t)


class Image:
    """stand-in for o3d.t.geometry.Image: a 2-D tensor gets its channel dimension"""

    def __init__(self, tensor):
        t = tensor.t
        self._t = t[..., None] if t.dim() == 2 else t
        self.rows, self.columns = self._t.shape[0], self._t.shape[1]

    def as_tensor(self):
        return T(self._t)


class DenseVolume:
    """every voxel of an n^3 grid active; layout [z][y][x]"""

    def __init__(self, n, origin, voxel_size):
        self.n, self.origin, self.vs = n, np.asarray(origin, np.float32), np.float32(voxel_size)
        self.attr = {"tsdf": torch.zeros(n ** 3, 1), "weight": torch.zeros(n ** 3, 1), "color": torch.zeros(n ** 3, 3)}

    def compute_unique_block_coordinates(self, depth, intrinsic, extrinsic, depth_scale, max_depth):
        return "all blocks"

    def hashmap(self):
        vol = self

        class H:
            def activate(self, coords):
                pass

            def find(self, coords):
                return "all buffers", None
        return H()

    def voxel_coordinates_and_flattened_indices(self, buf_indices):
        n = self.n
        iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        idx = np.stack([ix, iy, iz], -1).reshape(-1, 3).astype(np.float32)
        coords = (self.vs * idx + self.origin[None]).astype(np.float32)       # float32 products, then float32 sums
        return T(torch.from_numpy(coords)), T(torch.arange(n ** 3))

    def attribute(self, name):
        return T(self.attr[name])


class SE3:
    def __init__(self, data):
        self.data = data

    def matrix(self):
        t, q = self.data[..., :3].double(), self.data[..., 3:].double()
        x, y, z, w = q.unbind(-1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
        M = torch.eye(4, dtype=torch.float64).repeat(len(R), 1, 1)
        M[:, :3, :3] = R
        M[:, :3, 3] = t
        return M.float()                                 # lietorch returns the dtype of its data (fp32 poses)


def reference_class():
    src = open(REF).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TsdfFusion")
    funcs = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert len(funcs) == len(WANTED)
    body = "\n".join(ast.get_source_segment(src, f) for f in funcs)
    code = "class TsdfFusion:\n" + "\n".join("    " + l if not l.startswith("    ") else l for l in body.split("\n"))
    o3d = types.SimpleNamespace(
        core=types.SimpleNamespace(Tensor=T, float64=torch.float64, float32=torch.float32, int64=torch.int64,
                                   cuda=types.SimpleNamespace(synchronize=lambda: None)),
        t=types.SimpleNamespace(geometry=types.SimpleNamespace(Image=Image)),
        geometry=types.SimpleNamespace())
    ns = {"o3d": o3d, "torch": torch, "np": np, "SE3": SE3, "ic": lambda *a, **k: None, "print": lambda *a, **k: None}
    exec(compile(code, REF, "exec"), ns)
    return ns["TsdfFusion"]


def scenario(seed=11, n=24, H=30, W=40, frames=3):
    """inputs shared with the tests: voxel lattice through the world origin (as Open3D's), camera inside the grid"""
    rng = np.random.default_rng(seed)
    vs = 0.08
    origin = (-n // 2 * np.float32(vs) * np.ones(3, np.float32)).astype(np.float32) + np.array([0, 0, 1.0], np.float32)
    idepths = rng.uniform(0.45, 1.6, (frames, H, W)).astype(np.float32)
    covs = rng.uniform(0.004, 0.6, (frames, H, W)).astype(np.float32)
    covs[:, :3] = 4e8                                    # sqrt > 10000: masked ("uncertainty") or weight ~ 0 ("uniform")
    imgs = rng.integers(0, 255, (frames, 3, H, W), dtype=np.uint8)
    q = rng.normal(size=(frames, 4)) * 0.08 + np.array([0, 0, 0, 1.0]); q /= np.linalg.norm(q, axis=1, keepdims=True)
    poses = np.concatenate([rng.uniform(-0.15, 0.15, (frames, 3)), q], 1).astype(np.float32)
    intr = np.array([[W * 0.55, 0, W / 2 - 0.5], [0, W * 0.55, H / 2 - 0.5], [0, 0, 1]], np.float64)
    return dict(n=n, voxel_size=np.float32(vs), origin=origin, idepths=idepths, covs=covs, imgs=imgs, poses=poses, intr=intr)


def run_reference(sc, mask_type, max_weight):
    Ref = reference_class()
    self = Ref.__new__(Ref)
    self.device, self.o3d_device = "cpu", "cpu"
    self.depth_scale, self.max_depth, self.sdf_trunc, self.max_weight = 1.0, 6.0, 0.10, max_weight
    self.max_depth_sigma_thresh, self.use_old_volume, self.depth_mask_type = 10000, False, mask_type
    self.volume = DenseVolume(sc["n"], sc["origin"], sc["voxel_size"])
    snaps = []
    for k in range(len(sc["poses"])):                    # one keyframe per packet: snapshots in between
        packet = {"cam0_poses": torch.from_numpy(sc["poses"][k:k + 1]), "cam0_idepths_up": torch.from_numpy(sc["idepths"][k:k + 1]),
                  "cam0_depths_cov_up": torch.from_numpy(sc["covs"][k:k + 1].copy()), "cam0_images": torch.from_numpy(sc["imgs"][k:k + 1])}
        if mask_type == "uniform":                       # rebuild_volume's rule for the "tsdf" flavour (:228)
            packet["cam0_depths_cov_up"] = torch.ones_like(packet["cam0_depths_cov_up"])
        o3d_intr = types.SimpleNamespace(intrinsic_matrix=sc["intr"])
        self.build_volume(packet, o3d_intr, self.get_depth_masks(packet))
        snaps.append({k2: v.clone().numpy() for k2, v in self.volume.attr.items()})
    return snaps


def main():
    sc = scenario()
    out = {k: v for k, v in sc.items()}
    for mask_type, tag in (("uncertainty", "sigma"), ("uniform", "tsdf")):
        # max_weight 20 (the reference's value) never saturates in "sigma" here, 2.5 does
        for mw in (20.0, 2.5):
            snaps = run_reference(sc, mask_type, mw)
            n = sc["n"]
            for k, s in enumerate(snaps):
                out[f"{tag}_w{mw}_tsdf_{k}"] = s["tsdf"].reshape(n, n, n)
                out[f"{tag}_w{mw}_weight_{k}"] = s["weight"].reshape(n, n, n)
                out[f"{tag}_w{mw}_color_{k}"] = s["color"].reshape(n, n, n, 3)
            touched = int((snaps[-1]["weight"] > 0).sum())
            print(tag, "max_weight", mw, "voxels touched", touched, "max weight", float(snaps[-1]["weight"].max()))
    np.savez_compressed(os.path.join(HERE, "ref_tsdf_integrate.npz"), **out)
    print("wrote ref_tsdf_integrate.npz", os.path.getsize(os.path.join(HERE, "ref_tsdf_integrate.npz")) // 1024, "KB")


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    main()
