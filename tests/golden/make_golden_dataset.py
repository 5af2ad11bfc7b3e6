"""Record what the REFERENCE's own `NeRFDataset` (datasets/nerf_dataset.py, imported from /root/reference) returns for
datasets written by nerf_slam_b200.datasets.write_transforms_dataset — build container only.

  python tests/golden/make_golden_dataset.py        ->  tests/golden/ref_dataset_packets.json

Stubbed third-party imports (not installable, not executed on this path): open3d, icecream."""
import json
import os
import sys
import tempfile
import types

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    o3d = types.ModuleType("open3d"); o3d.geometry = types.SimpleNamespace(PointCloud=object)
    ic = types.ModuleType("icecream"); ic.ic = lambda *a, **k: None
    sys.modules.setdefault("open3d", o3d); sys.modules.setdefault("icecream", ic)
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import dataset_scenario as sc
    sys.path.insert(0, REF)
    from datasets.nerf_dataset import NeRFDataset        # the reference's class
    out = {}
    written = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, w, h, n, largs in sc.CASES:
            key = (w, h, n)
            if key not in written:
                written[key] = os.path.join(tmp, f"ds_{w}x{h}_{n}")
                sc.write_case(name, w, h, n, written[key])
            args = sc.loader_args(written[key], **largs)
            ds = NeRFDataset(args, "cpu")
            exact = not ds.resize_images
            out[name] = {"len": len(ds), "world_T_imu_t0": args.world_T_imu_t0.tolist(),
                         "packets": [sc.digest_packet(ds[k], exact) for k in range(len(ds))]}
            print(name, "frames", len(ds), "resized" if ds.resize_images else "native",
                  out[name]["packets"][0]["image_shape"], out[name]["packets"][0]["resolution"])
    path = os.path.join(HERE, "ref_dataset_packets.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
