"""Input side: nerf_slam_b200.datasets.NeRFDataset must return, frame for frame, what the
REFERENCE's own NeRFDataset returned for the same files (tests/golden/ref_dataset_packets.json, recorded by
tests/golden/make_golden_dataset.py): packet schema, dtypes, nerf->ngp pose convention, slicing by
initial_k/final_k/img_stride, the > 640x640 down-scaling rule with rescaled intrinsics.  Lossless path: identical
bytes (sha1); resized path: 8x8 block means within half a grey level (cv2.resize may differ by an LSB across CPUs)."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import dataset_scenario as sc   # noqa: E402

from nerf_slam_b200 import datasets as ds   # noqa: E402

with open(os.path.join(HERE, "golden", "ref_dataset_packets.json")) as f:
    GOLD = json.load(f)


@pytest.fixture(scope="module")
def written(tmp_path_factory):
    root = tmp_path_factory.mktemp("nerf_datasets")
    out = {}
    for name, w, h, n, _ in sc.CASES:
        if (w, h, n) not in out:
            out[(w, h, n)] = str(root / f"ds_{w}x{h}_{n}")
            sc.write_case(name, w, h, n, out[(w, h, n)])
    return out


@pytest.mark.parametrize("case", sc.CASES, ids=[c[0] for c in sc.CASES])
def test_reader_matches_reference_packets(case, written):
    name, w, h, n, largs = case
    gold = GOLD[name]
    args = sc.loader_args(written[(w, h, n)], **largs)
    data = ds.NeRFDataset(args, "cpu")
    assert len(data) == gold["len"]
    assert np.allclose(args.world_T_imu_t0, gold["world_T_imu_t0"], atol=1e-12)
    exact = not data.resize_images
    for k in range(len(data)):
        got, ref = sc.digest_packet(data[k], exact), gold["packets"][k]
        for key in ("k", "t_cams", "is_last_frame", "image_shape", "image_dtype", "depth_shape", "depth_dtype",
                    "resolution", "rate_hz"):
            assert got[key] == ref[key], (name, k, key)
        for key in ("poses", "intrinsics", "aabb", "depth_scale"):
            assert np.allclose(got[key], ref[key], rtol=0, atol=1e-12), (name, k, key)
        if exact:
            assert got["image_sha1"] == ref["image_sha1"] and got["depth_sha1"] == ref["depth_sha1"], (name, k)
        else:
            assert np.abs(np.array(got["image_means"]) - np.array(ref["image_means"])).max() < 0.5
            assert np.abs(np.array(got["depth_means"]) - np.array(ref["depth_means"])).max() < 2.0


def test_written_stream_equals_the_procedural_stream(written):
    """the files are a lossless copy of synthetic.SyntheticRoom: bench / tests and the reference CLI see the same frames"""
    from nerf_slam_b200.synthetic import SyntheticRoom
    room = SyntheticRoom(64, 48, 6, seed=3, step=0.05)
    data = ds.NeRFDataset(ds.dataset_args(written[(64, 48, 6)]))
    for k, p in enumerate(data.stream()):
        q = room.packet(k)
        assert np.array_equal(p["images"], q["images"]) and np.array_equal(p["depths"], q["depths"])
        assert np.allclose(p["poses"], q["poses"], atol=1e-12)
        assert p["is_last_frame"] == q["is_last_frame"] and p["depths"].dtype == np.int32


def test_pose_convention_round_trip():
    rng = np.random.default_rng(0)
    for _ in range(5):
        m = np.eye(4); m[:3] = rng.normal(size=(3, 4))
        assert np.allclose(ds.nerf_matrix_to_ngp(ds.ngp_matrix_to_nerf(m)), m, atol=1e-12)
        g = ds.nerf_matrix_to_ngp(m)
        # utils/utils.py:104-116 spelled out: columns 1,2 negated, position + 0.5, rows cycled (y, z, x)
        e = m.copy(); e[:3, 1] *= -1; e[:3, 2] *= -1; e[:3, 3] += 0.5
        assert np.allclose(g, e[[1, 2, 0, 3]], atol=1e-15)


def test_demo_cli_has_the_reference_options_and_defaults(written):
    """examples/slam_demo.py: option names and defaults of the reference's parse_args (examples/slam_demo.py:20-55)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("slam_demo", os.path.join(os.path.dirname(HERE), "examples", "slam_demo.py"))
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)
    a = demo.parse_args([])
    ref_defaults = dict(parallel_run=False, multi_gpu=False, initial_k=0, final_k=-1, img_stride=1, stereo=False,
                        weights="droid.pth", buffer=512, dataset_dir="/home/tonirv/Datasets/euroc/V1_01_easy",
                        dataset_name="euroc", mask_type="ours", slam=False, fusion="", gui=False, width=0, height=0,
                        network="", eval=False)
    for k, v in ref_defaults.items():
        assert getattr(a, k) == v, k
    a = demo.parse_args(["--dataset_dir", written[(64, 48, 6)], "--dataset_name", "nerf", "--buffer", "100", "--slam",
                         "--fusion", "nerf", "--screenshot_w", "320"])
    assert a.slam and a.fusion == "nerf" and a.buffer == 100 and a.width == 320
    data = demo.make_data(a)
    assert len(data) == 5                 # --final_k=-1 drops the last of the 6 frames, as in the reference
    assert np.allclose(a.world_T_imu_t0, data[0]["poses"][0])
    with pytest.raises(NotImplementedError):
        demo.make_data(demo.parse_args(["--dataset_name", "euroc"]))
