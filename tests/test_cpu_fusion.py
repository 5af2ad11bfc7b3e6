"""B1 / B2: SLAM packet -> training tuples.  Golden = what the REFERENCE's own NerfFusion.process_slam +
send_data hand to `update_training_images` (tests/golden/ref_process_slam.npz, recorded by
tests/golden/make_golden_process_slam.py).  Checked here on the CPU: the oracle restatement, and the host half of the
product's process_slam (pose conversion, mask types, slot ids, intrinsics) with the trainer replaced by a recorder; the
device half (sRGB->linear, 1/idepth in one kernel) is compared with the same golden in tests/test_gpu_ngp.py."""
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import process_slam_scenario as sc   # noqa: E402

from oracle import ngp as ongp   # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "ref_process_slam.npz"))


@pytest.mark.parametrize("mt", sc.MASK_TYPES)
def test_oracle_process_slam_matches_reference(mt):
    got = ongp.process_slam_tuples(sc.make_packet(), mt)
    assert got["ids"] == GOLD[f"{mt}.ids"].tolist()
    assert np.allclose(got["poses"], GOLD[f"{mt}.poses"], rtol=0, atol=2e-6)
    assert np.array_equal(got["images"], GOLD[f"{mt}.images"])            # same fp32 torch ops -> identical
    assert np.array_equal(got["depths"], GOLD[f"{mt}.depths"])
    assert np.array_equal(got["depths_cov"], GOLD[f"{mt}.covs"])
    assert tuple(GOLD[f"{mt}.scales"]) == got["scales"]
    if mt in ("ours_w_thresh", "no_depth"):
        assert (got["depths"] == -1.0).any()


@pytest.mark.parametrize("mt", sc.MASK_TYPES)
def test_product_process_slam_host_half(mt):
    from nerf_slam_b200.nerf_fusion import NerfFusion
    calls = []
    nf = object.__new__(NerfFusion)
    nf.mask_type, nf.ref_frames = mt, {}
    nf.evaluate, nf._gt_depths, nf.args = True, None, types.SimpleNamespace(buffer=32)
    training = types.SimpleNamespace(update_training_images_device=lambda *a, **kw: calls.append(a))
    nf.ngp = types.SimpleNamespace(device="cpu", nerf=types.SimpleNamespace(training=training))
    assert nf.process_slam([None, sc.make_packet()]) is False and len(calls) == 1
    ids, c2w, images, idepths, covs, focal, pp = calls[0]
    assert [int(i) for i in ids] == GOLD[f"{mt}.ids"].tolist()
    assert np.allclose(np.asarray(c2w), GOLD[f"{mt}.poses"], rtol=0, atol=2e-6)
    assert images.dtype == torch.uint8 and torch.equal(images, sc.make_packet()["cam0_images"])
    # what the ingest kernel computes from these operands is the golden's depth / covariance
    assert np.allclose((1.0 / idepths).numpy()[..., None], GOLD[f"{mt}.depths"], rtol=1e-6, atol=0)
    assert np.array_equal(covs.numpy()[..., None], GOLD[f"{mt}.covs"])
    assert np.allclose(focal, GOLD[f"{mt}.fl"]) and np.allclose(pp, GOLD[f"{mt}.pp"])
    # the packet itself is left untouched (the reference mutates its input tensors in place)
    # ... and is not retained: only the GT depth maps of the ingested frame ids are kept (for eval_gt_traj)
    ids_g = GOLD[f"{mt}.ids"].tolist()
    assert sorted(nf.ref_frames) == ids_g
    assert torch.equal(nf._gt_depths[ids_g], sc.make_packet()["gt_depths"][:, 0].float())
    assert not any(torch.is_tensor(v) or isinstance(v, dict) for v in nf.ref_frames.values())


def test_last_frame_packet_is_not_ingested():
    from nerf_slam_b200.nerf_fusion import NerfFusion
    assert bool(GOLD["last_frame_skipped"][0])
    nf = object.__new__(NerfFusion)
    nf.mask_type, nf.ref_frames = "ours", {}
    nf.ngp = types.SimpleNamespace(device="cpu", nerf=types.SimpleNamespace(training=types.SimpleNamespace(
        update_training_images_device=lambda *a, **kw: (_ for _ in ()).throw(AssertionError("ingested")))))
    last = sc.make_packet(); last["is_last_frame"] = True
    assert nf.process_slam([None, last]) is True and nf.process_slam(None) is True and nf.process_slam([None, None]) is True


def test_eval_metric_matches_reference_compute_error():
    """B4 metric (fusion/nerf_fusion.py:424-425, utils/utils.py:168-188): MSE over all four channels with the estimate
    clamped at 0 and non-finite values zeroed; value recorded from the reference's own compute_error / mse2psnr on these
    seeded arrays (generator: the snippet in this docstring, run against /root/reference/utils/utils.py):
        rng = default_rng(3); est = U(-0.2, 1.2, (12,16,4)) f32; ref = U(0, 1, (12,16,4)) f32;
        est[0,0,0] = nan; est[1,2,3] = inf; est[3,3,1] = -inf  ->  0.2410338968038559, 6.17921878931163 dB"""
    from nerf_slam_b200.nerf_fusion import compute_error, mse2psnr
    rng = np.random.default_rng(3)
    est = rng.uniform(-0.2, 1.2, (12, 16, 4)).astype(np.float32); ref = rng.uniform(0, 1, (12, 16, 4)).astype(np.float32)
    est[0, 0, 0] = np.nan; est[1, 2, 3] = np.inf; est[3, 3, 1] = -np.inf
    e = compute_error(est, ref)
    assert abs(e - 0.2410338968038559) < 1e-7 and abs(mse2psnr(e) - 6.17921878931163) < 1e-5
    assert np.isnan(est[0, 0, 0])                                     # the caller's array is left untouched


def test_product_process_data_matches_reference():
    """GT-fitting path (fusion/nerf_fusion.py:121-138): the tuples handed to update_training_images — poses scaled /
    offset into the unit cube by the calibration's aabb, colours u8/255 without linearisation, depth scale times the pose
    scale, unit covariances — equal what the reference's own process_data + send_data produced."""
    from nerf_slam_b200.nerf_fusion import NerfFusion, get_scale_and_offset
    calls = []
    nf = object.__new__(NerfFusion)
    training = types.SimpleNamespace(update_training_images=lambda *a: calls.append(a), optimize_extrinsics=True)
    nf.ngp = types.SimpleNamespace(nerf=types.SimpleNamespace(training=training))
    assert nf.process_data(sc.make_data_packet()) is False and len(calls) == 1 and training.optimize_extrinsics is False
    ids, poses, images, depths, covs, res, pp, fl, dscale, cscale = calls[0]
    assert list(ids) == GOLD["data.ids"].tolist()
    assert np.allclose(np.stack(poses), GOLD["data.poses"], rtol=0, atol=1e-9)
    assert np.array_equal(np.stack(images), GOLD["data.images"]) and np.stack(images).dtype == np.float32
    assert np.array_equal(np.stack(depths), GOLD["data.depths"]) and np.array_equal(np.stack(covs), GOLD["data.covs"])
    assert np.array_equal(res, GOLD["data.res"]) and np.allclose(pp, GOLD["data.pp"]) and np.allclose(fl, GOLD["data.fl"])
    assert np.allclose([dscale, cscale], GOLD["data.scales"], rtol=1e-12)
    s, o = get_scale_and_offset([[-2.0, -1.0, -2.0], [2.0, 3.0, 2.0]])
    assert s == 0.25 and np.allclose(o, [0.5, 0.25, 0.5])


def test_tsdf_fusion_history_matches_the_reference_methods():
    """TsdfFusion.update_history / get_history_packet against the reference's own methods executed verbatim
    (tests/golden/ref_tsdf_history.json, make_golden_tsdf_history.py): overlapping dirty windows overwrite their entries,
    insertion order and frame-id keys, the last-frame packet is refused, stacked packet shapes and contents"""
    import json
    import types
    from nerf_slam_b200.tsdf_fusion import TsdfFusion
    from tests.golden import make_golden_tsdf_history as mk
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tsdf_history.json")) as f:
        ref = json.load(f)
    fus = TsdfFusion("sigma", types.SimpleNamespace(eval=False, tsdf_resolution=8), device="cpu")
    returns = [bool(fus.update_history(p)) for p in mk.packets()]
    got = json.loads(json.dumps(mk.summarize(fus, returns)))
    gt_ref, gt_got = ref["packet"].pop("gt_depths_sum"), got["packet"].pop("gt_depths_sum")
    assert got == ref
    # same values x depth_scale; the reference additionally applies `.permute(2,0,1)` to the [1,H,W] maps (:528, its own TODO
    # questions that line) — consumed only by its Open3D evaluation rendering, which is out of scope
    assert abs(gt_ref - gt_got) < 1e-3


def test_nerf_fusion_control_loop_matches_the_reference_methods():
    """NerfFusion.fuse / fit_volume / fit_volume_once / stop_condition against the reference's own methods executed verbatim
    around a recording `ngp` (tests/golden/ref_nerf_fusion_loop.json, make_golden_nerf_fusion_loop.py): which packets trigger
    fitting, iterations per call, annealing and evaluation cadence, the stop rule, the error for unknown packet names"""
    import json
    from nerf_slam_b200.nerf_fusion import NerfFusion
    from tests.golden import make_golden_nerf_fusion_loop as mk
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_nerf_fusion_loop.json")) as f:
        ref = json.load(f)
    for e in (False, True):
        for a in (False, True):
            got = json.loads(json.dumps(mk.drive(NerfFusion.__new__(NerfFusion), e, a)))
            assert got == ref[f"eval{int(e)}_anneal{int(a)}"], (e, a)
