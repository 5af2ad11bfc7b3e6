"""End-to-end front-end on the procedural stream (GPU): warm-up, initialisation, steady-state
updates; checks state sanity, the graph-size contract and (with droid.pth) pose accuracy."""
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS = os.path.join(ROOT, "oracle", "_ref", "droid.pth")


def _run(n_frames, W=320, H=240, buffer=24, step=0.03, impl="product", args=None):
    """impl: "product" = RaftVisualFrontend; "reference-cuda" = the reference's own kernels + library PyTorch in the
    reference's sequencing (oracle/ref_cuda_frontend.py, the A/B arm of bench.py)"""
    from nerf_slam_b200.synthetic import SyntheticRoom
    if impl == "product":
        from nerf_slam_b200.frontend import RaftVisualFrontend as Frontend
    else:
        from oracle import build_ref
        if build_ref.load("nslam_ref_droid") is None:
            pytest.skip("oracle/_ref not built")
        from oracle.ref_cuda_frontend import RefCudaFrontend as Frontend
    if args is None:
        args = types.SimpleNamespace(buffer=buffer, stereo=False, multi_gpu=False,
                                     weights=WEIGHTS if os.path.exists(WEIGHTS) else None)
    room = SyntheticRoom(W, H, n_frames, seed=0, step=step)
    fe = Frontend(np.linalg.inv(room.packet(0)["poses"][0]), np.eye(4), args, "cuda:0")
    outs = []
    for k in range(n_frames):
        x0, f, viz = fe.forward(room.packet(k))
        outs.append(viz)
        if fe.stop_condition():
            break
    return fe, room, outs


def _traj_error(fe):
    """relative error of the camera-centre distances to the first keyframe, after the monocular scale fit"""
    from oracle import se3
    n = fe.kf_idx
    est = fe.cam0_T_world[:n].cpu().numpy().astype(np.float64)
    gt = fe.gt_poses[:n].cpu().numpy().astype(np.float64)          # w2c
    c_est = np.stack([se3.inv_se3(e[:3], e[3:])[0] for e in est])   # camera centres in world
    c_gt = np.stack([np.linalg.inv(g)[:3, 3] for g in gt])
    d_est = np.linalg.norm(c_est[1:] - c_est[0], axis=-1)
    d_gt = np.linalg.norm(c_gt[1:] - c_gt[0], axis=-1)
    s = (d_est * d_gt).sum() / (d_est ** 2).sum()
    return np.abs(s * d_est - d_gt).max() / d_gt.max()


@pytest.mark.parametrize("impl", ["product", "reference-cuda"])
def test_frontend_runs_and_tracks(impl):
    fe, room, outs = _run(60, impl=impl)
    torch.cuda.synchronize()
    assert fe.is_initialized, f"not initialised after 60 frames (kf_idx={fe.kf_idx})"
    n = fe.kf_idx
    assert torch.isfinite(fe.cam0_T_world[:n]).all() and torch.isfinite(fe.cam0_idepths[:n]).all()
    assert (fe.cam0_idepths[:n] >= 1e-3).all()
    assert len(fe.ii_h) <= fe.max_factors + 2          # graph size contract
    assert len(set(zip(fe.ii_h.tolist(), fe.jj_h.tolist()))) == len(fe.ii_h)   # no duplicate edges
    if impl == "product":
        assert fe.corr_pool.capacity - len(fe.corr_pool.free) == len(fe.ii_h)   # arena accounting
        assert fe.ba_failures(wait=True) == 0
    q = fe.cam0_T_world[:n, 3:]
    assert torch.allclose(q.norm(dim=-1), torch.ones(n, device=q.device), atol=1e-4)
    viz = [v for v in outs if v is not None and "cam0_poses" in v]
    assert viz and viz[-1]["cam0_idepths_up"].shape[-2:] == (240, 320)
    if os.path.exists(WEIGHTS):
        # with the trained weights the monocular trajectory must match GT up to scale
        err = _traj_error(fe)
        assert err < 0.15, f"relative trajectory error {err:.3f}"


def test_frontend_at_the_benchmark_configuration():
    """640x480, the stream pace and the EXACT constructor arguments of bench.py (bench.make_args): initialises, tracks the
    ground-truth trajectory up to the monocular scale, no failed BA factorisation, covariances finite and positive,
    packets carry full-resolution maps"""
    import bench
    args = bench.make_args(100)
    fe, room, outs = _run(72, W=bench.W_IMG, H=bench.H_IMG, step=bench.STREAM_STEP, args=args)
    torch.cuda.synchronize()
    assert fe.is_initialized and fe.kf_idx >= 10
    n = fe.kf_idx
    assert (fe.ht, fe.wd) == (60, 80)
    assert torch.isfinite(fe.cam0_T_world[:n]).all() and (fe.cam0_idepths[:n] >= 1e-3).all()
    assert torch.isfinite(fe.cam0_idepths_cov[:n]).all() and (fe.cam0_idepths_cov[:n] > 0).all()
    assert torch.isfinite(fe.cam0_depths_cov_up[:n]).all()
    assert fe.ba_failures(wait=True) == 0
    assert len(fe.ii_h) <= fe.max_factors + 2
    viz = [v for v in outs if v is not None and "cam0_poses" in v]
    assert viz and viz[-1]["cam0_idepths_up"].shape[-2:] == (480, 640) and viz[-1]["cam0_images"].dtype == torch.uint8
    if os.path.exists(WEIGHTS):
        err = _traj_error(fe)
        assert err < 0.15, f"relative trajectory error {err:.3f}"


def test_prefetched_proximity_distances_are_the_distances():
    """the asynchronous prefetch of the next candidate's pairwise distances (frontend._prefetch_proximity) must hand
    add_proximity_factors exactly what the direct computation gives at that point (bit for bit: same kernel, same state)"""
    from nerf_slam_b200.frontend import RaftVisualFrontend
    hits = []
    orig = RaftVisualFrontend._take_prefetched_distances

    def checked(self, kf0, kf1, t, beta):
        d = orig(self, kf0, kf1, t, beta)
        if d is not None:
            ii, jj = np.meshgrid(np.arange(kf0, t), np.arange(kf1, t), indexing="ij")
            direct = self.distance(ii.reshape(-1), jj.reshape(-1), beta=beta).cpu().numpy()
            assert np.array_equal(d, direct)
            hits.append(len(d))
        return d
    RaftVisualFrontend._take_prefetched_distances = checked
    try:
        fe, _, _ = _run(60)
    finally:
        RaftVisualFrontend._take_prefetched_distances = orig
    assert fe.is_initialized and len(hits) >= 3, hits          # every steady-state candidate was served by the prefetch
