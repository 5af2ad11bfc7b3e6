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


def _run(n_frames, W=320, H=240, buffer=24, step=0.03, conv_backend="tcgen05"):
    from nerf_slam_b200.frontend import RaftVisualFrontend
    from nerf_slam_b200.synthetic import SyntheticRoom
    args = types.SimpleNamespace(buffer=buffer, stereo=False, multi_gpu=False, conv_backend=conv_backend,
                                 weights=WEIGHTS if os.path.exists(WEIGHTS) else None)
    room = SyntheticRoom(W, H, n_frames, seed=0, step=step)
    fe = RaftVisualFrontend(np.linalg.inv(room.packet(0)["poses"][0]), np.eye(4), args, "cuda:0")
    outs = []
    for k in range(n_frames):
        x0, f, viz = fe.forward(room.packet(k))
        outs.append(viz)
        if fe.stop_condition():
            break
    return fe, room, outs


@pytest.mark.parametrize("conv_backend", ["tcgen05", "cudnn"])
def test_frontend_runs_and_tracks(conv_backend):
    fe, room, outs = _run(60, conv_backend=conv_backend)
    torch.cuda.synchronize()
    assert fe.is_initialized, f"not initialised after 60 frames (kf_idx={fe.kf_idx})"
    n = fe.kf_idx
    assert torch.isfinite(fe.cam0_T_world[:n]).all() and torch.isfinite(fe.cam0_idepths[:n]).all()
    assert (fe.cam0_idepths[:n] >= 1e-3).all()
    assert len(fe.ii_h) <= fe.max_factors + 2          # graph size contract
    assert len(set(zip(fe.ii_h.tolist(), fe.jj_h.tolist()))) == len(fe.ii_h)   # no duplicate edges
    assert fe.corr_pool.capacity - len(fe.corr_pool.free) == len(fe.ii_h)       # arena accounting
    q = fe.cam0_T_world[:n, 3:]
    assert torch.allclose(q.norm(dim=-1), torch.ones(n, device=q.device), atol=1e-4)
    viz = [v for v in outs if v is not None and "cam0_poses" in v]
    assert viz and viz[-1]["cam0_idepths_up"].shape[-2:] == (240, 320)
    if os.path.exists(WEIGHTS):
        # with the trained weights the monocular trajectory must match GT up to scale
        est = fe.cam0_T_world[:n].cpu().numpy().astype(np.float64)
        gt = fe.gt_poses[:n].cpu().numpy().astype(np.float64)          # w2c
        from oracle import se3
        c_est = np.stack([se3.inv_se3(e[:3], e[3:])[0] for e in est])   # camera centres in world
        c_gt = np.stack([np.linalg.inv(g)[:3, 3] for g in gt])
        d_est = np.linalg.norm(c_est[1:] - c_est[0], axis=-1)
        d_gt = np.linalg.norm(c_gt[1:] - c_gt[0], axis=-1)
        s = (d_est * d_gt).sum() / (d_est ** 2).sum()
        err = np.abs(s * d_est - d_gt).max() / d_gt.max()
        assert err < 0.15, f"relative trajectory error {err:.3f}"
