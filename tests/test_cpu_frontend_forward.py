"""A19, per-frame control flow of the live class: RaftVisualFrontend.forward + initialize_buffers + get_viz_out + _store_frame
run on the CPU (encoders, motion decision and the keyframe routines replaced by the scenario's stand-ins) must reproduce
the traces recorded by executing the REFERENCE's own forward / initialize_buffers / get_viz_out verbatim
(tests/golden/make_golden_forward.py): which frames become keyframes, kf_idx / last_k bookkeeping, the frame <-> keyframe
maps, when initialisation / update / rm_keyframe / terminate are called, the last-frame and buffer-full rules, the stored
per-keyframe inputs, the output packet (keys, dirty indices, shapes, contents) and the initial values of every buffer."""
import gzip
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import forward_scenario as sc   # noqa: E402

with gzip.open(os.path.join(HERE, "golden", "ref_forward_traces.json.gz"), "rt") as f:
    GOLD = json.load(f)


def _frontend(case, monkeypatch):
    from nerf_slam_b200 import _lib, corr, frontend as fr
    monkeypatch.setattr(_lib, "h2d", lambda a, device, dtype=None: (torch.from_numpy(np.ascontiguousarray(a)) if dtype is None
                                                                    else torch.from_numpy(np.ascontiguousarray(a)).to(dtype)))
    import contextlib
    monkeypatch.setattr(_lib, "fixed_stream", contextlib.nullcontext)
    motion, accept = sc.plan(case["seed"], case["n_frames"], case["last_has_motion"])
    log = []

    class CpuFrontend(fr.RaftVisualFrontend):
        def __init__(self):
            self.device, self.buffer, self.dsf, self.stereo = "cpu", case["buffer"], 8, False
            self.args = types.SimpleNamespace(multi_gpu=False, corr_slots=4)
            self.kf_idx, self.last_kf_idx, self.last_k = 0, 0, None
            self.kf_idx_to_f_idx, self.f_idx_to_kf_idx = {}, {}
            self.is_initialized, self.keyframe_warmup, self.stop, self.global_ba = False, 8, False, False
            self.max_factors, self.motion_filter_thresh = 48, 2.4
            self.cam0_t0_T_world = np.array([0.1, 0.2, 0.3, 0, 0, 0, 1.0])
            self.world_T_body_t0 = np.array([-0.1, -0.2, -0.3, 0, 0, 0, 1.0])
            self.world_T_cam0_t0 = np.array([-0.1, -0.2, -0.3, 0, 0, 0, 1.0])
            self.g_prior_cov = torch.block_diag(0.01 ** 2 * torch.eye(3), 0.01 ** 2 * torch.eye(3))
            self.idepth_prior_cov = 0.1 ** 2
            self.timers = fr._Timers()
            self.stats = {"updates": 0}
            self._img_static = None
            self.use_cuda_graphs = False

        def _normalize_imgs(self, images):
            return images[:, :, :3].float()

        def _frame_front(self, imgs_k):
            k = int(imgs_k[0, 0, 0, 0, 0])
            self._img_static = imgs_k                   # what the real per-frame front leaves for the context encoder
            self.last_motion = torch.tensor(10.0 if motion[k] else 0.0)
            return torch.full((1, 128, self.ht, self.wd), float(k), dtype=torch.half)

        def _feature_encoder(self, imgs_norm):
            return torch.full((1, 128, self.ht, self.wd), float(int(imgs_norm[0, 0, 0, 0, 0])), dtype=torch.half)

        def _context_encoder(self, imgs_norm):
            k = float(int(imgs_norm[0, 0, 0, 0, 0]))
            return (torch.full((1, self.ht, self.wd, 128), k + 0.25, dtype=torch.half),
                    torch.full((1, self.ht, self.wd, 128), k + 0.5, dtype=torch.half))

        def _initialize(self):
            log.append(["initialize", self.kf_idx]); self.is_initialized = True; self.viz_idx[:self.kf_idx + 1] = True

        def _update(self):
            ok = bool(accept[self.kf_idx_to_f_idx[self.kf_idx]])
            log.append(["update", self.kf_idx, ok]); self.viz_idx[max(self.kf_idx - 2, 0):self.kf_idx + 1] = True
            return ok

        def rm_keyframe(self, k):
            log.append(["rm_keyframe", k])

        def terminate(self):
            log.append(["terminate", self.kf_idx]); self.stop = True

    return CpuFrontend(), log


@pytest.mark.parametrize("n", range(len(GOLD)))
def test_forward_replays_the_reference_method(n, monkeypatch):
    case, ref, init = GOLD[n]["case"], GOLD[n]["trace"], GOLD[n]["init"]
    fe, log = _frontend(case, monkeypatch)
    got = []
    for k in range(case["n_frames"]):
        x0, factors, viz = fe.forward(sc.packet(k, case["n_frames"]))
        assert x0 is not None and factors is not None
        got.append({"k": k, "kf_idx": int(fe.kf_idx), "last_k": None if fe.last_k is None else int(fe.last_k),
                    "last_kf_idx": int(fe.last_kf_idx), "is_initialized": bool(fe.is_initialized), "stop": bool(fe.stop),
                    "kf2f": {int(a): int(b) for a, b in fe.kf_idx_to_f_idx.items()},
                    "f2kf": {int(a): int(b) for a, b in fe.f_idx_to_kf_idx.items()},
                    "viz": sc.summarize_viz(viz), "log": list(log),
                    "feat_ids": [int(v) for v in fe.features_imgs[:, 0, 0, 0, 0].tolist()],
                    "ctx_ids": [round(float(v), 2) for v in fe.contexts_imgs[:, 0, 0, 0, 0].tolist()],
                    "tstamps": [float(v) for v in fe.cam0_timestamps.tolist()]})
        log.clear()
        if fe.stop:
            break
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        for key in ("k", "kf_idx", "last_k", "last_kf_idx", "is_initialized", "stop", "log", "feat_ids", "ctx_ids", "tstamps"):
            assert g[key] == r[key], (g["k"], key, g[key], r[key])
        assert {int(a): b for a, b in r["kf2f"].items()} == g["kf2f"] and {int(a): b for a, b in r["f2kf"].items()} == g["f2kf"], g["k"]
        gv, rv = g["viz"], r["viz"]
        assert (gv is None) == (rv is None), g["k"]
        if rv is not None:
            assert set(rv["keys"]) <= set(gv["keys"]), (g["k"], set(rv["keys"]) - set(gv["keys"]))       # ours adds host copies
            assert gv["is_last_frame"] == rv["is_last_frame"]
            if "viz_idx" in rv:
                assert gv["viz_idx"] == rv["viz_idx"] and gv["kf_idx"] == rv["kf_idx"], g["k"]
                assert gv["kf_idx_to_f_idx"] == {int(a): b for a, b in rv["kf_idx_to_f_idx"].items()}
                assert gv["shapes"] == rv["shapes"] and gv["images_sum"] == rv["images_sum"], g["k"]
                assert np.isclose(gv["gt_depth_sum"], rv["gt_depth_sum"]) and np.allclose(gv["intr"], rv["intr"])
                assert np.isclose(gv["poses_sum"], rv["poses_sum"], atol=1e-4)
    # initial values / shapes of the buffers (untouched last slot)
    assert np.isclose(float(fe.cam0_idepths_cov[-1, 0, 0]), init["idepths_cov"]) and float(fe.cam0_depths_cov[-1, 0, 0]) == init["depths_cov"]
    assert float(fe.cam0_idepths[-1, 0, 0]) == init["idepths"] and float(fe.cam0_idepths_up[-1, 0, 0]) == init["idepths_up"]
    assert float(fe.cam0_depths_cov_up[-1, 0, 0]) == init["depths_cov_up"] and np.isclose(float(fe.damping[-1, 0, 0]), init["damping"])
    assert np.allclose(fe.cam0_T_world[-1].tolist(), init["T_world"]) and np.allclose(fe.world_T_body[-1].tolist(), init["wTb"])
    assert np.allclose(torch.diagonal(fe.world_T_body_cov[-1]).tolist(), init["wTb_cov_diag"])
    assert fe.coords0[-1, -1].tolist() == init["coords0_last"]
    for name, shape in init["shapes"].items():
        mine = list(getattr(fe, name).shape)
        if name in ("features_imgs", "contexts_imgs", "cst_contexts_imgs"):       # ours are channels-last
            assert mine == [shape[0], shape[1], shape[3], shape[4], shape[2]], name
        else:
            assert mine == shape, name
