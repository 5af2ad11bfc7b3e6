"""Drop-in check: the REFERENCE's own entry point examples/slam_demo.py, loaded UNCHANGED from
/root/reference, runs against this repo when `nerf_slam_b200/shim` is first on sys.path — argument parsing, module
construction, queue wiring, the sequential spin loop, shutdown — with the two GPU workers (RaftVisualFrontend,
NerfFusion) replaced by recording stand-ins (no GPU in this container; tests/test_gpu_demo.py runs the same wiring with the
real workers on hardware).  Skipped where /root/reference does not exist (the GPU box)."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

REF_DEMO = "/root/reference/examples/slam_demo.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "nerf_slam_b200", "shim")


@pytest.mark.skipif(not os.path.exists(REF_DEMO), reason="reference tree not present")
def test_reference_slam_demo_runs_unchanged_on_the_shim(tmp_path, monkeypatch):
    from nerf_slam_b200 import datasets, frontend, nerf_fusion, synthetic
    room = synthetic.SyntheticRoom(64, 48, 7)
    datasets.write_transforms_dataset(room, str(tmp_path))
    log = {"frames": [], "fused": 0, "fit_only": 0, "ctor": None}

    class FakeFrontend:
        def __init__(self, world_T_body_t0, body_T_cam0, args, device="cuda:0"):
            log["ctor"] = (np.asarray(world_T_body_t0).shape, np.asarray(body_T_cam0).shape, device, args.buffer)
            self.n = 0

        def __call__(self, batch):
            self.n += 1
            log["frames"].append((int(batch["k"][0]), batch["images"].shape, bool(batch["is_last_frame"])))
            return frontend.EmptyValues(), frontend.EmptyFactorGraph(), {"kf_idx": self.n, "is_last_frame": batch["is_last_frame"]}

        def stop_condition(self):
            return False

    class FakeFusion:
        def __init__(self, name, args, device):
            assert name == "nerf" and device == "cuda:0"

        def fuse(self, packets):
            if packets:
                assert set(packets) == {"slam"} and packets["slam"][1]["kf_idx"] >= 1       # [state, viz_out] (meta_slam.py:47)
                log["fused"] += 1
            else:
                log["fit_only"] += 1
            return True

        def stop_condition(self):
            return log["fit_only"] >= 3
    monkeypatch.setattr(frontend, "RaftVisualFrontend", FakeFrontend)
    monkeypatch.setattr(nerf_fusion, "NerfFusion", FakeFusion)
    monkeypatch.setattr(sys, "path", [SHIM, ROOT] + [p for p in sys.path if p not in (SHIM, ROOT)])
    for m in [k for k in sys.modules if k.split(".")[0] in ("datasets", "slam", "fusion", "pipeline", "gui", "icecream")]:
        monkeypatch.delitem(sys.modules, m)
    monkeypatch.setattr(sys, "argv", ["slam_demo.py", f"--dataset_dir={tmp_path}", "--dataset_name=nerf", "--buffer=50",
                                      "--slam", "--fusion=nerf"])
    spec = importlib.util.spec_from_file_location("reference_slam_demo", REF_DEMO)
    demo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(demo)                          # the reference's file, byte for byte
    args = demo.parse_args()
    assert args.buffer == 50 and args.fusion == "nerf" and args.slam and not args.parallel_run
    demo.run(args)
    # the reference's reader drops the last frame with the default --final_k=-1 (datasets/nerf_dataset.py:65): 6 of 7
    assert [f[0] for f in log["frames"]] == list(range(6))
    assert log["frames"][0][1] == (1, 48, 64, 4)
    assert log["ctor"] == ((4, 4), (4, 4), "cuda:0", 50)
    assert log["fused"] == 6 and log["fit_only"] >= 3      # every SLAM output reached the fusion module; then it kept fitting
    assert "slam.slam_module" in sys.modules and sys.modules["slam.slam_module"].__file__.startswith(SHIM)


def test_pipeline_modules_replay_the_reference_modules():
    """DataModule / SlamModule / FusionModule (+ the MIMO base) of nerf_slam_b200/pipeline.py against traces of the
    reference's OWN classes under the same scripted scenarios (tests/golden/ref_pipeline_traces.json, recorded by
    make_golden_pipeline.py from /root/reference with only `colored_glog` / `icecream` stubbed): sequential spin loops incl. a
    falsy and a None SLAM output and the stop condition, parallel loops until self-shutdown, shutdown / restart, a raising
    consumer, unknown module names, empty inputs"""
    import json
    from nerf_slam_b200 import pipeline
    from tests.golden import pipeline_scenario as sc
    with open(os.path.join(ROOT, "tests", "golden", "ref_pipeline_traces.json")) as f:
        ref = json.load(f)
    got = json.loads(json.dumps(sc.run((pipeline.DataModule, pipeline.SlamModule, pipeline.FusionModule))))
    assert set(got) == set(ref)
    for name in ref:
        assert got[name] == ref[name], name
