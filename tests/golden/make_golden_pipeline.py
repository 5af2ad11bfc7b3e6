"""Golden traces of the reference's OWN pipeline modules (pipeline/pipeline_module.py, datasets/data_module.py,
slam/slam_module.py, fusion/fusion_module.py — pure Python; only `colored_glog` and `icecream` are stubbed) under the
scripted scenarios of pipeline_scenario.py.

    python tests/golden/make_golden_pipeline.py        (needs /root/reference; writes tests/golden/ref_pipeline_traces.json)"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def reference_classes():
    log = types.ModuleType("colored_glog")
    for n in ("info", "debug", "warn", "error", "check", "log", "check_eq", "check_lt", "check_gt", "check_le", "check_ge", "check_ne"):
        setattr(log, n, lambda *a, **k: None)
    ice = types.ModuleType("icecream"); ice.ic = lambda *a, **k: None
    sys.modules["colored_glog"], sys.modules["icecream"] = log, ice
    for m in [k for k in sys.modules if k.split(".")[0] in ("datasets", "slam", "fusion", "pipeline")]:
        del sys.modules[m]
    sys.path.insert(0, REF)
    try:
        from datasets.data_module import DataModule
        from slam.slam_module import SlamModule
        from fusion.fusion_module import FusionModule
    finally:
        sys.path.remove(REF)
    assert DataModule.__module__ == "datasets.data_module" and "/root/reference" in sys.modules["pipeline.pipeline_module"].__file__
    return DataModule, SlamModule, FusionModule


def main():
    sys.path.insert(0, ROOT)
    from tests.golden import pipeline_scenario as sc
    classes = reference_classes()
    out = sc.run(classes)
    with open(os.path.join(HERE, "ref_pipeline_traces.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: (len(v["steps"]) if "steps" in v else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
