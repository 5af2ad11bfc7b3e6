"""Golden for NerfFusion's control loop — fuse / fit_volume / fit_volume_once / stop_condition (fusion/nerf_fusion.py:237-307),
executed VERBATIM (cut out of the reference file with `ast`) around a recording stand-in for `ngp` and scripted
process_data / process_slam results.

    python tests/golden/make_golden_nerf_fusion_loop.py    (needs /root/reference; writes ref_nerf_fusion_loop.json)"""
import ast
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/fusion/nerf_fusion.py"
METHODS = ("fuse", "fit_volume", "fit_volume_once", "stop_condition")


class FakeNgp:
    def __init__(self):
        self.frames = 0
        self.elapsed_training_time = 0.0
        self.loss = 0.0
        self.rgba = object()
        self.nerf = types.SimpleNamespace(training=types.SimpleNamespace(depth_supervision_lambda=1.0))

    def frame(self):
        self.frames += 1
        self.elapsed_training_time += 0.01

    def apply_camera_smoothing(self, dt):
        pass


def drive(obj, evaluate, anneal):
    """the same scripted sequence of fuse() inputs for the reference's methods and for ours"""
    obj.ngp = FakeNgp()
    obj.iters, obj.total_iters, obj.stop_iters = 3, 0, 14
    obj.anneal, obj.anneal_every_iters, obj.annealing_rate = anneal, 4, 0.5
    obj.evaluate, obj.eval_every_iters = evaluate, 5
    log = {"evals": [], "calls": [], "stops": [], "raised": None}
    obj.eval_gt_traj = lambda *a, **k: log["evals"].append(obj.total_iters)
    script = iter([False, True, False, True, True])
    obj.process_slam = lambda p: (log["calls"].append("slam"), next(script))[1]
    obj.process_data = lambda p: (log["calls"].append("data"), next(script))[1]
    for packets in ({"slam": 1}, None, {"data": 1}, {}, {"slam": 1, "data": 2}, False, {"slam": 3}, None):
        r = obj.fuse(packets)
        log["stops"].append([bool(r), int(obj.total_iters), int(obj.ngp.frames), bool(obj.stop_condition()),
                             round(obj.ngp.nerf.training.depth_supervision_lambda, 6)])
    try:
        obj.fuse({"gui": 1})
    except NotImplementedError:
        log["raised"] = "NotImplementedError"
    return log


def reference_class():
    src = open(REF).read()
    cls = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "NerfFusion")
    body = "\n".join(ast.get_source_segment(src, f) for f in cls.body if isinstance(f, ast.FunctionDef) and f.name in METHODS)
    code = "class NerfFusion:\n" + "\n".join("    " + l if not l.startswith("    ") else l for l in body.split("\n"))
    ns = {"print": lambda *a, **k: None}
    exec(compile(code, REF, "exec"), ns)
    return ns["NerfFusion"]


def main():
    Ref = reference_class()
    out = {f"eval{int(e)}_anneal{int(a)}": drive(Ref.__new__(Ref), e, a) for e in (False, True) for a in (False, True)}
    with open(os.path.join(HERE, "ref_nerf_fusion_loop.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: v["stops"][-1] for k, v in out.items()})


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    main()
