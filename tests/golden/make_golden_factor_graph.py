"""Record golden edge-set traces from the REFERENCE's own `FactorGraph` (networks/factor_graph.py), imported
from /root/reference and run on CPU — build container only (the reference tree does not travel).

  python tests/golden/make_golden_factor_graph.py        ->  tests/golden/ref_factor_graph_traces.json.gz

Third-party imports of the reference that cannot be installed (lietorch, matplotlib, droid_backends,
icecream) are stubbed as far as import resolution needs; none of them is executed by the graph-management
methods.  `CorrBlock` is the reference's own (pure torch on CPU: matmul + avg_pool, corr.py:23-38,52-72).
The driver is `factor_graph_scenario.run_scenario` (shared with the replay test).

Ties: the reference orders candidate edges with `torch.argsort(d)` and retires edges with `torch.argsort(age)`
(networks/factor_graph.py:366,107) — UNSTABLE sorts, so the order of equal keys is implementation-defined (torch's
CPU sort permutes ties once n > 16; the CUDA sort has its own order) and equal ages are the normal case.  The
recording pins the one reproducible choice, index order (`stable=True`), by wrapping `torch.argsort` for the
duration of the run; this repo's implementation makes the same choice (numpy stable argsort).
"""
import json
import os
import sys
import types
import warnings

import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
SCENARIOS = [dict(seed=s, n_frames=n, stereo=st, slope=sl, max_factors=mf) for s, n, st, sl, mf in
             [(0, 30, False, 4.0, 48), (1, 34, False, 1.5, 48), (2, 26, True, 2.0, 48), (3, 40, False, 1.0, 48),
              (4, 22, False, 1.5, 24), (5, 30, True, 1.0, 32)]]


def _stub_modules():
    lt = types.ModuleType("lietorch"); lt.SE3 = object; lt.Sim3 = object
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot"); mpl.pyplot = plt
    dbk = types.ModuleType("droid_backends")
    ic = types.ModuleType("icecream"); ic.ic = lambda *a, **k: None
    for name, m in (("lietorch", lt), ("matplotlib", mpl), ("matplotlib.pyplot", plt), ("droid_backends", dbk),
                    ("icecream", ic)):
        sys.modules.setdefault(name, m)


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    sys.path.insert(0, HERE)
    warnings.filterwarnings("ignore")
    from networks.factor_graph import FactorGraph
    from factor_graph_scenario import run_scenario
    out = []
    unstable_argsort = torch.argsort
    torch.argsort = lambda x, *a, **k: unstable_argsort(x, *a, **{**k, "stable": True})
    for sc in SCENARIOS:
        make = lambda video, mf, net=None: FactorGraph(video, net, device="cpu", corr_impl="volume", max_factors=mf)
        trace = run_scenario(make, **sc)
        out.append({"scenario": sc, "trace": trace})
        last = trace[-1]
        print(sc, "snapshots", len(trace), "max edges", max(len(t["ii"]) for t in trace), "final edges", len(last["ii"]), "inactive", len(last["ii_inac"]), "bad", len(last["ii_bad"]))
    torch.argsort = unstable_argsort
    # update() / update_lowmem() with stand-ins for the operator, the BA and the two correlation kernels (zeros of
    # the real shapes: src/droid.cpp:280-288 corr_index_forward -> [n,7,7,h,w]; :303-313 altcorr_forward -> [b,n,49,h,w])
    import droid_backends as dbk
    dbk.corr_index_forward = lambda volume, coords, radius: (torch.zeros(volume.shape[0], 2 * radius + 1, 2 * radius + 1, *coords.shape[-2:]),)
    dbk.altcorr_forward = lambda f1, f2, coords, radius: (torch.zeros(coords.shape[0], coords.shape[1], (2 * radius + 1) ** 2, *coords.shape[2:4]),)
    from factor_graph_scenario import run_update_scenario
    upd = []
    for sc in [dict(seed=11, n_kf=9, stereo=False), dict(seed=12, n_kf=12, stereo=False), dict(seed=13, n_kf=10, stereo=True)]:
        make = lambda video, mf, net=None: FactorGraph(video, net, device="cpu", corr_impl="volume", max_factors=mf)
        trace = run_update_scenario(make, **sc)
        upd.append({"scenario": sc, "trace": trace})
        print("update scenario", sc, "snapshots", len(trace), "BA calls", sum(len(t["ba_calls"]) for t in trace))
    # DroidFrontend (networks/droid_frontend.py) on the same stand-ins; its FactorGraph is created on the CPU
    # (the class hard-codes the default device "cuda:0") and lietorch's SE3 — used for one unused local — is a no-op
    import lietorch
    lietorch.SE3 = lambda data: data
    import networks.droid_frontend as rdf
    rdf.SE3 = lietorch.SE3
    rdf.FactorGraph = lambda video, net, max_factors=-1: FactorGraph(video, net, device="cpu", max_factors=max_factors)
    from factor_graph_scenario import run_frontend_scenario
    torch.argsort = lambda x, *a, **k: unstable_argsort(x, *a, **{**k, "stable": True})
    fr = []
    for sc in [dict(seed=21, n_steps=22, slope=1.0), dict(seed=22, n_steps=26, slope=0.5, keyframe_thresh=8.0),
               dict(seed=23, n_steps=24, slope=2.0, keyframe_thresh=10.0)]:
        trace = run_frontend_scenario(lambda net, video, args: rdf.DroidFrontend(net, video, args), **sc)
        fr.append({"scenario": sc, "trace": trace})
        print("frontend scenario", sc, "final t1", trace[-1]["t1"], "counter", trace[-1]["counter"],
              "dropped keyframes", sum(1 for a, b in zip(trace, trace[1:]) if b["counter"] == a["counter"]))
    torch.argsort = unstable_argsort
    path = os.path.join(HERE, "ref_factor_graph_traces.json.gz")
    import gzip
    with gzip.GzipFile(path, "wb", mtime=0) as gz, __import__("io").TextIOWrapper(gz) as f:   # mtime=0: reproducible bytes
        json.dump({"torch": torch.__version__, "scenarios": out, "update_scenarios": upd, "frontend_scenarios": fr}, f,
                  separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
