"""Record the keyframe loop of the reference's LIVE front end by executing its own methods verbatim — build container only.

  python tests/golden/make_golden_live_frontend.py        ->  tests/golden/ref_live_frontend_traces.json.gz

slam/visual_frontends/visual_frontend.py cannot be imported (gtsam, lietorch, droid_backends, ...).  Its class body is
parsed with `ast` instead, and the methods that make up the keyframe logic — rm_keyframe, __update, __initialize,
add_neighborhood_factors, add_proximity_factors, add_factors, rm_factors, __filter_repeated_edges, backend,
clear_edges, normalize — are compiled from the
reference's source text into a class of the same name (so that private names mangle identically).  Nothing of it is
copied into this repository.  Stand-ins, as in live_frontend_scenario.py: update(), distance(), reproject(); CorrBlock
is the reference's own class.  Unstable `torch.argsort` calls are pinned to index order for ties (see
make_golden_factor_graph.py)."""
import ast
import gzip
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = ["rm_keyframe", "__update", "__initialize", "add_neighborhood_factors", "add_proximity_factors", "add_factors",
           "rm_factors", "__filter_repeated_edges", "backend", "clear_edges", "normalize", "update"]


def reference_class():
    src = open(os.path.join(REF, "slam", "visual_frontends", "visual_frontend.py")).read()
    tree = ast.parse(src)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RaftVisualFrontend"][0]
    keep = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in METHODS]
    assert sorted(n.name for n in keep) == sorted(METHODS)
    new = ast.Module(body=[ast.ClassDef(name="RaftVisualFrontend", bases=[], keywords=[], body=keep, decorator_list=[])],
                     type_ignores=[])
    ast.fix_missing_locations(new)
    for name, m in (("droid_backends", types.ModuleType("droid_backends")),):
        sys.modules.setdefault(name, m)
    sys.path.insert(0, REF)
    from networks.modules.corr import CorrBlock
    import droid_backends as dbk                      # the lookup kernel: zeros of its real shape (src/droid.cpp:280-288)
    dbk.corr_index_forward = lambda volume, coords, radius: (torch.zeros(volume.shape[0], 2 * radius + 1, 2 * radius + 1, *coords.shape[-2:]),)
    sys.modules.setdefault("icecream", types.ModuleType("icecream")).ic = lambda *a, **k: None
    from utils.flow_viz import cvx_upsample             # the reference's own (pure torch)
    ns = {"torch": torch, "np": np, "ic": lambda *a, **k: None, "CorrBlock": CorrBlock, "cvx_upsample": cvx_upsample}
    exec(compile(new, "visual_frontend.py (reference, selected methods)", "exec"), ns)
    return ns["RaftVisualFrontend"]


def make_shim(Ref, sc, D, feats, ctx, buffer, keyframe_thresh=4.0):
    P = sc.PARAMS

    class Shim(Ref):
        def __init__(self):
            self.device, self.stereo, self.corr_impl = "cpu", False, "volume"
            for k, v in P.items():
                setattr(self, k, v)
            self.keyframe_thresh = keyframe_thresh
            self.buffer, self.kf_idx, self.is_initialized = buffer, 0, False
            B, h, w, C = buffer, sc.HT8, sc.WD8, sc.CH
            z = torch.zeros
            self.gt_poses, self.gt_depths, self.cam0_images = z(B, 4, 4), z(B, 1, 2, 2), z(B, 3, 2, 2)
            self.cam0_timestamps, self.cam0_T_world, self.world_T_body = z(B), z(B, 7), z(B, 7)
            self.world_T_body_cov, self.cam0_idepths = z(B, 6, 6), torch.ones(B, h, w)
            self.cam0_idepths_cov, self.cam0_depths_cov, self.cam0_idepths_sensed = torch.ones(B, h, w), torch.ones(B, h, w), z(B, h, w)
            self.cam0_intrinsics = z(B, 4)
            self.features_imgs, self.contexts_imgs, self.cst_contexts_imgs = z(B, 1, C, h, w), z(B, 1, C, h, w), z(B, 1, C, h, w)
            self.viz_idx = z(B, dtype=torch.bool)
            L = lambda: torch.as_tensor([], dtype=torch.long)
            self.ii, self.jj, self.age = L(), L(), L()
            self.ii_inactive, self.jj_inactive, self.ii_bad, self.jj_bad = L(), L(), L(), L()
            self.correlation_volumes = self.gru_hidden_states = self.gru_contexts_input = None
            e = lambda: z(1, 0, h, w, 2)
            self.gru_estimated_flow, self.gru_estimated_flow_weight = e(), e()
            self.gru_estimated_flow_inactive, self.gru_estimated_flow_weight_inactive = e(), e()
            self.coords0 = sc.coords0()
            self.ht, self.wd, self.lowmem_log, self.ba_log = h, w, [], []
            self.damping = 1e-6 * torch.ones_like(self.cam0_idepths)
            self.cam0_T_body = torch.tensor([0, 0, 0, 0, 0, 0, 1.0])
            self.compute_covariances, self.viz = False, False
            self.cam0_idepths_up, self.cam0_depths_cov_up = z(B, 8 * h, 8 * w), z(B, 8 * h, 8 * w)
            sys.path.insert(0, HERE)
            import factor_graph_scenario as fgs
            self.update_net = fgs.fake_update_net

        def _ids(self, ix):
            return self.cam0_intrinsics[torch.as_tensor(ix).long().reshape(-1), 0].long()

        def distance(self, ii=None, jj=None, beta=0.3, bidirectional=True):
            return D[self._ids(ii), self._ids(jj)].clone()

        def reproject(self, ii, jj, cam_T_body=None, jacobian=False):
            off = (self._ids(ii) * 100 + self._ids(jj)).float() + 10.0 * (self.cam0_idepths[ii, 0, 0] - 1.0)
            c = self.coords0[None, None] + off.view(1, -1, 1, 1, 1)
            return c, torch.ones_like(c[..., :1]), (None, None, None)

        real_update = Ref.update                         # the reference's own update() (used by the update scenario)

        def update(self, kf0=None, kf1=None, itrs=2, use_inactive=False, EP=1e-7, motion_only=False):
            self.age += 1
            self.cam0_idepths[:self.kf_idx + 1] *= 1.001
            return None, None

        def ba(self, flow, weight, damping, ii, jj, kf0=0, kf1=None, itrs=2, lm=1e-4, ep=0.1, motion_only=False,
               compute_covariances=True):
            self.ba_log.append({"ii": [int(v) for v in ii.tolist()], "jj": [int(v) for v in jj.tolist()], "kf0": int(kf0),
                                "kf1": None if kf1 is None else int(kf1), "itrs": int(itrs), "motion_only": bool(motion_only),
                                "target": sc.digest(flow), "weight": sc.digest(weight), "damping": sc.digest(damping),
                                "contig": bool(flow.is_contiguous() and weight.is_contiguous() and damping.is_contiguous())})
            self.cam0_idepths[torch.unique(ii)] *= 1.01
            return None, None

        def update_lowmem(self, itrs=2, EP=1e-7, steps=8):
            self.lowmem_log.append([[int(v) for v in self.ii.tolist()], [int(v) for v in self.jj.tolist()], int(steps)])

    fe = Shim()

    class Acc:
        kf_idx = property(lambda s: fe.kf_idx, lambda s, v: setattr(fe, "kf_idx", v))
        is_initialized = property(lambda s: fe.is_initialized)

        def backend(self, steps):
            fe.lowmem_log = []
            fe.backend(steps)
            return fe.lowmem_log

        # ---- update scenario
        def add_neighborhood(self, kf0, kf1, radius):
            fe.add_neighborhood_factors(kf0, kf1, radius)

        def retire(self, first):
            fe.rm_factors(fe.ii < first, store=True)

        def live_update(self, use_inactive):
            Shim.real_update(fe, kf0=None, kf1=None, itrs=2, use_inactive=use_inactive)

        def snapshot_update(self):
            d = {"ii": [int(v) for v in fe.ii.tolist()], "age": [int(v) for v in fe.age.tolist()],
                 "viz": [int(v) for v in fe.viz_idx.tolist()],
                 "flow": sc.digest(fe.gru_estimated_flow), "weight": sc.digest(fe.gru_estimated_flow_weight),
                 "hidden": sc.digest(fe.gru_hidden_states), "damping00": [round(float(v), 7) for v in fe.damping[:, 0, 0].tolist()],
                 "idepth00": [round(float(v), 6) for v in fe.cam0_idepths[:, 0, 0].tolist()], "ba": fe.ba_log}
            fe.ba_log = []
            return d

        def put_frame(self, slot, fid):
            fe.cam0_intrinsics[slot, 0] = float(fid)
            fe.cam0_T_world[slot, 0] = 0.1 * fid
            fe.features_imgs[slot, 0] = feats[fid]; fe.contexts_imgs[slot, 0] = ctx[fid]; fe.cst_contexts_imgs[slot, 0] = -ctx[fid]

        def initialize(self):
            fe._RaftVisualFrontend__initialize()

        def update(self):
            return fe._RaftVisualFrontend__update()

        def rm_keyframe(self, k):
            fe.rm_keyframe(k)

        def snapshot(self):
            tl = lambda t: [int(v) for v in t.tolist()]
            hid = fe.gru_hidden_states
            return {"kf_idx": int(fe.kf_idx), "is_initialized": bool(fe.is_initialized),
                    "ii": tl(fe.ii), "jj": tl(fe.jj), "age": tl(fe.age), "ii_inac": tl(fe.ii_inactive), "jj_inac": tl(fe.jj_inactive),
                    "ids": tl(fe.cam0_intrinsics[:, 0]), "viz": tl(fe.viz_idx),
                    "idepth00": [round(float(v), 6) for v in fe.cam0_idepths[:, 0, 0].tolist()],
                    "tx": [round(float(v), 6) for v in fe.cam0_T_world[:, 0].tolist()], "max_factors": int(fe.max_factors),
                    "flow00": [float(v) for v in fe.gru_estimated_flow[0, :, 0, 0, 0].tolist()],
                    "flow_inac00": [float(v) for v in fe.gru_estimated_flow_inactive[0, :, 0, 0, 0].tolist()],
                    "hidden00": [] if hid is None else [round(float(v), 5) for v in hid[0, :, 0, 0, 0].tolist()],
                    "n_volumes": 0 if fe.correlation_volumes is None else int(fe.correlation_volumes.corr_pyramid[0].shape[0])}
    return Acc()


def main():
    warnings.filterwarnings("ignore")
    sys.path.insert(0, HERE)
    import live_frontend_scenario as sc
    Ref = reference_class()
    unstable = torch.argsort
    torch.argsort = lambda x, *a, **k: unstable(x, *a, **{**k, "stable": True})
    out = []
    for case in [dict(seed=31, n_steps=30, slope=1.0), dict(seed=32, n_steps=36, slope=0.4), dict(seed=33, n_steps=30, slope=2.0),
                 dict(seed=34, n_steps=44, slope=0.7), dict(seed=35, n_steps=40, slope=1.0, keyframe_thresh=9.0)]:
        D, feats, ctx = sc.bank(case["seed"], case["n_steps"] + 4, case["slope"])
        acc = make_shim(Ref, sc, D, feats, ctx, buffer=case["n_steps"] + 6, keyframe_thresh=case.get("keyframe_thresh", 4.0))
        trace = sc.run(acc, **case)
        out.append({"case": case, "trace": trace})
        print(case, "keyframes", trace[-1]["kf_idx"], "rejected", sum(not t["accepted"] for t in trace),
              "max edges", max(len(t["ii"]) for t in trace), "inactive", len(trace[-1]["ii_inac"]))
    upd = []
    for case in [dict(seed=41, n_kf=10), dict(seed=42, n_kf=13)]:
        D, feats, ctx = sc.bank(case["seed"], case["n_kf"] + 4, 1.0)
        acc = make_shim(Ref, sc, D, feats, ctx, buffer=case["n_kf"] + 4)
        trace = sc.run_update(acc, **case)
        upd.append({"case": case, "trace": trace})
        print("update()", case, "snapshots", len(trace), "BA edges", [len(t["ba"][0]["ii"]) for t in trace])
    torch.argsort = unstable
    path = os.path.join(HERE, "ref_live_frontend_traces.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as gz, __import__("io").TextIOWrapper(gz) as f:   # mtime=0: reproducible bytes
        json.dump({"loops": out, "updates": upd}, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
