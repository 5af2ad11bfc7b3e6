"""A deterministic factor-graph life cycle, written against the PUBLIC surface of `FactorGraph`
(networks/factor_graph.py) only, so that the same driver runs the reference's class (to record the
golden trace, `make_golden_factor_graph.py`) and this repo's class (to replay it, `tests/test_cpu_droid.py`).

The sequence is the one `DroidFrontend` issues (networks/droid_frontend.py:35-121): neighbourhood
factors + 8 updates + proximity factors + 8 updates at initialisation, then per keyframe: age-based
retirement, proximity factors with `remove=True`, 4 (+2) updates, and — for some frames — `rm_keyframe`.
`update()` itself (network + BA) is replaced by its bookkeeping side effects (age += 1, new confidences),
because the trace pins the EDGE SET, which is the bit-exact part of the contract.
"""
import contextlib
import types

import numpy as np
import torch

HT8, WD8, CH = 16, 16, 4       # 1/8-resolution grid (4 pooling steps need >= 16) and channels of the fake video


class FakeVideo:
    """the attributes `FactorGraph` touches (networks/factor_graph.py:20-29,122-135,173-182,313,326,334)"""

    def __init__(self, n, seed, stereo=False, slope=4.0):
        g = torch.Generator().manual_seed(seed)
        self.ht, self.wd = HT8 * 8, WD8 * 8
        self.stereo = stereo
        self.counter = types.SimpleNamespace(value=0)
        self.ready = types.SimpleNamespace(value=0)
        self.images = torch.zeros(n, 3, 2, 2)
        self.tstamp = torch.arange(n).float()
        self.device = "cpu"
        self.poses = torch.zeros(n, 7)
        self.disps = torch.ones(n, HT8, WD8)
        self.disps_sens = torch.zeros(n, HT8, WD8)
        self.intrinsics = torch.ones(n, 4)
        self.intrinsics[:, 0] = torch.arange(n).float()     # slot -> frame identity (moves with rm_keyframe)
        self.dirty = torch.zeros(n, dtype=torch.bool)
        rig = 2 if stereo else 1
        self.fmaps = torch.randn(n, rig, CH, HT8, WD8, generator=g)
        self.nets = torch.randn(n, CH, HT8, WD8, generator=g)
        self.inps = torch.randn(n, CH, HT8, WD8, generator=g)
        # pairwise "flow distance" of frame identities: grows with the index gap, plus seeded noise and ties
        a = torch.arange(n).float()
        base = slope * (a[:, None] - a[None, :]).abs()
        noise = torch.rand(n, n, generator=g) * 14.0
        d = base + 0.5 * (noise + noise.t())
        d[torch.rand(n, n, generator=g) < 0.08] = 7.5        # exact ties
        d[torch.rand(n, n, generator=g) < 0.03] = 150.0      # > 100 -> treated as inf
        self.D = d.float()
        y, x = torch.meshgrid(torch.arange(HT8).float(), torch.arange(WD8).float(), indexing="ij")
        self.coords0 = torch.stack([x, y], dim=-1)

    def get_lock(self):
        return contextlib.nullcontext()

    def _ids(self, ix):
        return self.intrinsics[torch.as_tensor(ix).long().reshape(-1), 0].long()

    def reproject(self, ii, jj):
        """coords that encode which frames the edge connects: the trace checks that every per-edge tensor
        follows its edge through add / remove / shift"""
        ii = torch.as_tensor(ii).long().reshape(-1); jj = torch.as_tensor(jj).long().reshape(-1)
        off = (self._ids(ii) * 100 + self._ids(jj)).float() + (self.poses[ii, 1] - self.poses[jj, 1]) + self.disps[ii, 0, 0] - 1.0
        c = self.coords0[None, None] + off.view(1, -1, 1, 1, 1)
        return c, torch.ones_like(c[..., :1])

    def distance(self, ii, jj, beta=0.3, bidirectional=True):
        return self.D[self._ids(ii), self._ids(jj)].clone()

    # ---- used by the update() scenario only
    ba_log = None

    def ba(self, target, weight, eta, ii, jj, t0=1, t1=None, itrs=2, lm=1e-4, ep=0.1, motion_only=False):
        """records what FactorGraph hands to the dense BA and applies a deterministic stand-in for its effect"""
        ii = torch.as_tensor(np.asarray(ii.cpu() if torch.is_tensor(ii) else ii)).long().reshape(-1)
        jj = torch.as_tensor(np.asarray(jj.cpu() if torch.is_tensor(jj) else jj)).long().reshape(-1)
        if self.ba_log is None:
            self.ba_log = []
        f = lambda t: [list(t.shape), round(float(t.double().sum()), 4), round(float(t.reshape(-1)[0]), 5), bool(t.is_contiguous())]
        self.ba_log.append({"ii": ii.tolist(), "jj": jj.tolist(), "t0": int(t0), "t1": None if t1 is None else int(t1),
                            "itrs": int(itrs), "lm": float(lm), "ep": float(ep), "motion_only": bool(motion_only),
                            "target": f(target), "weight": f(weight), "eta": f(eta)})
        end = int(max(ii.max(), jj.max())) + 1 if t1 is None else int(t1)
        self.disps[torch.unique(ii)] *= 1.01
        self.poses[int(t0):end, 1] += 0.125

    def upsample(self, ix, mask):
        pass


def snapshot(graph, tag):
    tolist = lambda t: [int(v) for v in torch.as_tensor(t).reshape(-1).tolist()]
    flow = graph.gru_estimated_flow
    wgt = graph.gru_estimated_flow_weight
    hid = graph.gru_hidden_states
    return {
        "tag": tag,
        "ii": tolist(graph.ii), "jj": tolist(graph.jj), "age": tolist(graph.age),
        "ii_inac": tolist(graph.ii_inac), "jj_inac": tolist(graph.jj_inac),
        "ii_bad": tolist(graph.ii_bad), "jj_bad": tolist(graph.jj_bad),
        # per-edge tensors, reduced to one scalar per edge
        "flow00": [float(v) for v in flow[0, :, 0, 0, 0].tolist()],
        "weight00": [round(float(v), 6) for v in wgt[0, :, 0, 0, 0].tolist()],
        "target_inac00": [float(v) for v in graph.target_inac[0, :, 0, 0, 0].tolist()],
        "n_hidden": 0 if hid is None else int(hid.shape[1]),
        "hidden00": [] if hid is None else [round(float(v), 5) for v in hid[0, :, 0, 0, 0].float().tolist()],
    }


def fake_update(graph, rng):
    """the bookkeeping side effects of FactorGraph.update (networks/factor_graph.py:202-255)"""
    E = int(graph.ii.shape[0])
    w = torch.as_tensor(rng.random(E).astype(np.float32))
    w[torch.as_tensor(rng.random(E) < 0.08)] = 1e-4         # low-confidence edges (filter_edges removes |i-j|>2 ones)
    graph.gru_estimated_flow_weight = w.view(1, E, 1, 1, 1).expand(1, E, HT8, WD8, 2).contiguous().to(graph.gru_estimated_flow_weight.device)
    graph.age += 1


def run_scenario(make_graph, seed, n_frames=30, stereo=False, slope=4.0, max_factors=48, warmup=8, max_age=25, window=25,
                 nms=1, radius=2, thresh=16.0, beta=0.3):
    """-> list of snapshots.  make_graph(video, max_factors) -> FactorGraph(video, None, "cpu", "volume", max_factors);
    slope = distance per frame of index gap (smaller -> more frames within `thresh` -> the max_factors regime)"""
    rng = np.random.default_rng(seed)
    video = FakeVideo(n_frames + 2, seed, stereo, slope)
    graph = make_graph(video, max_factors)
    trace = []
    # ---- DroidFrontend.__initialize
    video.counter.value = warmup
    t1 = warmup
    graph.add_neighborhood_factors(0, t1, r=3)
    trace.append(snapshot(graph, "init.neighborhood"))
    for _ in range(8):
        fake_update(graph, rng)
    graph.add_proximity_factors(0, 0, rad=2, nms=2, thresh=thresh, remove=False)
    trace.append(snapshot(graph, "init.proximity"))
    for _ in range(8):
        fake_update(graph, rng)
    graph.rm_factors(graph.ii < warmup - 4, store=True)
    trace.append(snapshot(graph, "init.rm_old"))
    # ---- DroidFrontend.__update per keyframe
    step = 0
    while video.counter.value < n_frames:
        video.counter.value += 1
        t1 += 1
        step += 1
        if graph.correlation_volumes is not None:
            graph.rm_factors(graph.age > max_age, store=True)
        trace.append(snapshot(graph, f"kf{step}.rm_age"))
        graph.add_proximity_factors(t1 - 5, max(t1 - window, 0), rad=radius, nms=nms, thresh=thresh, beta=beta, remove=True)
        trace.append(snapshot(graph, f"kf{step}.proximity"))
        for _ in range(4):
            fake_update(graph, rng)
        if step % 7 == 3:
            graph.filter_edges()
            trace.append(snapshot(graph, f"kf{step}.filter_edges"))
        if rng.random() < 0.3:
            graph.rm_keyframe(t1 - 2)
            video.counter.value -= 1
            t1 -= 1
            n_frames -= 1                                   # the stream is finite: a dropped frame is gone
            trace.append(snapshot(graph, f"kf{step}.rm_keyframe"))
        else:
            for _ in range(2):
                fake_update(graph, rng)
    return trace


# ---------------------------------------------------------------------------------------------------------------
# update() / update_lowmem(): the operator and the BA are replaced by deterministic stand-ins with the REAL call
# conventions (UpdateModule.forward's, networks/droid_net.py:118-150; video.ba's), so that the trace pins everything
# FactorGraph itself does around them: motion features, state write-back, damping scatter by source frame, inactive
# edges in the BA window, argument order / shapes / contiguity of the BA call, age.
def fake_update_net(net, inp, corr, flow=None, ii=None, jj=None):
    """net [1,E,C,h,w]; flow = motion features [1,E,4,h,w]; -> net', delta [1,E,h,w,2], weight [1,E,h,w,2],
    eta [1,K,h,w], upmask [1,K,576,h,w] (K = number of distinct source frames, in sorted order)"""
    b, E, C, h, w = net.shape
    flow = flow.float()
    net2 = 0.9 * net.float() + 0.1 * torch.tanh(0.01 * flow[:, :, 0:1]) + 0.001 * corr.float().mean(dim=2, keepdim=True)
    delta = (0.5 * torch.tanh(0.05 * flow[:, :, 2:4]) + 0.01).permute(0, 1, 3, 4, 2).contiguous()
    weight = torch.sigmoid(0.02 * flow[:, :, 0:2]).permute(0, 1, 3, 4, 2).contiguous()
    if ii is None:
        return net2, delta, weight
    ii = torch.as_tensor(ii).long().cpu()
    ux, inv = torch.unique(ii, return_inverse=True)
    K = len(ux)
    s = torch.zeros(K, h, w).index_add_(0, inv, net2[0, :, 0])
    cnt = torch.zeros(K).index_add_(0, inv, torch.ones(E))
    eta = (0.01 * (s / cnt.view(-1, 1, 1)).abs() + 0.001 * (ux.float().view(-1, 1, 1) + 1))[None]
    return net2, delta, weight, eta, torch.zeros(1, K, 576, h, w)


def snapshot_update(graph, video, tag):
    d = snapshot(graph, tag)
    d["damping_sum"] = round(float(graph.damping.double().sum()), 6)
    d["damping00"] = [round(float(v), 6) for v in graph.damping[:, 0, 0].tolist()]
    d["flow_sum"] = round(float(graph.gru_estimated_flow.double().sum()), 3)
    d["weight_sum"] = round(float(graph.gru_estimated_flow_weight.double().sum()), 4)
    d["hidden_sum"] = round(float(graph.gru_hidden_states.double().sum()), 4)
    d["ba_calls"] = list(video.ba_log or [])
    d["dirty"] = [int(v) for v in video.dirty.tolist()]
    video.ba_log = []
    return d


def run_update_scenario(make_graph, seed, n_kf=9, stereo=False):
    """make_graph(video, max_factors, update_net) -> FactorGraph"""
    video = FakeVideo(n_kf + 3, seed, stereo)
    video.counter.value = n_kf
    graph = make_graph(video, 48, fake_update_net)
    trace = []
    graph.add_neighborhood_factors(0, n_kf, r=3)
    for k in range(3):
        graph.update(1, use_inactive=True)                       # DroidFrontend.__initialize: t0 = 1
        trace.append(snapshot_update(graph, video, f"init.update{k}"))
    graph.rm_factors(graph.ii < 3, store=True)
    for k in range(2):
        graph.update(None, None, use_inactive=True)              # DroidFrontend.__update: t0 from the edges
        trace.append(snapshot_update(graph, video, f"steady.update{k}"))
    graph.update(None, None, itrs=3, use_inactive=False, EP=1e-3, motion_only=True)
    trace.append(snapshot_update(graph, video, "steady.update_active_only"))
    graph.update_lowmem(steps=2)                                 # global BA path (alt-corr, chunks of 8 source frames)
    trace.append(snapshot_update(graph, video, "lowmem"))
    return trace


# ---------------------------------------------------------------------------------------------------------------
# DroidFrontend (networks/droid_frontend.py:9-121): the keyframe loop on top of FactorGraph, with a stand-in for the
# motion filter (every step delivers one new frame with a fresh identity), the operator and the BA.
def frontend_args(keyframe_thresh=4.0):
    return types.SimpleNamespace(warmup=8, beta=0.3, frontend_nms=1, keyframe_thresh=keyframe_thresh, frontend_window=25,
                                 frontend_thresh=16.0, frontend_radius=2)


def run_frontend_scenario(make_frontend, seed, n_steps=20, slope=1.0, keyframe_thresh=4.0):
    """make_frontend(droid_net, video, args) -> DroidFrontend; droid_net.update_net = fake_update_net"""
    video = FakeVideo(n_steps + 12, seed, False, slope)
    net = types.SimpleNamespace(update_net=fake_update_net)
    front = make_frontend(net, video, frontend_args(keyframe_thresh))
    next_id = 0
    trace = []
    for step in range(n_steps):
        slot = video.counter.value                               # MotionFilter.track -> video.append
        video.intrinsics[slot, 0] = float(next_id)
        video.tstamp[slot] = float(next_id)
        next_id += 1
        video.counter.value += 1
        front()
        d = snapshot_update(front.graph, video, f"step{step}") if front.is_initialized else {"tag": f"step{step}"}
        d.update({"t1": int(front.t1), "counter": int(video.counter.value), "is_initialized": bool(front.is_initialized),
                  "count": int(front.count), "ready": int(video.ready.value),
                  "ids": [int(v) for v in video.intrinsics[:, 0].tolist()],
                  "poses1": [round(float(v), 5) for v in video.poses[:, 1].tolist()],
                  "disps00": [round(float(v), 6) for v in video.disps[:, 0, 0].tolist()]})
        trace.append(d)
    return trace
