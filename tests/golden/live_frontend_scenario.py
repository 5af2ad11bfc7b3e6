"""Keyframe loop of the LIVE front end (RaftVisualFrontend.forward's keyframe branch, visual_frontend.py:301-365, with
__initialize / __update / rm_keyframe and the graph-management methods underneath), driven identically for
  * the reference's own methods, executed verbatim from its source file (make_golden_live_frontend.py), and
  * this repo's RaftVisualFrontend on a CPU shim (tests/test_cpu_droid.py).
Stand-ins on both sides: `update()` = its bookkeeping (age += 1, a deterministic drift of the inverse depths),
`distance()` = a seeded matrix over frame identities, `reproject()` = coordinates that encode the edge's frames.
The frame identity of a keyframe slot lives in cam0_intrinsics[:, 0] (moved by rm_keyframe like every other buffer)."""
import numpy as np
import torch

HT8, WD8, CH = 16, 16, 4
PARAMS = dict(max_factors=48, max_age=25, frontend_window=25, frontend_radius=2, frontend_nms=1, frontend_thresh=16.0,
              beta=0.3, iters1=4, iters2=2, keyframe_warmup=8, backend_thresh=22.0, backend_radius=2, backend_nms=3)


def bank(seed, n_ids, slope):
    g = torch.Generator().manual_seed(seed)
    a = torch.arange(n_ids).float()
    noise = torch.rand(n_ids, n_ids, generator=g) * 14.0
    d = slope * (a[:, None] - a[None, :]).abs() + 0.5 * (noise + noise.t())
    d[torch.rand(n_ids, n_ids, generator=g) < 0.08] = 7.5
    d[torch.rand(n_ids, n_ids, generator=g) < 0.03] = 150.0
    feats = torch.randn(n_ids, CH, HT8, WD8, generator=g)
    ctx = torch.randn(n_ids, CH, HT8, WD8, generator=g)
    return d.float(), feats, ctx


def coords0():
    y, x = torch.meshgrid(torch.arange(HT8).float(), torch.arange(WD8).float(), indexing="ij")
    return torch.stack([x, y], dim=-1)


def run(acc, seed, n_steps, slope=1.0, keyframe_thresh=4.0):
    """acc: accessor object with
         .put_frame(slot, frame_id)   store features / contexts / identity of the arriving frame in keyframe slot `slot`
         .kf_idx (get/set), .is_initialized, .initialize(), .update() -> bool, .rm_keyframe(k), .snapshot() -> dict,
         .backend(steps) -> [(ii, jj, steps) seen by update_lowmem]"""
    trace = []
    next_id = 0
    acc.put_frame(0, next_id); next_id += 1          # first frame: always a keyframe (forward(), :262-289)
    acc.kf_idx = 1
    for step in range(n_steps):
        k = acc.kf_idx
        acc.put_frame(k, next_id); next_id += 1
        accepted = True
        if not acc.is_initialized:
            if k >= PARAMS["keyframe_warmup"]:
                acc.initialize()
        else:
            if not acc.update():
                acc.rm_keyframe(k - 1)
                accepted = False
        if accepted:
            acc.kf_idx = k + 1
        d = acc.snapshot()
        d.update({"step": step, "accepted": accepted})
        trace.append(d)
    # global BA at the end of the stream (terminate(), visual_frontend.py:1308-1335): backend(7), backend(12);
    # update_lowmem is a stand-in that records the edge set it is given
    for steps in (7, 12):
        log = acc.backend(steps)
        d = acc.snapshot()
        d.update({"step": f"backend{steps}", "accepted": True, "lowmem": log})
        trace.append(d)
    return trace


# ---------------------------------------------------------------------------------------------------------------
# update() of the live class (visual_frontend.py:371-470) around stand-ins for the operator and the BA: what it writes
# back (flow, confidence, hidden state, damping by source frame, age, viz flags) and what it hands to the dense BA
# (active + inactive edges inside the window, kf0, planar targets / weights, 0.2 * damping + EP).
def run_update(acc, seed, n_kf=10, n_updates=3):
    """acc: .put_frame(slot, id), .kf_idx, .add_neighborhood(kf0, kf1, radius), .retire(first_frames),
            .live_update(use_inactive) -> None, .snapshot_update() -> dict (incl. the BA calls since the last snapshot)"""
    for k in range(n_kf):
        acc.put_frame(k, k)
    acc.kf_idx = n_kf - 1
    acc.add_neighborhood(0, n_kf - 1, 3)
    trace = []
    for k in range(2):
        acc.live_update(True)
        trace.append(acc.snapshot_update())
    acc.retire(4)                                     # edges leaving frames 0..3 become inactive (stored targets)
    for k in range(n_updates):
        acc.live_update(True)
        trace.append(acc.snapshot_update())
    acc.live_update(False)
    trace.append(acc.snapshot_update())
    return trace


def digest(t):
    return [list(t.shape), round(float(t.double().sum()), 3), round(float(t.reshape(-1)[0]), 5)]
