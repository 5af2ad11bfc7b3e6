"""Kernel micro-benchmarks (CUDA events, L2 flushed between iterations) at the SURVEY.md §8(d)
sizes; prints one JSON line per kernel: time, algorithmic bytes/flops, achieved vs measured peak.
Also times the reference's own CUDA kernels (oracle/_ref) on the same tensors when present."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nerf_slam_b200 import droid_backends as db  # noqa: E402
from oracle import build_ref  # noqa: E402
from tests.util import make_targets, make_window  # noqa: E402

DEV = "cuda:0"
PEAKS = {"hbm_gbs": 6486.1, "bf16_tflops": 1710.9}
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    PEAKS.update(json.load(open(p)))


def timeit(fn, iters=10, warmup=3, flush=True):
    buf = torch.empty(256 * 1024 * 1024 // 4, device=DEV) if flush else None
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            buf.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def report(name, ms, nbytes=None, flops=None, **kw):
    r = dict(kernel=name, ms=round(ms, 4), **kw)
    if nbytes:
        r["GBps"] = round(nbytes / ms / 1e6, 1); r["hbm_frac"] = round(r["GBps"] / PEAKS["hbm_gbs"], 3)
    if flops:
        r["TFLOPs"] = round(flops / ms / 1e9, 2); r["tensor_frac"] = round(r["TFLOPs"] / PEAKS["bf16_tflops"], 3)
    print(json.dumps(r), flush=True)


def main():
    E = int(os.environ.get("NSLAM_E", 48)); H, W, C = 60, 80, 128
    HW = H * W
    g = torch.Generator().manual_seed(1236)
    NF = 12
    fm = torch.randn(NF, H, W, C, generator=g).half().to(DEV)
    ii = torch.randint(0, NF, (E,), generator=g).int().to(DEV); jj = torch.randint(0, NF, (E,), generator=g).int().to(DEV)
    lv = sum((H >> l) * (W >> l) for l in range(4))
    vol_bytes = E * (2 * C * HW * 2 + HW * lv * 2)
    Eb = min(E, 16)
    ms = timeit(lambda: db.corr_volume_build(fm, ii[:Eb], jj[:Eb]))
    report("corr_volume_build_tc", ms, nbytes=Eb * (2 * C * HW * 2 + HW * lv * 2), flops=2.0 * Eb * HW * HW * C, E=Eb)
    ms = timeit(lambda: torch.matmul(fm.view(NF, HW, C)[ii[:Eb].long()], fm.view(NF, HW, C)[jj[:Eb].long()].transpose(1, 2)))
    report("torch_matmul_level0_only(cuBLAS)", ms, nbytes=Eb * (2 * C * HW * 2 + HW * HW * 2), E=Eb)
    pyr = db.corr_volume_build(fm, ii, jj)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    coords = (torch.stack([xx, yy], 0)[None].float() + torch.rand(E, 2, H, W, generator=g) * 16 - 8).to(DEV).contiguous()
    ms = timeit(lambda: db.corr_lookup_pyramid(pyr, coords, 3))
    report("corr_lookup_pyramid", ms, nbytes=E * HW * (4 * 64 * 2 + 8 + 196 * 2), E=E)
    refc = build_ref.load("nslam_ref_corr")
    if refc is not None:
        def ref_lookup():
            return [refc.corr_index_forward(pyr[l], coords / 2 ** l, 3)[0] for l in range(4)]
        ms = timeit(ref_lookup)
        report("REFERENCE corr_index_forward x4", ms, nbytes=E * HW * (4 * 64 * 2 + 8 + 196 * 2), E=E)
    # altcorr
    f32 = (fm.float() / 4)
    Ea = 8
    co = coords[:Ea].permute(0, 2, 3, 1)[:, None].contiguous()
    ms = timeit(lambda: db.altcorr_forward(f32[:Ea], f32[:Ea], co, 3))
    report("altcorr_forward(level0)", ms, nbytes=Ea * (2 * HW * C * 4 + HW * 8 + HW * 49 * 4), flops=2.0 * Ea * HW * 64 * C, E=Ea)
    if refc is not None:
        ms = timeit(lambda: refc.altcorr_forward(f32[:Ea], f32[:Ea], co, 3))
        report("REFERENCE altcorr_forward(level0)", ms, nbytes=Ea * (2 * HW * C * 4 + HW * 8 + HW * 49 * 4), E=Ea)
    # BA
    rng = np.random.default_rng(1236)
    poses, disps, intr, ei, ej = make_window(rng, 14, H, W, extra_edges=8)
    target, weight = make_targets(rng, poses, disps, intr, ei, ej)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(DEV)
    ext = T(np.array([0, 0, 0, 0, 0, 0, 1], np.float32))
    eta = T(np.full((14, H, W), 1e-2, np.float32))
    a = (T(poses), T(poses), T(disps), T(intr), ext, T(np.zeros_like(disps)), T(target), T(weight), eta, T(ei), T(ej), 0, 14)
    nE = len(ei)
    prob = db.BAProblem(a[0], a[2], a[3], a[4], a[5], a[6], a[7], a[8], ei, ej, 0, 14)
    ms = timeit(prob.linearize, flush=False)
    report("ba_reduced_camera_matrix(linearise+schur+assemble)", ms, nbytes=nE * HW * (5 + 14) * 4, E=nE, P=14)
    err6 = torch.zeros(6, device=DEV)
    ms = timeit(lambda: prob.solve(0, err6, 1e8), flush=False)
    report("ba_solve(n=%d)" % (6 * 14), ms)
    dx, _, _ = prob.solve(0, err6, 1e8)
    ms = timeit(lambda: prob.depth_update(dx * 0), flush=False)
    report("ba_depth", ms, nbytes=(14 + nE) * 6 * HW * 4 + 3 * 14 * HW * 4)
    refd = build_ref.load("nslam_ref_droid")
    if refd is not None:
        ms = timeit(lambda: refd.reduced_camera_matrix(*a), flush=False)
        report("REFERENCE reduced_camera_matrix(kernels + dense fp64 host glue)", ms, E=nE, P=14)
    # frame distance
    fi, fj = np.meshgrid(np.arange(14), np.arange(14), indexing="ij")
    fi, fj = T(fi.reshape(-1)), T(fj.reshape(-1))
    ms = timeit(lambda: db.frame_distance(a[0], a[2], a[3], fi, fj, 0.3), flush=False)
    report("frame_distance(196 pairs)", ms)
    if refd is not None:
        ms = timeit(lambda: refd.frame_distance(a[0], a[2], a[3], fi, fj, 0.3), flush=False)
        report("REFERENCE frame_distance(196 pairs)", ms)
    # upsample
    K = 14
    mask = torch.randn(K, 576, H, W, generator=g).half().to(DEV)
    d = T(disps).unsqueeze(-1)
    ms = timeit(lambda: db.cvx_upsample(d, mask))
    report("cvx_upsample", ms, nbytes=K * (576 * HW * 2 + HW * 4 + 64 * HW * 4), K=K)
    ms = timeit(lambda: db.reproject(a[0], a[2], a[3], T(ei), T(ej)), flush=False)
    report("reproject", ms, nbytes=nE * HW * (4 + 8 + 4), E=nE)


if __name__ == "__main__":
    main()
