"""component timing of the steady-state hot loop (CUDA events + host wall clock)"""
import os, sys, time, types, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from nerf_slam_b200 import droid_backends as db
torch.set_grad_enabled(False)
job = bench.SlamNerfJob(0, 1, 0)
fe = job.fe
while not (fe.is_initialized and fe.kf_idx >= 14):
    for p in job.make_frames(4, True):
        job.step(p, False)
torch.cuda.synchronize()
E = len(fe.ii_h)
print("edges", E, "kf", fe.kf_idx)

def T(fn, n=10, name=""):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    print(json.dumps(dict(op=name, gpu_ms=round(e0.elapsed_time(e1) / n, 3), host_issue_ms=round(t_host * 1e3, 3))), flush=True)

coords1, _ = fe.reproject(fe.ii, fe.jj)
T(lambda: fe.reproject(fe.ii, fe.jj), name="reproject")
T(lambda: fe.corr_pool.lookup(fe.slots_d, coords1, nhwc=True), name="corr_lookup(nhwc)")
corr = fe.corr_pool.lookup(fe.slots_d, coords1, nhwc=True)
inp = fe.cst_contexts_imgs[fe.ii, 0]
T(lambda: fe._run_update_net(fe.gru_hidden_states, inp, corr, coords1, fe.gru_estimated_flow, fe.ii_h), name=f"update operator ({fe.conv_backend}) E={E}")
net, delta, weight, damping, upmask = fe._run_update_net(fe.gru_hidden_states, inp, corr, coords1, fe.gru_estimated_flow, fe.ii_h)
ii, jj = fe.ii_h, fe.jj_h
target = fe.gru_estimated_flow.permute(0, 3, 1, 2).contiguous(); wgt = fe.gru_estimated_flow_weight.permute(0, 3, 1, 2).contiguous()
dmp = .2 * fe.damping[torch.as_tensor(np.unique(ii), device=fe.device)].contiguous() + 1e-7
kf0 = max(0, int(ii.min()))
T(lambda: fe.ba(target, wgt, dmp, ii, jj, kf0, None, itrs=2), name="ba(2 iters + cov)")
T(lambda: fe.ba(target, wgt, dmp, ii, jj, kf0, None, itrs=2, compute_covariances=False), name="ba(2 iters, no cov)")
kx = torch.as_tensor(np.unique(ii), device=fe.device)
T(lambda: db.cvx_upsample(fe.cam0_idepths[kx].unsqueeze(-1), upmask, mask_nhwc=True), name="cvx_upsample")
T(lambda: fe.update(use_inactive=True), n=6, name="update() total")
img = job.make_frames(1, True)[0]
imgs = fe._normalize_imgs(img["images"].to(fe.device)[None].permute(0, 1, 4, 2, 3))
T(lambda: fe._feature_encoder(imgs), name="feature_encoder")
T(lambda: fe._context_encoder(imgs), name="context_encoder")
T(lambda: (fe._frame_front(img["images"].to(fe.device)[None].permute(0, 1, 4, 2, 3)), fe.last_motion.item()), name="frame front: fnet + motion filter (graph) + .item sync")
# NeRF
job2 = bench.SlamNerfJob(0, 1, 1)
while not (job2.fe.is_initialized and job2.fe.kf_idx >= 12):
    for p in job2.make_frames(4, True):
        job2.step(p, False)
torch.cuda.synchronize()
tb = job2.nf.ngp
for _ in range(40): tb.train_step()
T(tb.train_step, n=32, name=f"nerf train_step rays={tb.rays_per_batch}")
tb.sync_stats()
print("nerf samples/rays last:", tb._measured, "loss", tb.loss)
