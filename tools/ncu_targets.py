"""workload for `ncu --profile-from-start off`: primes the SLAM+NeRF job, then runs ONE eager
update() (no CUDA graph), one per-frame front, one corr-volume build and one NeRF training step
inside cudaProfilerStart/Stop."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
torch.set_grad_enabled(False)
job = bench.SlamNerfJob(0, 1, 1)
fe = job.fe
fe.use_cuda_graphs = False
while not (fe.is_initialized and fe.kf_idx >= 14):
    for p in job.make_frames(4, True):
        job.step(p, False)
tb = job.nf.ngp
for _ in range(40):
    tb.train_step()
fe.update(use_inactive=True)
img = job.make_frames(1, True)[0]
x = img["images"].to(fe.device)[None].permute(0, 1, 4, 2, 3)
fe._frame_front(x)
torch.cuda.synchronize()
print("edges", len(fe.ii_h), "kf", fe.kf_idx, "rays", tb.rays_per_batch, flush=True)
torch.cuda.profiler.start()
fe.update(use_inactive=True)
fe.use_cuda_graphs = False
fe._frame_front_body()
tb.training_step = 17            # not a density-grid step
tb.train_step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
