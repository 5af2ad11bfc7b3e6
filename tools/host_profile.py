"""host-side profile of the steady-state frame loop (cProfile): where does the Python thread spend its time?"""
import os, sys, cProfile, pstats, io, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
torch.set_grad_enabled(False)
job = bench.SlamNerfJob(0, 1, int(os.environ.get("NERF_ITERS", 2)))
fe = job.fe
while not (fe.is_initialized and fe.kf_idx >= 12):
    for p in job.make_frames(8, True):
        job.step(p, False)
frames = job.make_frames(72, True)
for p in frames[:8]:
    job.step(p, False)
torch.cuda.synchronize()
pr = cProfile.Profile()
fe.timers.t.clear()
t0 = time.perf_counter()
if os.environ.get("NSLAM_CPROFILE", "1") == "1":
    pr.enable()
for p in frames[8:]:
    job.step(p, False)
torch.cuda.synchronize()
pr.disable()
dt = time.perf_counter() - t0
print("host section timers (ms total, calls):", fe.timers.report())
print(f"64 frames in {dt * 1e3:.1f} ms -> {dt / 64 * 1e3:.2f} ms/frame, kf now {fe.kf_idx}, updates {fe.stats['updates']}")
for key in (("cumulative", "tottime") if os.environ.get("NSLAM_CPROFILE", "1") == "1" else ()):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue()[:9000])
