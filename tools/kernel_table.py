"""per-kernel GPU time of the three steady-state pieces (torch.profiler / CUPTI):
update() graph replay, the per-frame front graph, one NeRF training step."""
import os, sys, json, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity
torch.set_grad_enabled(False)


def table(fn, n, title, top=40):
    fn(); fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            a = agg.setdefault(ev.name, [0, 0.0])
            a[0] += 1; a[1] += ev.device_time
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    tot = sum(v[1] for _, v in rows)
    print(f"== {title}: {tot / n:.1f} us of kernel time per call, {sum(v[0] for _, v in rows) // n} launches per call")
    for name, (c, t) in rows[:top]:
        print(json.dumps(dict(us_per_call=round(t / n, 1), launches_per_call=round(c / n, 2), avg_us=round(t / c, 2), kernel=name[:110])))
    sys.stdout.flush()


job = bench.SlamNerfJob(0, 1, 1)
fe = job.fe
while not (fe.is_initialized and fe.kf_idx >= 14):
    for p in job.make_frames(4, True):
        job.step(p, False)
torch.cuda.synchronize()
print("edges", len(fe.ii_h), "kf", fe.kf_idx)
for _ in range(3):
    fe.update(use_inactive=True)
table(lambda: fe.update(use_inactive=True), 4, "update()")
img = job.make_frames(1, True)[0]
x = img["images"].to(fe.device)[None].permute(0, 1, 4, 2, 3)
for _ in range(3):
    fe._frame_front(x)
table(lambda: fe._frame_front(x), 4, "frame front [graph replay]")
imgs = fe._normalize_imgs(x)
table(lambda: fe._context_encoder(imgs), 4, "context encoder (keyframes only)")
tb = job.nf.ngp
for _ in range(40):
    tb.train_step()
table(tb.train_step, 8, f"nerf train_step ({tb.mlp_backend}) rays={tb.rays_per_batch}")
tb.mlp_backend = "simt"
for _ in range(4):
    tb.train_step()
table(tb.train_step, 8, f"nerf train_step (simt) rays={tb.rays_per_batch}")
