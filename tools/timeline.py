"""Where does a frame's time go at one GPU?  CUPTI timeline (torch.profiler chrome trace) of N steady-state frames of the
bench job: per stream busy time / idle gaps, the kernels' durations inside the co-run next to their durations alone
(tools/kernel_table.py), and what the SLAM stream was waiting for in its largest gaps.

  python tools/timeline.py [frames=48] [nerf_iters=2] > gpurun_out/timeline.log
"""
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from torch.profiler import profile, ProfilerActivity

torch.set_grad_enabled(False)
FRAMES = int(sys.argv[1]) if len(sys.argv) > 1 else 48
NERF_ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 2

job = bench.SlamNerfJob(0, 1, NERF_ITERS)
fe = job.fe
n = 0
while not (fe.is_initialized and fe.kf_idx >= 14) or n < 64:
    for p in job.make_frames(4, True):
        job.step(p, False); n += 1
torch.cuda.synchronize()
frames = job.make_frames(FRAMES, True)
import time
t0 = time.perf_counter()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for p in frames:
        job.step(p, False)
    torch.cuda.synchronize()
wall = time.perf_counter() - t0
path = "/tmp/nslam_timeline.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
print(f"frames {FRAMES} nerf_iters {NERF_ITERS}: wall {wall * 1e3:.1f} ms under the profiler = {wall / FRAMES * 1e3:.2f} ms/frame; "
      f"device span {span / 1e3:.1f} ms; keyframes {fe.kf_idx}, edges {len(fe.ii_h)}")
streams = collections.defaultdict(list)
for e in ev:
    streams[e["args"].get("stream", -1)].append(e)
slam_stream = max(streams, key=lambda s: sum(1 for e in streams[s] if "conv_igemm" in e["name"] or "ba_" in e["name"]))
for s, es in sorted(streams.items(), key=lambda kv: -len(kv[1])):
    busy = sum(e["dur"] for e in es)
    print(f"stream {s}{' (SLAM)' if s == slam_stream else ''}: {len(es)} launches, busy {busy / 1e3:.1f} ms = {busy / span * 100:.1f} % of the span, "
          f"{busy / FRAMES:.0f} us per frame")
es = streams[slam_stream]
gaps = []
for a, b in zip(es[:-1], es[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0:
        gaps.append((g, a["name"][:60], b["name"][:60]))
tot_gap = sum(g for g, _, _ in gaps)
print(f"SLAM stream idle between its kernels: {tot_gap / 1e3:.1f} ms ({tot_gap / span * 100:.1f} % of the span); "
      f"gaps > 20 us: {sum(g for g, _, _ in gaps if g > 20) / 1e3:.1f} ms in {sum(1 for g, _, _ in gaps if g > 20)} gaps; "
      f"> 100 us: {sum(g for g, _, _ in gaps if g > 100) / 1e3:.1f} ms in {sum(1 for g, _, _ in gaps if g > 100)}")
by_next = collections.defaultdict(lambda: [0, 0.0])
for g, a, b in gaps:
    if g > 20:
        k = f"{a.split('(')[0][-44:]} -> {b.split('(')[0][-44:]}"
        by_next[k][0] += 1; by_next[k][1] += g
print("largest idle classes on the SLAM stream (previous kernel -> next kernel):")
for k, (c, t) in sorted(by_next.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t / 1e3:7.2f} ms in {c:4d} gaps (avg {t / c:6.0f} us)  {k}")
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[(e["args"].get("stream", -1) == slam_stream, e["name"][:90])]
    a[0] += 1; a[1] += e["dur"]
print("kernels by total time (co-run durations):")
for (is_slam, name), (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    print(f"  {'SLAM' if is_slam else 'other'} {t / FRAMES:7.1f} us/frame  {c / FRAMES:5.2f} launches/frame  avg {t / c:7.1f} us  {name}")
