"""run the front-end on the procedural stream (or, with DATASET_DIR, on a transforms.json dataset) and print timing /
state statistics (GPU)"""
import os, sys, time, types, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nerf_slam_b200.frontend import RaftVisualFrontend
from nerf_slam_b200.synthetic import SyntheticRoom
W, H, N = int(os.environ.get("W", 640)), int(os.environ.get("H", 480)), int(os.environ.get("N", 120))
wp = os.path.join(ROOT, "oracle", "_ref", "droid.pth")
args = types.SimpleNamespace(buffer=int(os.environ.get("BUF", 100)), stereo=False, multi_gpu=False, weights=wp if os.path.exists(wp) else None)
if os.environ.get("DATASET_DIR"):
    # a transforms.json + PNG directory in the reference's format (nerf_slam_b200.datasets; write one with
    # `python -c "from nerf_slam_b200 import datasets, synthetic; datasets.write_transforms_dataset(synthetic.SyntheticRoom(640, 480, 120), 'out_dir')"`)
    from nerf_slam_b200.datasets import NeRFDataset, dataset_args
    data = NeRFDataset(dataset_args(os.environ["DATASET_DIR"]))
    N = min(N, len(data))
    pk = [data[k] for k in range(N)]
else:
    room = SyntheticRoom(W, H, N, seed=0, step=float(os.environ.get("STEP", 0.012)))
    pk = [room.packet(k) for k in range(N)]
fe = RaftVisualFrontend(np.linalg.inv(pk[0]["poses"][0]), np.eye(4), args, "cuda:0")
ts = []
for k in range(N):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fe.forward(pk[k])
    torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print(f"frame {k:4d} kf_idx {fe.kf_idx:3d} edges {len(getattr(fe,'ii_h',[])):3d} motion {float(getattr(fe,'last_motion',0)):.2f} {ts[-1]*1e3:8.2f} ms", flush=True)
    if fe.stop_condition(): break
ts = np.array(ts)
print(json.dumps(dict(frames=len(ts), kf=fe.kf_idx, total_s=float(ts.sum()), fps=float(len(ts)/ts.sum()),
                      steady_fps=float(len(ts[40:])/ts[40:].sum()) if len(ts) > 50 else None, weights=fe.weights_source, updates=fe.stats["updates"])))
