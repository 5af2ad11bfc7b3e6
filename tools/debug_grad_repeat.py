"""is nslam_ngp_train_step_tc's gradient a pure function of (model, images, seed)?  (diagnostic for tests/test_gpu_dist.py)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.test_gpu_dist import _testbed_with_images, _gradient
torch.set_grad_enabled(False)
tb = _testbed_with_images(0)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
g = [_gradient(tb, 1000) for _ in range(3)] + [_gradient(tb, 1017) for _ in range(2)] + [_gradient(tb, 1000)]
print("seed 1000 call 0 vs 1:", rel(g[0][0], g[1][0]), rel(g[0][1], g[1][1]))
print("seed 1000 call 1 vs 2:", rel(g[1][0], g[2][0]), rel(g[1][1], g[2][1]))
print("seed 1017 call 0 vs 1:", rel(g[3][0], g[4][0]), rel(g[3][1], g[4][1]))
print("seed 1000 call 2 vs after 1017:", rel(g[2][0], g[5][0]), rel(g[2][1], g[5][1]))
print("seed 1000 vs 1017 (different batches):", rel(g[0][0], g[3][0]), rel(g[0][1], g[3][1]))
tb2 = _testbed_with_images(0)
h = _gradient(tb2, 1000)
print("second testbed, seed 1000 vs first testbed:", rel(h[0], g[0][0]), rel(h[1], g[0][1]))
print("counters", tb._bufs["counters"][:4].tolist() if "counters" in tb._bufs else None)
