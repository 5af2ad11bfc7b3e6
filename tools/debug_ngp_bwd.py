"""debug: per-layer gradient error of the tensor-core backward vs the SIMT backward"""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nerf_slam_b200 import _lib, pyngp
DEV = "cuda"
lib = _lib.load()
for (R, per) in ((1, 8), (8, 16), (24, 11), (300, 13)):
    tb = pyngp.Testbed(seed=5, max_samples=1 << 16, max_rays=1 << 13)
    tb.create_empty_nerf_dataset(8, 1.0, None, 4, None)
    g = torch.Generator().manual_seed(1)
    tb.grid_master.copy_(((torch.rand(tb.grid_master.shape, generator=g) * 2 - 1) * 0.5).to(DEV)); tb.grid_half.copy_(tb.grid_master.half())
    n = R * per
    x = torch.rand(n, 3, generator=g) * 0.6 + 0.2
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).repeat_interleave(per, 0)
    dt = torch.rand(n, generator=g) * 0.05 + 0.01
    tdist = (torch.arange(per).float()[None] * 0.05 + 0.3).repeat(R, 1).reshape(-1)
    coords = torch.cat([x, dt[:, None], d], -1).contiguous()
    rays = torch.zeros(R, 16); rays[:, 3:6] = d[::per]; rays[:, 6] = 0.9
    rays[:, 7] = torch.rand(R, generator=g) * 0.5 + 0.2; rays[:, 8] = torch.rand(R, generator=g) * 0.5 + 0.1; rays[:, 9:12] = torch.rand(R, 3, generator=g)
    ri = rays.view(torch.int32); ri[:, 12] = torch.arange(R, dtype=torch.int32) * per; ri[:, 13] = per
    tb._bufs["rays"][:R].copy_(rays.to(DEV)); tb._bufs["coords"][:n].copy_(coords.to(DEV)); tb._bufs["tdist"][:n].copy_(tdist.to(DEV))
    res = {}
    for be in ("simt", "tc"):
        tb.mlp_grad.zero_(); tb.grid_grad.zero_()
        if be == "simt":
            _lib.check(lib.nslam_ngp_loss_backward(ctypes.byref(tb.model), ctypes.byref(tb.batch), R, n, 1.0, .2, .4, .6, tb.num_sms, _lib.stream_ptr()), "a")
        else:
            tb.pack_weights()
            _lib.check(lib.nslam_ngp_loss_backward_tc(ctypes.byref(tb.model), ctypes.byref(tb.batch), _lib.ptr(tb.packed), R, n, 1.0, .2, .4, .6, 1024.0, tb.num_sms, _lib.stream_ptr()), "b")
        torch.cuda.synchronize()
        res[be] = (tb.mlp_grad.clone(), tb.grid_grad.clone(), tb._bufs["dout"][:n].clone(), float(tb._bufs["loss"].item()))
    print(f"R={R} per={per} n={n} loss simt {res['simt'][3]:.6f} tc {res['tc'][3]:.6f}; max|dout| {float(res['tc'][2].abs().max()):.3e} (w {float(res['tc'][2][:,3].abs().max()):.3e})")
    off = 0
    for name, (i, o) in dict(W1=(32, 64), W2=(64, 16), W3=(32, 64), W4=(64, 64), W5=(64, 16)).items():
        a = res["simt"][0][off:off + i * o].view(i, o); b = res["tc"][0][off:off + i * o].view(i, o); off += i * o
        err = float((a - b).abs().max() / (a.abs().max() + 1e-20))
        rowerr = ((a - b).abs().max(1).values / (a.abs().max() + 1e-20))
        fro = float((a - b).norm() / (a.norm() + 1e-20))
        print(f"  {name}: fro {fro:.3e} rel err {err:.3e}  |ref|max {float(a.abs().max()):.3e}  finite {bool(torch.isfinite(b).all())}  worst rows {rowerr.topk(3).indices.tolist()} cols {((a-b).abs().max(0).values).topk(3).indices.tolist()}")
    a, b = res["simt"][1], res["tc"][1]
    print(f"  grid: fro {float((a - b).norm() / (a.norm() + 1e-20)):.3e} rel err {float((a - b).abs().max() / (a.abs().max() + 1e-20)):.3e}")
