"""fused glue kernels around the update operator and the encoder's instance norm vs plain torch"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_inorm_stats_apply_vs_torch():
    from nerf_slam_b200 import networks as nw
    g = torch.Generator().manual_seed(0)
    for (B, C, H, W) in ((2, 32, 60, 81), (1, 64, 33, 40), (1, 128, 17, 23)):
        x = (torch.randn(B, C, H, W, generator=g) * 2 + 0.3).half().to(DEV).contiguous(memory_format=torch.channels_last)
        r = torch.randn(B, C, H, W, generator=g).half().to(DEV).contiguous(memory_format=torch.channels_last)
        ref1 = torch.relu(torch.nn.functional.instance_norm(x))
        ref2 = torch.relu(r + ref1)
        ref3 = torch.relu(torch.nn.functional.instance_norm(r) + ref1)
        y = x.clone(memory_format=torch.preserve_format); nw._inorm_apply(y, nw._inorm_stats(y))
        assert float((y.float() - ref1.float()).abs().max()) < 4e-3
        y = x.clone(memory_format=torch.preserve_format); nw._inorm_apply(y, nw._inorm_stats(y), res=r)
        assert float((y.float() - ref2.float()).abs().max()) < 8e-3
        y = x.clone(memory_format=torch.preserve_format); nw._inorm_apply(y, nw._inorm_stats(y), res=r, res_st=nw._inorm_stats(r))
        assert float((y.float() - ref3.float()).abs().max()) < 8e-3


def test_motion_im2col_matches_unfold():
    from nerf_slam_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    E, H, W = 3, 11, 14
    yy, xx = torch.meshgrid(torch.arange(H).float(), torch.arange(W).float(), indexing="ij")
    c0 = torch.stack([xx, yy], -1).contiguous().to(DEV)
    c1 = (c0[None] + torch.randn(E, H, W, 2, generator=g).to(DEV) * 40).contiguous()      # some beyond the +-64 clamp
    tg = (c1 + torch.randn(E, H, W, 2, generator=g).to(DEV) * 40).contiguous()
    out = torch.full((E, H, W, 200), 7.0, dtype=torch.float16, device=DEV)
    _lib.check(lib.nslam_motion_im2col(_lib.ptr(c1), _lib.ptr(c0), _lib.ptr(tg), _lib.ptr(out), E, H, W, _lib.stream_ptr()), "im2col")
    motion = torch.cat([c1 - c0, tg - c1], -1).clamp(-64, 64).half().permute(0, 3, 1, 2).float()   # [E,4,H,W]
    cols = torch.nn.functional.unfold(motion, 7, padding=3).view(E, 4, 49, H, W).permute(0, 3, 4, 2, 1).reshape(E, H, W, 196)
    assert torch.equal(out[..., :196].float(), cols)
    assert float(out[..., 196:].abs().max()) == 0.0
    # target = NULL -> zero residual channels
    _lib.check(lib.nslam_motion_im2col(_lib.ptr(c1), _lib.ptr(c0), None, _lib.ptr(out), E, H, W, _lib.stream_ptr()), "im2col")
    assert float(out.view(E, H, W, 50, 4)[..., :49, 2:].abs().max()) == 0.0


def test_flow_heads_post_and_eta_damping():
    from nerf_slam_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(2)
    E, H, W, K = 4, 9, 13, 3
    h2 = torch.randn(E, H, W, 16, generator=g).half().to(DEV)
    c1 = torch.randn(E, H, W, 2, generator=g).to(DEV) * 10
    flow = torch.empty(E, H, W, 2, device=DEV); conf = torch.empty_like(flow)
    bt = torch.zeros(E + 2, 2, H, W, device=DEV); bw = torch.zeros_like(bt)
    _lib.check(lib.nslam_flow_heads_post(_lib.ptr(h2), _lib.ptr(c1), _lib.ptr(flow), _lib.ptr(conf), _lib.ptr(bt[2:]), _lib.ptr(bw[2:]),
                                         E, H * W, _lib.stream_ptr()), "post")
    rf = c1 + h2[..., 0:2].float(); rc = torch.sigmoid(h2[..., 2:4].float())
    assert torch.allclose(flow, rf, atol=1e-6) and torch.allclose(conf, rc, atol=1e-6)
    assert torch.allclose(bt[2:], rf.permute(0, 3, 1, 2), atol=1e-6) and torch.allclose(bw[2:], rc.permute(0, 3, 1, 2), atol=1e-6)
    assert float(bt[:2].abs().max()) == 0.0
    e16 = (torch.randn(K, H, W, 16, generator=g) * 8).half().to(DEV)
    ux = torch.tensor([1, 4, 5], device=DEV); kx = torch.tensor([0, 1, 4, 5], device=DEV)
    damping = torch.full((8, H, W), 0.5, device=DEV); ba = torch.empty(4, H, W, device=DEV)
    _lib.check(lib.nslam_eta_damping(_lib.ptr(e16), _lib.ptr(ux), _lib.ptr(damping), K, _lib.ptr(kx), _lib.ptr(ba), 4, H * W, 1e-7,
                                     _lib.stream_ptr()), "eta")
    ref = torch.full((8, H, W), 0.5, device=DEV); ref[ux] = 0.01 * torch.nn.functional.softplus(e16[..., 0].float())
    assert torch.allclose(damping, ref, rtol=1e-5, atol=1e-7)
    assert torch.allclose(ba, 0.2 * ref[kx] + 1e-7, rtol=1e-5, atol=1e-8)


def test_segment_mean():
    from nerf_slam_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    E, HW = 7, 50
    a = torch.randn(E, HW, 128, generator=g).half().to(DEV)
    ii = np.array([3, 3, 5, 9, 5, 3, 9])
    ux, inv = np.unique(ii, return_inverse=True)
    order = np.argsort(inv, kind="stable").astype(np.int32)
    ptr = np.zeros(len(ux) + 1, np.int32); np.cumsum(np.bincount(inv), out=ptr[1:])
    out = torch.empty(len(ux), HW, 128, dtype=torch.float16, device=DEV)
    ptr_d, order_d = torch.as_tensor(ptr, device=DEV), torch.as_tensor(order, device=DEV)     # keep alive across the launch
    _lib.check(lib.nslam_segment_mean(_lib.ptr(a), _lib.ptr(ptr_d), _lib.ptr(order_d), _lib.ptr(out), len(ux), HW,
                                      _lib.stream_ptr()), "segmean")
    for k, u in enumerate(ux):
        ref = a[torch.as_tensor(ii == u, device=DEV)].float().mean(0)
        assert float((out[k].float() - ref).abs().max()) < 2e-3


def test_cvx_upsample2_indexed_matches_single_plane():
    from nerf_slam_b200 import droid_backends as db
    g = torch.Generator().manual_seed(4)
    K, ht, wd, B = 3, 6, 9, 7
    mask = torch.randn(K, ht, wd, 576, generator=g).half().to(DEV)
    d1 = torch.rand(B, ht, wd, generator=g).to(DEV); d2 = torch.rand(B, ht, wd, generator=g).to(DEV)
    idx = torch.tensor([5, 0, 2], device=DEV)
    o1 = torch.full((B, 8 * ht, 8 * wd), -1.0, device=DEV); o2 = torch.full_like(o1, -2.0)
    db.cvx_upsample2(d1, d2, mask, o1, o2, index=idx)
    r1 = db.cvx_upsample(d1[idx].unsqueeze(-1), mask, mask_nhwc=True).squeeze(-1)
    r2 = db.cvx_upsample(d2[idx].unsqueeze(-1), mask, mask_nhwc=True).squeeze(-1)
    assert torch.equal(o1[idx], r1) and torch.equal(o2[idx], r2)
    keep = torch.tensor([1, 3, 4, 6], device=DEV)
    assert float((o1[keep] + 1).abs().max()) == 0.0 and float((o2[keep] + 2).abs().max()) == 0.0
