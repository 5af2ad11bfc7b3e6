"""Generate golden vectors by importing the REFERENCE's own Python modules (CPU, fp32) from
/root/reference — run in the build container only (the reference tree does not exist on the GPU box).

  python tests/golden/make_golden_py.py

Missing third-party imports of the reference (lietorch, torch_scatter, droid_backends — empty
submodules / uninstallable) are stubbed ONLY as far as import resolution needs:
`torch_scatter.scatter_mean` gets an index_add implementation (the one call site is
networks/droid_net.py:67); `lietorch`/`droid_backends` are never executed on the paths used here.
Outputs (small, committed): tests/golden/ref_py_*.npz
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    def scatter_mean(src, index, dim=0, dim_size=None):
        n = int(index.max()) + 1 if dim_size is None else dim_size
        shape = list(src.shape); shape[dim] = n
        out = torch.zeros(shape, dtype=src.dtype).index_add_(dim, index, src)
        cnt = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        view = [1] * src.dim(); view[dim] = n
        return out / cnt.view(view)
    ts = types.ModuleType("torch_scatter"); ts.scatter_mean = scatter_mean; ts.scatter_sum = None
    lt = types.ModuleType("lietorch"); lt.SE3 = object; lt.Sim3 = object
    dbk = types.ModuleType("droid_backends")
    ic = types.ModuleType("icecream"); ic.ic = lambda *a, **k: None
    for name, m in (("torch_scatter", ts), ("lietorch", lt), ("droid_backends", dbk), ("icecream", ic)):
        sys.modules.setdefault(name, m)
    # networks/__init__ pulls geometry helpers that import lietorch symbols at module import time only


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(20220922)
    from networks.modules.extractor import BasicEncoder
    from networks.modules.corr import CorrBlock
    from networks.droid_net import UpdateModule
    sd = torch.load(os.path.join(REF, "droid.pth"), map_location="cpu")
    sd = {k.replace("module.", ""): v for k, v in sd.items()}

    def sub(prefix):
        return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    # A1: feature / context encoders on a 64x96 image
    fnet = BasicEncoder(output_dim=128, norm_fn="instance"); fnet.load_state_dict(sub("fnet.")); fnet.eval()
    cnet = BasicEncoder(output_dim=256, norm_fn="none"); cnet.load_state_dict(sub("cnet.")); cnet.eval()
    img = torch.randn(1, 1, 3, 64, 96, generator=g)
    with torch.no_grad():
        f = fnet(img); c = cnet(img)
    np.savez_compressed(os.path.join(OUT, "ref_py_encoders.npz"), img=img.numpy(), fnet=f.numpy(), cnet=c.numpy())
    # A5: update operator, 3 edges on 8x12 maps, with graph aggregation
    um = UpdateModule()
    usd = sub("update.")
    for k in ("weight.2.weight", "weight.2.bias", "delta.2.weight", "delta.2.bias"):
        usd[k] = usd[k][:2]
    um.load_state_dict(usd); um.eval()
    E, H, W = 3, 8, 12
    net = torch.tanh(torch.randn(1, E, 128, H, W, generator=g)); inp = torch.relu(torch.randn(1, E, 128, H, W, generator=g))
    corr = torch.randn(1, E, 196, H, W, generator=g) * 2; flow = torch.randn(1, E, 4, H, W, generator=g) * 3
    ii = torch.tensor([0, 0, 1]); jj = torch.tensor([1, 2, 0])
    with torch.no_grad():
        o = um(net, inp, corr, flow, ii, jj)
    np.savez_compressed(os.path.join(OUT, "ref_py_update.npz"), net=net.numpy(), inp=inp.numpy(), corr=corr.numpy(),
                        flow=flow.numpy(), ii=ii.numpy(), jj=jj.numpy(), out_net=o[0].numpy(), delta=o[1].numpy(),
                        weight=o[2].numpy(), eta=o[3].numpy(), upmask=o[4].numpy())
    # A2: all-pairs correlation + pyramid (fp32 on CPU), 2 edges of 16x16 maps with 128 channels
    f1 = torch.randn(1, 2, 128, 16, 16, generator=g); f2 = torch.randn(1, 2, 128, 16, 16, generator=g)
    blk = CorrBlock(f1, f2)
    np.savez_compressed(os.path.join(OUT, "ref_py_corr.npz"), f1=f1.numpy(), f2=f2.numpy(),
                        **{f"l{i}": p.numpy() for i, p in enumerate(blk.corr_pyramid)})
    # A17: live cvx_upsample (utils/flow_viz.py) incl. its in-place -inf border masking
    try:
        from utils.flow_viz import cvx_upsample
        data = torch.rand(2, 5, 6, 1, generator=g) + 0.1
        mask = torch.randn(2, 576, 5, 6, generator=g)
        up = cvx_upsample(data.clone(), mask.clone())
        up2 = cvx_upsample(data.clone(), mask.clone(), pow=0.5)
        np.savez_compressed(os.path.join(OUT, "ref_py_upsample.npz"), data=data.numpy(), mask=mask.numpy(), up=up.numpy(), up_pow=up2.numpy())
    except Exception as e:      # flow_viz imports plotting libs that may be absent
        print("cvx_upsample golden skipped:", e)
    print("written:", sorted(f for f in os.listdir(OUT) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
