"""CUDA product path vs the golden vectors produced by the reference's own Python modules."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHTS = os.path.join(ROOT, "oracle", "_ref", "droid.pth")


def test_corr_volume_tc_vs_reference_python():
    from nerf_slam_b200 import droid_backends as db
    d = np.load(os.path.join(G, "ref_py_corr.npz"))
    f1, f2 = torch.from_numpy(d["f1"][0]), torch.from_numpy(d["f2"][0])
    E = f1.shape[0]
    fm = torch.cat([f1, f2], 0).half().permute(0, 2, 3, 1).contiguous().to(DEV)
    ii = torch.arange(E, dtype=torch.int32, device=DEV)
    pyr = db.corr_volume_build(fm, ii, ii + E)
    for l in range(4):
        err = np.abs(pyr[l].float().cpu().numpy() - d[f"l{l}"]).max()
        assert err < 6e-2, (l, err)           # fp16 inputs + fp16 volume vs the fp32 reference run


def test_cvx_upsample_vs_reference_python():
    from nerf_slam_b200 import droid_backends as db
    d = np.load(os.path.join(G, "ref_py_upsample.npz"))
    T = lambda x: torch.from_numpy(x).to(DEV)
    up = db.cvx_upsample(T(d["data"]), T(d["mask"]), 1.0).cpu().numpy()
    assert np.allclose(up, d["up"], atol=1e-5)
    up2 = db.cvx_upsample(T(d["data"]), T(d["mask"]), 0.5).cpu().numpy()
    assert np.allclose(up2, d["up_pow"], atol=1e-4)
    # channels-last mask variant
    up3 = db.cvx_upsample(T(d["data"]), T(d["mask"]).permute(0, 2, 3, 1).contiguous(), 1.0, mask_nhwc=True).cpu().numpy()
    assert np.allclose(up3, d["up"], atol=1e-5)


def test_update_operator_tc_vs_reference_python():
    """fused tensor-core update operator (fp16) vs the reference UpdateModule run in fp32 on CPU"""
    if not os.path.exists(WEIGHTS):
        pytest.skip("droid.pth not shipped")
    from nerf_slam_b200.conv import CORR_PAD, UpdateOperatorTC
    from nerf_slam_b200.networks import UpdateModule, load_droid_weights
    d = np.load(os.path.join(G, "ref_py_update.npz"))
    um = UpdateModule(); um.load_state_dict(load_droid_weights(WEIGHTS), "update_net.")
    um.to(device=DEV, dtype=torch.float16)
    op = UpdateOperatorTC(um, DEV)
    nhwc = lambda x: torch.from_numpy(x[0]).permute(0, 2, 3, 1).contiguous()
    corr = torch.zeros(3, 8, 12, CORR_PAD); corr[..., :196] = nhwc(d["corr"])
    o = op.call_reference_convention(nhwc(d["net"]).half().to(DEV), nhwc(d["inp"]).half().to(DEV), corr.half().to(DEV),
                                     torch.from_numpy(d["flow"][0]).to(DEV), torch.from_numpy(d["ii"]).to(DEV))
    ref_net = torch.from_numpy(d["out_net"][0]).permute(0, 2, 3, 1)
    assert float((o[0].float().cpu() - ref_net).abs().max()) < 3e-2
    assert float((o[1].cpu() - torch.from_numpy(d["delta"][0])).abs().max()) < 8e-2        # px
    assert float((o[2].cpu() - torch.from_numpy(d["weight"][0])).abs().max()) < 2e-2
    assert float((o[3].cpu() - torch.from_numpy(d["eta"][0])).abs().max()) < 2e-3
    assert float((o[4].float().cpu() - torch.from_numpy(d["upmask"][0]).permute(0, 2, 3, 1)).abs().max()) < 0.1


def test_encoders_vs_reference_python():
    if not os.path.exists(WEIGHTS):
        pytest.skip("droid.pth not shipped")
    from nerf_slam_b200.networks import BasicEncoder, load_droid_weights
    sd = load_droid_weights(WEIGHTS)
    d = np.load(os.path.join(G, "ref_py_encoders.npz"))
    f = BasicEncoder(128, "instance"); f.load_state_dict(sd, "feature_net."); f.to(device=DEV, dtype=torch.float16)
    c = BasicEncoder(256, "none"); c.load_state_dict(sd, "context_net."); c.to(device=DEV, dtype=torch.float16)
    x = torch.from_numpy(d["img"]).to(DEV)
    assert float((f(x).float().cpu() - torch.from_numpy(d["fnet"])).abs().max()) < 5e-2
    assert float((c(x).float().cpu() - torch.from_numpy(d["cnet"])).abs().max()) < 5e-2
    # the same networks on the tensor-core convolution kernel (im2col first layer, fused statistics, stride-2 stores)
    from nerf_slam_b200.conv import EncoderTC
    ft, ct = EncoderTC(f, DEV), EncoderTC(c, DEV)
    xb = x.reshape(-1, *x.shape[-3:])
    got_f = ft(xb).float().cpu().reshape(d["fnet"].shape); got_c = ct(xb).float().cpu().reshape(d["cnet"].shape)
    assert float((got_f - torch.from_numpy(d["fnet"])).abs().max()) < 5e-2
    assert float((got_c - torch.from_numpy(d["cnet"])).abs().max()) < 5e-2
