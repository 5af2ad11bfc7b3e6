"""host-side logic that needs no GPU: weight packing for the tensor-core convolution, hash-grid level table,
pose conversion of the NeRF hand-off"""
import math

import numpy as np
import torch

from oracle import ngp as ongp
from oracle import se3 as ose3


def test_pack_weights_layout():
    """packed image = [tap][source][64-channel block] blocks of [N_pad][64] fp16, 16-byte chunk j of row n stored at
    chunk position j ^ (n & 7)  (the 128-byte swizzle the UMMA shared-memory descriptor expects)"""
    from nerf_slam_b200.conv import pack_weights
    g = torch.Generator().manual_seed(0)
    for (N, srcs, k, npad) in ((16, [128], 3, None), (128, [128, 128, 128, 64], 3, None), (32, [152], 1, None),
                               (4, [256], 3, 16), (64, [32], 1, None)):
        cin = sum(srcs)
        w = torch.randn(N, cin, k, k, generator=g)
        img = pack_weights(w, srcs, n_pad=npad).float()
        Np = npad or N
        nblk = k * k * sum((c + 63) // 64 for c in srcs)
        assert img.numel() == nblk * Np * 64
        img = img.view(nblk, Np, 8, 8)
        b = 0
        for ky in range(k):
            for kx in range(k):
                off = 0
                for C in srcs:
                    for cb in range((C + 63) // 64):
                        blk = torch.zeros(Np, 64)
                        cs, ce = cb * 64, min(C, cb * 64 + 64)
                        blk[:N, :ce - cs] = w[:, off + cs:off + ce, ky, kx].half().float()
                        for n in range(0, Np, max(1, Np // 7)):           # sample rows
                            for j in range(8):
                                assert torch.equal(img[b, n, j ^ (n & 7)], blk[n, 8 * j:8 * j + 8]), (N, b, n, j)
                        b += 1
                    off += C


def test_level_table_matches_oracle():
    from nerf_slam_b200 import pyngp
    for aabb in (1.0, 4.0, 16.0):
        rows, total = pyngp.level_table(aabb)
        lv, tot = ongp.level_params(aabb)
        assert total == tot and len(rows) == len(lv) == 16
        for (sc, res, n, off, dense), o in zip(rows, lv):
            assert (sc, res, n, off) == (o[0], o[1], o[2], o[3])
            assert bool(dense) == (res ** 3 <= n)


def test_pose_tq_to_c2w_is_inverse_of_cam_T_world():
    from nerf_slam_b200.nerf_fusion import _pose_tq_to_c2w
    rng = np.random.default_rng(3)
    q = rng.normal(size=(5, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    t = rng.normal(size=(5, 3))
    tq = np.concatenate([t, q], 1)
    c2w = _pose_tq_to_c2w(tq)
    for k in range(5):
        T = ose3.matrix(tq[k].astype(np.float64)) if hasattr(ose3, "matrix") else None
        if T is None:
            x, y, z, w = q[k]
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t[k]
        assert np.allclose(c2w[k] @ T, np.eye(4), atol=1e-9)
