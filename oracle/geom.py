"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatements of the projective-geometry operators of the reference:
  A6   pops.projective_transform (jacobian=False)   networks/geom/projective_ops.py:20-145
  A16  frame_distance_kernel                        src/droid_kernels.cu:630-769
       projmap_kernel / iproj_kernel / depth_filter_kernel   src/droid_kernels.cu:539-628,896-967,773-892
  A17  cvx_upsample                                 utils/flow_viz.py:166-183
"""
import numpy as np

from . import se3

F32 = np.float32
MIN_DEPTH_CUDA = 0.25   # src/droid_kernels.cu:26
MIN_DEPTH_PY = 0.2      # networks/geom/projective_ops.py:8


def _grid(ht, wd, dtype):
    v, u = np.meshgrid(np.arange(ht, dtype=dtype), np.arange(wd, dtype=dtype), indexing="ij")
    return u, v


def _rel(poses, ii, jj, stereo_fix=True):
    t, q = se3.rel_se3(poses[ii, :3], poses[ii, 3:], poses[jj, :3], poses[jj, 3:])
    if stereo_fix:
        s = ii == jj
        t[s] = np.array([-0.1, 0, 0], dtype=t.dtype)
        q[s] = np.array([0, 0, 0, 1], dtype=q.dtype)
    return t, q


def reproject(poses, disps, intrinsics, ii, jj, dtype=np.float64):
    """-> coords [E,ht,wd,2], valid [E,ht,wd,1]; intrinsics [N,4] (per frame) or [4]"""
    poses = poses.astype(dtype); disps = disps.astype(dtype)
    K = intrinsics.astype(dtype)
    if K.ndim == 1:
        K = np.broadcast_to(K, (poses.shape[0], 4))
    ht, wd = disps.shape[1:]
    u, v = _grid(ht, wd, dtype)
    Ki, Kj = K[ii][:, None, None, :], K[jj][:, None, None, :]
    X0 = np.stack([(u - Ki[..., 2]) / Ki[..., 0], (v - Ki[..., 3]) / Ki[..., 1],
                   np.ones((len(ii), ht, wd), dtype), disps[ii]], -1)
    t, q = _rel(poses, ii, jj)
    X1 = se3.act_se3(t[:, None, None, :], q[:, None, None, :], X0)
    Z = X1[..., 2]
    Zs = np.where(Z < 0.5 * MIN_DEPTH_PY, 1.0, Z)
    d = 1.0 / Zs
    x = Kj[..., 0] * (X1[..., 0] * d) + Kj[..., 2]
    y = Kj[..., 1] * (X1[..., 1] * d) + Kj[..., 3]
    valid = ((Z > MIN_DEPTH_PY) & (X0[..., 2] > MIN_DEPTH_PY)).astype(dtype)
    return np.stack([x, y], -1), valid[..., None]


def frame_distance(poses, disps, intr, ii, jj, beta, dtype=np.float32):
    """single direction, like the kernel (`for n<1`, src/droid_kernels.cu:682).
    dtype=float32 reproduces the kernel's thread-strided partial sums and reduction tree
    (without FMA contraction, so agreement with the GPU is ~1 ulp, not bitwise)."""
    poses = poses.astype(dtype); disps = disps.astype(dtype)
    fx, fy, cx, cy = [dtype(x) for x in intr]
    beta = dtype(beta)
    ht, wd = disps.shape[1:]
    hw = ht * wd
    u, v = _grid(ht, wd, dtype)
    u, v = u.reshape(-1), v.reshape(-1)
    out = np.zeros(len(ii), dtype=dtype)
    for b, (i, j) in enumerate(zip(ii, jj)):
        t, q = se3.rel_se3(poses[i, :3], poses[i, 3:], poses[j, :3], poses[j, 3:])
        Xi = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones(hw, dtype), disps[i].reshape(-1)], -1)
        Xj = se3.act_se3(t, q, Xi)
        with np.errstate(divide="ignore", invalid="ignore"):
            du = fx * (Xj[:, 0] / Xj[:, 2]) + cx - u
            dv = fy * (Xj[:, 1] / Xj[:, 2]) + cy - v
            d1 = np.sqrt(du * du + dv * dv)
            ok1 = Xj[:, 2] > MIN_DEPTH_CUDA
            X2 = Xi[:, :3] + Xi[:, 3:4] * t
            du = fx * (X2[:, 0] / X2[:, 2]) + cx - u
            dv = fy * (X2[:, 1] / X2[:, 2]) + cy - v
            d2 = np.sqrt(du * du + dv * dv)
            ok2 = X2[:, 2] > MIN_DEPTH_CUDA
        if dtype == np.float32:
            acc = np.zeros(256, dtype); val = np.zeros(256, dtype); tot = np.zeros(256, dtype)
            nrounds = (hw + 255) // 256
            one_m = dtype(1) - beta
            for r in range(nrounds):
                k = np.arange(r * 256, min(hw, (r + 1) * 256))
                th = k - r * 256
                tot[th] += beta
                acc[th] = np.where(ok1[k], acc[th] + beta * d1[k], acc[th])
                val[th] = np.where(ok1[k], val[th] + beta, val[th])
                tot[th] += one_m
                acc[th] = np.where(ok2[k], acc[th] + one_m * d2[k], acc[th])
                val[th] = np.where(ok2[k], val[th] + one_m, val[th])

            def tree(s):
                s = s.copy()
                for st in (128, 64, 32, 16, 8, 4, 2, 1):
                    s[:st] = s[:st] + s[st:2 * st]
                return s[0]
            A, V, T = tree(acc), tree(val), tree(tot)
        else:
            A = (beta * d1[ok1]).sum() + ((1 - beta) * d2[ok2]).sum()
            V = beta * ok1.sum() + (1 - beta) * ok2.sum()
            T = dtype(hw)
        out[b] = 1000.0 if (float(V) / (float(T) + 1e-8) < 0.75) else A / V
    return out


def projmap(poses, disps, intr, ii, jj, dtype=np.float64):
    poses = poses.astype(dtype); disps = disps.astype(dtype)
    fx, fy, cx, cy = [dtype(x) for x in intr]
    ht, wd = disps.shape[1:]
    u, v = _grid(ht, wd, dtype)
    Xi = np.stack([np.broadcast_to((u - cx) / fx, (len(ii), ht, wd)),
                   np.broadcast_to((v - cy) / fy, (len(ii), ht, wd)),
                   np.ones((len(ii), ht, wd), dtype), disps[ii]], -1)
    t, q = _rel(poses, np.asarray(ii), np.asarray(jj), stereo_fix=False)
    Xj = se3.act_se3(t[:, None, None, :], q[:, None, None, :], Xi)
    ok = Xj[..., 2] > 0.01
    Zs = np.where(ok, Xj[..., 2], 1.0)
    c0 = np.where(ok, fx * (Xj[..., 0] / Zs) + cx, u)
    c1 = np.where(ok, fy * (Xj[..., 1] / Zs) + cy, v)
    coords = np.stack([c0, c1, np.zeros_like(c0)], -1)
    valid = (Xj[..., 2] > MIN_DEPTH_CUDA).astype(dtype)[..., None]
    return coords, valid


def iproj(poses, disps, intr, dtype=np.float64):
    poses = poses.astype(dtype); disps = disps.astype(dtype)
    fx, fy, cx, cy = [dtype(x) for x in intr]
    n, ht, wd = disps.shape
    u, v = _grid(ht, wd, dtype)
    Xi = np.stack([np.broadcast_to((u - cx) / fx, (n, ht, wd)), np.broadcast_to((v - cy) / fy, (n, ht, wd)),
                   np.ones((n, ht, wd), dtype), disps], -1)
    Xj = se3.act_se3(poses[:, None, None, :3], poses[:, None, None, 3:], Xi)
    return Xj[..., :3] / Xj[..., 3:4]


def depth_filter(poses, disps, intr, inds, thresh, dtype=np.float64):
    poses = poses.astype(dtype); disps = disps.astype(dtype)
    fx, fy, cx, cy = [dtype(x) for x in intr]
    num, ht, wd = disps.shape
    u, v = _grid(ht, wd, dtype)
    out = np.zeros((len(inds), ht, wd), dtype)
    for b, ix in enumerate(inds):
        for nb in range(6):
            jx = ix - nb - 1 if nb < 3 else ix + nb
            if jx < 0 or jx >= num:
                continue
            t, q = se3.rel_se3(poses[ix, :3], poses[ix, 3:], poses[jx, :3], poses[jx, 3:])
            Xi = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones((ht, wd), dtype), disps[ix]], -1)
            Xj = se3.act_se3(t, q, Xi)
            with np.errstate(divide="ignore", invalid="ignore"):
                uj = fx * (Xj[..., 0] / Xj[..., 2]) + cx
                vj = fy * (Xj[..., 1] / Xj[..., 2]) + cy
                dj = Xj[..., 3] / Xj[..., 2]
                u0 = np.floor(uj); v0 = np.floor(vj)
                ok = (u0 >= 0) & (v0 >= 0) & (u0 < wd - 1) & (v0 < ht - 1) & np.isfinite(uj) & np.isfinite(vj)
                u0c = np.clip(np.nan_to_num(u0), 0, wd - 2).astype(int)
                v0c = np.clip(np.nan_to_num(v0), 0, ht - 2).astype(int)
                hit = np.zeros((ht, wd), bool)
                for (a, c) in ((0, 0), (0, 1), (1, 0), (1, 1)):
                    dd = disps[jx][v0c + a, u0c + c]
                    hit |= np.abs(1.0 / dj - 1.0 / dd) < thresh[b]
            out[b] += (ok & hit)
    return out


def cvx_upsample(data, mask, pw=1.0, half_weights=False):
    """data [K,ht,wd] , mask [K,576,ht,wd] -> [K,8ht,8wd]  (utils/flow_viz.py:166-183)"""
    K, ht, wd = data.shape
    m = mask.astype(np.float64).reshape(K, 9, 8, 8, ht, wd).copy()
    m[:, [0, 1, 2], :, :, 0, :] = -np.inf
    m[:, [6, 7, 8], :, :, ht - 1, :] = -np.inf
    m[:, [0, 3, 6], :, :, :, 0] = -np.inf
    m[:, [2, 5, 8], :, :, :, wd - 1] = -np.inf
    m = m - m.max(axis=1, keepdims=True)
    e = np.exp(m)
    sm = e / e.sum(axis=1, keepdims=True)
    if half_weights:
        sm = sm.astype(np.float32).astype(np.float16).astype(np.float64)
    sm = sm ** pw
    pad = np.pad(data.astype(np.float64), ((0, 0), (1, 1), (1, 1)))
    nb = np.stack([pad[:, dy:dy + ht, dx:dx + wd] for dy in range(3) for dx in range(3)], 1)  # [K,9,ht,wd]
    up = (sm * nb[:, :, None, None]).sum(1)  # [K,8,8,ht,wd]
    return up.transpose(0, 3, 1, 4, 2).reshape(K, 8 * ht, 8 * wd)
