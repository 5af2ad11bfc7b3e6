"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement (numpy) of the per-voxel update of TsdfFusion.custom_volume_integrate, reference
fusion/tsdf_fusion.py:231-296 (and of the pixel masking / weighting of build_volume, :185-203), applied to EVERY voxel of a
dense grid.  Open3D's block activation is absent from /root/reference — parity of the ACTIVE SET is unpinned; the update rule
itself is pinned: tests/golden/ref_tsdf_integrate.npz holds the output of the reference's own `build_volume` +
`custom_volume_integrate` executed verbatim on torch stand-ins for Open3D's tensors (tests/golden/make_golden_tsdf.py), and
this restatement follows it operation for operation: float32 voxel coordinates and pose matrix promoted to fp64 for the
projection, Tensor.round(), fp32 sdf / weights / running averages (tests/test_cpu_golden.py)."""
import numpy as np

F32 = np.float32


def pose_tq_to_matrix(tq):
    x, y, z, w = [float(v) for v in tq[3:]]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(tq[:3], np.float64)
    return T.astype(np.float32).astype(np.float64)      # SE3(poses).matrix() is fp32; build_volume promotes it (:209)


def integrate(tsdf, weight, color, origin, voxel_size, idepth_up, depth_cov_up, rgb_chw, intr, cam_T_world_tq,
              max_depth=6.0, sdf_trunc=0.10, max_weight=20.0, max_depth_sigma=10000.0):
    """in place on tsdf / weight [nz,ny,nx] fp32 and color [nz,ny,nx,3] fp32"""
    nz, ny, nx = tsdf.shape
    H, W = idepth_up.shape
    # build_volume (:192-203): depth, weights, mask
    depth = (F32(1.0) / idepth_up.astype(F32)).astype(F32)
    if depth_cov_up is None:
        wimg = np.ones((H, W), F32)
    else:
        cov = depth_cov_up.astype(F32)
        wimg = np.sqrt(F32(1.0) / cov).astype(F32)
        depth = np.where(np.sqrt(cov) < F32(max_depth_sigma), depth, F32(max_depth + 1.0)).astype(F32)
    T = pose_tq_to_matrix(np.asarray(cam_T_world_tq, np.float32))
    fx, fy, cx, cy = [float(F32(v)) for v in intr]
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    # voxel coordinates are float32 tensors in the reference (Open3D hands out float32 metric coordinates), promoted to
    # float64 for the projection (:246): float32 product, float32 sum
    org = np.asarray(origin, np.float32)
    vs = F32(voxel_size)
    wx = (vs * ix.astype(F32) + org[0]).astype(F32).astype(np.float64)
    wy = (vs * iy.astype(F32) + org[1]).astype(F32).astype(np.float64)
    wz = (vs * iz.astype(F32) + org[2]).astype(F32).astype(np.float64)
    x = T[0, 0] * wx + T[0, 1] * wy + T[0, 2] * wz + T[0, 3]
    y = T[1, 0] * wx + T[1, 1] * wy + T[1, 2] * wz + T[1, 3]
    d = T[2, 0] * wx + T[2, 1] * wy + T[2, 2] * wz + T[2, 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = np.rint((fx * x + cx * d) / d)
        v = np.rint((fy * y + cy * d) / d)
    proj = (d > 0) & (u >= 0) & (v >= 0) & (u < W) & (v < H)
    ui = np.where(proj, u, 0).astype(np.int64); vi = np.where(proj, v, 0).astype(np.int64)
    dr = depth[vi, ui]
    sdf = (dr - d.astype(F32)).astype(F32)
    inl = proj & (dr > 0) & (dr < F32(max_depth)) & (sdf >= -F32(sdf_trunc))
    sdf = (np.minimum(sdf, F32(sdf_trunc)) / F32(sdf_trunc)).astype(F32)
    wr = wimg[vi, ui]
    w = weight
    wp = (w + wr).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        new_t = ((w * tsdf + wr * sdf) / wp).astype(F32)
        rgb = rgb_chw.astype(F32)
        new_c = ((w[..., None] * color + wr[..., None] * np.stack([rgb[c][vi, ui] for c in range(3)], -1)) / wp[..., None]).astype(F32)
    tsdf[inl] = new_t[inl]
    color[inl] = new_c[inl]
    weight[inl] = np.minimum(wp, F32(max_weight))[inl]
    return int(inl.sum())
