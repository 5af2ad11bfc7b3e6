"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp64 restatement of the reference's dense bundle adjustment:
  A7   projective_transform_kernel            src/droid_kernels.cu:192-536
  A8   accum_cuda / accum_kernel              src/droid_kernels.cu:971-991,1065-1115
  A9   schur_block / EEt6x6 / Ev6x1           src/droid_kernels.cu:1118-1210,1349-1438
  A10  SparseBlock (dense here)               src/droid_kernels.cu:1240-1346
  A11  reduced_camera_matrix_cuda             src/droid_kernels.cu:1681-1768
  A12  dense solve + gtsam retract            slam/visual_frontends/visual_frontend.py:1097-1162
  A13  solve_depth_cuda / EvT6x1 / disp_retr  src/droid_kernels.cu:1213-1238,1050-1063,1772-1825
  A14  covariances                            slam/visual_frontends/visual_frontend.py:1164-1230
  A15  ba_cuda                                src/droid_kernels.cu:1441-1568
The per-pixel Jacobian chain is evaluated exactly as the kernel does (adjSE3 through G_ij, then
through cam_T_body, two sign flips, reorder to [omega, t]) — NOT with the precomputed 6x6 maps the
CUDA product path uses — so it independently checks that restructuring.
gtsam (ToniRV/gtsam-1@df2ac901) is absent from /root/reference: the solve/retract boundary is
"parity unpinned"; pinned instead by H dx = v residuals and the closed-form Pose3 retraction.
"""
import numpy as np

from . import se3

MIN_DEPTH = 0.25


def linearize_edge(target, weight, poses, disps, intr, ext, i, j, dtype=np.float64):
    """one edge of projective_transform_kernel.
    target, weight: [2,ht,wd].  returns dict(Hs [4,6,6], vs [2,6], Eiz [6,HW], Ejz [6,HW], Cii [HW], bz [HW])"""
    fx, fy, cx, cy = [dtype(x) for x in intr]
    ht, wd = disps.shape[1:]
    hw = ht * wd
    v, u = np.meshgrid(np.arange(ht, dtype=dtype), np.arange(wd, dtype=dtype), indexing="ij")
    u, v = u.reshape(-1), v.reshape(-1)
    P = poses.astype(dtype)
    if i == j:
        tij = np.array([-0.1, 0, 0], dtype); qij = np.array([0, 0, 0, 1], dtype)
    else:
        tij, qij = se3.rel_se3(P[i, :3], P[i, 3:], P[j, :3], P[j, 3:])
    ext = ext.astype(dtype)
    Xi = np.stack([(u - cx) / fx, (v - cy) / fy, np.ones(hw, dtype), disps[i].reshape(-1).astype(dtype)], -1)
    Xj = se3.act_se3(tij, qij, Xi)
    x, y, Z, h = Xj[:, 0], Xj[:, 1], Xj[:, 2], Xj[:, 3]
    bad = Z < MIN_DEPTH
    d = np.where(bad, 0.0, 1.0 / np.where(bad, 1.0, Z))
    d2 = d * d
    wu = np.where(bad, 0.0, 0.001 * weight[0].reshape(-1).astype(dtype))
    wv = np.where(bad, 0.0, 0.001 * weight[1].reshape(-1).astype(dtype))
    ru = target[0].reshape(-1).astype(dtype) - (fx * d * x + cx)
    rv = target[1].reshape(-1).astype(dtype) - (fy * d * y + cy)

    def chain(Jj):
        Ji = -se3.adj_se3(tij, qij, Jj)
        Jj2 = se3.adj_se3(ext[:3], ext[3:], Jj)
        Ji2 = se3.adj_se3(ext[:3], ext[3:], Ji)
        Jj2, Ji2 = -Jj2, -Ji2
        perm = [3, 4, 5, 0, 1, 2]
        return Ji2[:, perm], Jj2[:, perm]

    zero = np.zeros(hw, dtype)
    Jzu = fx * (tij[0] * d - tij[2] * (x * d2))
    Cii = wu * Jzu * Jzu
    bz = wu * ru * Jzu
    if i == j:
        wu = zero
    Jju = np.stack([fx * (h * d), zero, fx * (-x * h * d2), fx * (-x * y * d2), fx * (1.0 + x * x * d2), fx * (-y * d)], -1)
    Jiu, Jju = chain(Jju)
    Jzv = fy * (tij[1] * d - tij[2] * (y * d2))
    Cii = Cii + wv * Jzv * Jzv
    bz = bz + wv * rv * Jzv
    if i == j:
        wv = zero
    Jjv = np.stack([zero, fy * (h * d), fy * (-y * h * d2), fy * (-1 - y * y * d2), fy * (x * y * d2), fy * (x * d)], -1)
    Jiv, Jjv = chain(Jjv)

    Ju = np.concatenate([Jiu, Jju], -1)  # [hw,12]
    Jv = np.concatenate([Jiv, Jjv], -1)
    H = np.einsum("p,pa,pb->ab", wu, Ju, Ju) + np.einsum("p,pa,pb->ab", wv, Jv, Jv)
    b = np.einsum("p,pa->a", wu * ru, Ju) + np.einsum("p,pa->a", wv * rv, Jv)
    Hs = np.stack([H[:6, :6], H[:6, 6:], H[6:, :6], H[6:, 6:]])
    vs = np.stack([b[:6], b[6:]])
    Eiz = (wu * Jzu)[None] * Jiu.T + (wv * Jzv)[None] * Jiv.T
    Ejz = (wu * Jzu)[None] * Jju.T + (wv * Jzv)[None] * Jjv.T
    return dict(Hs=Hs, vs=vs, Eiz=Eiz, Ejz=Ejz, Cii=Cii, bz=bz)


def reduced_camera_matrix(poses, disps, intr, ext, disps_sens, targets, weights, eta, ii, jj, kf0, kf1,
                          dtype=np.float64):
    """-> dict(H [6P,6P], v [6P], Q [K,HW], E [P+M,6,HW], w [K,HW], Hs [4,M,6,6], vs [2,M,6], kx)"""
    ii = np.asarray(ii); jj = np.asarray(jj)
    M = len(ii)
    P = kf1 - kf0
    ht, wd = disps.shape[1:]
    hw = ht * wd
    ts = np.arange(kf0, kf1)
    ii_exp = np.concatenate([ts, ii]); jj_exp = np.concatenate([ts, jj])
    kx, kk_exp = np.unique(ii_exp, return_inverse=True)
    K = len(kx)
    lin = [linearize_edge(targets[m], weights[m], poses, disps, intr, ext, int(ii[m]), int(jj[m]), dtype)
           for m in range(M)]
    Hs = np.stack([l["Hs"] for l in lin], 1) if M else np.zeros((4, 0, 6, 6), dtype)
    vs = np.stack([l["vs"] for l in lin], 1) if M else np.zeros((2, 0, 6), dtype)
    n = 6 * P
    A = np.zeros((n, n), dtype); bA = np.zeros(n, dtype)
    for which, (ri, ci) in enumerate([(ii, ii), (ii, jj), (jj, ii), (jj, jj)]):
        for m in range(M):
            a, c = ri[m] - kf0, ci[m] - kf0
            if a >= 0 and c >= 0 and a < P and c < P:
                A[6 * a:6 * a + 6, 6 * c:6 * c + 6] += Hs[which, m]
    for which, ri in enumerate([ii, jj]):
        for m in range(M):
            a = ri[m] - kf0
            if 0 <= a < P:
                bA[6 * a:6 * a + 6] += vs[which, m]
    # depth blocks
    C = np.zeros((K, hw), dtype); w = np.zeros((K, hw), dtype)
    Ei = np.zeros((P, 6, hw), dtype)
    for m in range(M):
        k = np.searchsorted(kx, ii[m])
        C[k] += lin[m]["Cii"]; w[k] += lin[m]["bz"]
        if kf0 <= ii[m] < kf1:
            Ei[ii[m] - kf0] += lin[m]["Eiz"]
    alpha = 0.05
    ds = disps_sens[kx].reshape(K, hw).astype(dtype) if disps_sens is not None else np.zeros((K, hw), dtype)
    mm = (ds > 0).astype(dtype)
    C = C + mm * alpha + (1 - mm) * eta.reshape(K, hw).astype(dtype)
    w = w - mm * alpha * (disps[kx].reshape(K, hw).astype(dtype) - ds)
    Q = 1.0 / C
    Ejz = np.stack([l["Ejz"] for l in lin]) if M else np.zeros((0, 6, hw), dtype)
    E = np.concatenate([Ei, Ejz], 0)
    # Schur complement over expanded rows (schur_block)
    S = np.zeros((n, n), dtype); bS = np.zeros(n, dtype)
    rows = [(r, int(jj_exp[r] - kf0), int(kk_exp[r])) for r in range(P + M) if kf0 <= jj_exp[r] < kf1]
    for (ra, pa, ka) in rows:
        bS[6 * pa:6 * pa + 6] += (E[ra] * (Q[ka] * w[ka])[None]).sum(-1)
        for (rb, pb, kb) in rows:
            if ka == kb:
                S[6 * pa:6 * pa + 6, 6 * pb:6 * pb + 6] += (E[ra] * Q[ka][None]) @ E[rb].T
    return dict(H=A - S, v=bA - bS, Q=Q, E=E, w=w, Hs=Hs, vs=vs, kx=kx, A=A, S=S)


def dense_solve(H, v, prior_idx=-1, prior_err=None, prior_info=0.0, lm=0.0, ep=0.0):
    """(H + damping + prior) dx = v (+ prior), fp64 Cholesky. returns dx [P,6], L"""
    H = H.astype(np.float64).copy(); v = v.astype(np.float64).reshape(-1).copy()
    n = H.shape[0]
    dg = np.arange(n)
    H[dg, dg] += ep + lm * H[dg, dg]
    if prior_idx >= 0:
        s = slice(6 * prior_idx, 6 * prior_idx + 6)
        H[s, s] += prior_info * np.eye(6)
        v[s] -= prior_info * np.asarray(prior_err, np.float64)
    L = np.linalg.cholesky(H)
    y = np.linalg.solve(L, v)
    dx = np.linalg.solve(L.T, y)
    return dx.reshape(-1, 6), L


def solve_depth(dx, disps, Q, E, w, ii, jj, kf0, kf1):
    """in-place on a copy; returns new disps. Keeps the `idx <= 0` skip (src/droid_kernels.cu:1225)."""
    ii = np.asarray(ii); jj = np.asarray(jj)
    P = kf1 - kf0
    ts = np.arange(kf0, kf1)
    ii_exp = np.concatenate([ts, ii]); jj_exp = np.concatenate([ts, jj])
    kx, kk_exp = np.unique(ii_exp, return_inverse=True)
    K = len(kx)
    hw = Q.shape[1]
    dw = np.zeros((K, hw), np.float64)
    for r in range(len(ii_exp)):
        idx = jj_exp[r] - kf0
        if idx <= 0 or idx >= P:
            continue
        dw[kk_exp[r]] += E[r].T.astype(np.float64) @ dx[idx].astype(np.float64)
    dz = Q * (w - dw)
    out = disps.astype(np.float64).copy()
    out[kx] += dz.reshape(K, *disps.shape[1:])
    return out, dz


def gtsam_retract(world_T_body, cam_T_body, dx, kf0):
    """world_T_body[kf0+i] <- world_T_body * Expmap(dx[i]) ; cam_T_world = cam_T_body * world_T_body^-1"""
    wTb = world_T_body.astype(np.float64).copy()
    P = dx.shape[0]
    t, q = se3.pose3_retract(wTb[kf0:kf0 + P, :3], wTb[kf0:kf0 + P, 3:], dx.astype(np.float64))
    wTb[kf0:kf0 + P, :3] = t; wTb[kf0:kf0 + P, 3:] = q
    w32 = wTb.astype(np.float32).astype(np.float64)
    ti, qi = se3.inv_se3(w32[:, :3], w32[:, 3:])
    tc, qc = se3.mul_se3(cam_T_body[None, :3].astype(np.float64), cam_T_body[None, 3:].astype(np.float64), ti, qi)
    return wTb, np.concatenate([tc, qc], -1)


def covariances_reference(L, E, Q, ii, jj, kf0, kf1, disps):
    """What the reference's covariance block REALLY computes (visual_frontend.py:1171-1230, pinned by
    tests/golden/ref_covariances.npz): as `covariances` below, except that its assignment
    `Ej[range(P), kf0-min:kf1-min, :, :] = Ei[range(P), :, :]` (:1214) broadcasts over the pose dimension — for a depth
    map of an OPTIMISED frame q, every pose row p holds Ei[q] and the Ejz blocks of that column are overwritten; depth maps
    of fixed frames keep their Ejz blocks.  Depth maps = unique(ii) (:1201), assumed contiguous from min(ii, jj)."""
    ii = np.asarray(ii); jj = np.asarray(jj)
    P = kf1 - kf0
    n = 6 * P
    Linv = np.linalg.solve(L, np.eye(n))
    sig = Linv.T @ Linv
    sigma_g = np.stack([sig[6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(P)])
    kx = np.unique(ii)
    K = len(kx)
    hw = Q.shape[1]
    z_cov = np.zeros((K, hw))
    for k in range(K):
        x = np.zeros((n, hw))
        if kf0 <= kx[k] < kf1:
            for p in range(P):
                x[6 * p:6 * p + 6] = E[kx[k] - kf0]
        else:
            for m in range(len(ii)):
                if ii[m] == kx[k] and kf0 <= jj[m] < kf1:
                    p = jj[m] - kf0
                    x[6 * p:6 * p + 6] = E[P + m]
        F = (Q[k][:, None] * x.T) @ Linv
        z_cov[k] = Q[k] + (F ** 2).sum(-1)
    d = disps[kx].reshape(K, hw).astype(np.float64)
    return sigma_g, z_cov, z_cov / d ** 4


def covariances(L, E, Q, ii, jj, kf0, kf1, disps):
    """The INTENDED formula of visual_frontend.py:1171-1230 (Ei on the diagonal, Ejz off the diagonal) — what
    csrc/ba.cu::nslam_ba_cov computes; it equals the reference's output for the depth maps of fixed frames only, see
    covariances_reference.  returns sigma_g [P,6,6], z_cov [K,HW], depth_cov [K,HW]"""
    ii = np.asarray(ii); jj = np.asarray(jj)
    P = kf1 - kf0
    n = 6 * P
    Linv = np.linalg.solve(L, np.eye(n))
    sig = Linv.T @ Linv
    sigma_g = np.stack([sig[6 * i:6 * i + 6, 6 * i:6 * i + 6] for i in range(P)])
    ts = np.arange(kf0, kf1)
    kx = np.unique(np.concatenate([ts, ii]))
    K = len(kx)
    hw = Q.shape[1]
    z_cov = np.zeros((K, hw))
    for k in range(K):
        x = np.zeros((n, hw))
        for m in range(len(ii)):
            if ii[m] == kx[k] and kf0 <= jj[m] < kf1:
                p = jj[m] - kf0
                x[6 * p:6 * p + 6] = E[P + m]
        if kf0 <= kx[k] < kf1:
            p = kx[k] - kf0
            x[6 * p:6 * p + 6] = E[p]
        F = (Q[k][:, None] * x.T) @ Linv
        z_cov[k] = Q[k] + (F ** 2).sum(-1)
    d = disps[kx].reshape(K, hw).astype(np.float64)
    return sigma_g, z_cov, z_cov / d ** 4


def pose_prior_error(x, prior):
    """Logmap(prior^-1 * x) in [omega, t] (gtsam Pose3::Logmap)"""
    x = x.astype(np.float64); prior = prior.astype(np.float64)
    ti, qi = se3.inv_se3(prior[:3], prior[3:])
    t, q = se3.mul_se3(ti, qi, x[:3], x[3:])
    if q[3] < 0:
        q = -q
    sn = np.linalg.norm(q[:3])
    if sn < 1e-12:
        w = 2 * q[:3]; th = 0.0
    else:
        th = 2 * np.arctan2(sn, q[3]); w = th / sn * q[:3]
    if th < 1e-5:
        c = 1.0 / 12
    else:
        c = (1 - th * np.sin(th) / (2 * (1 - np.cos(th)))) / th ** 2
    wt = np.cross(w, t)
    u = t - 0.5 * wt + c * np.cross(w, wt)
    return np.concatenate([w, u])
