"""CPU tests of the oracle itself and of the host-side logic (no GPU)."""
import numpy as np
import pytest

from oracle import ba, corr, geom, se3
from tests.util import make_targets, make_window


def test_se3_group_laws(rng):
    p = se3.random_poses(rng, 4, 0.5, 40, np.float64)
    ti, qi = se3.inv_se3(p[:, :3], p[:, 3:])
    t, q = se3.mul_se3(p[:, :3], p[:, 3:], ti, qi)
    assert np.allclose(t, 0, atol=1e-12) and np.allclose(np.abs(q[:, 3]), 1, atol=1e-12)
    # rel(i,j) = Gj * Gi^-1
    tr, qr = se3.rel_se3(p[0, :3], p[0, 3:], p[1, :3], p[1, 3:])
    t2, q2 = se3.mul_se3(p[1, :3], p[1, 3:], ti[0], qi[0])
    assert np.allclose(tr, t2, atol=1e-12) and np.allclose(qr, q2, atol=1e-12)
    # exp/retract: small twist composes to first order
    xi = rng.normal(size=(3, 6)) * 1e-3
    t, q = se3.exp_se3(xi)
    assert np.allclose(t, xi[:, :3], atol=1e-5) and np.allclose(2 * q[:, :3], xi[:, 3:], atol=1e-5)


def test_adj_is_adjoint_transpose(rng):
    """Ad(G)^T as applied by adjSE3 satisfies <Ad^T X, xi> = <X, Ad xi> with Ad from exp conjugation"""
    p = se3.random_poses(rng, 1, 0.5, 40, np.float64)[0]
    X = rng.normal(size=(6,))
    xi = rng.normal(size=(6,)) * 1e-6
    # G exp(xi) G^-1 = exp(Ad xi)
    te, qe = se3.exp_se3(xi)
    ti, qi = se3.inv_se3(p[:3], p[3:])
    t1, q1 = se3.mul_se3(p[:3], p[3:], te, qe)
    t2, q2 = se3.mul_se3(t1, q1, ti, qi)
    ad_xi = np.concatenate([t2, 2 * q2[:3]])
    lhs = se3.adj_se3(p[:3], p[3:], X[None])[0] @ xi
    rhs = X @ ad_xi
    assert abs(lhs - rhs) < 1e-10


def test_lookup_is_bilinear_interpolation(rng):
    """A3 property: with fp32 volumes the 7x7 window equals dense bilinear sampling of the volume"""
    n, h1, w1, h2, w2, r = 2, 3, 4, 9, 11, 3
    vol = rng.normal(size=(n, h1, w1, h2, w2)).astype(np.float32)
    coords = np.stack([rng.uniform(-2, w2 + 1, (n, h1, w1)), rng.uniform(-2, h2 + 1, (n, h1, w1))], 1).astype(np.float32)
    out = corr.corr_index_forward(vol, coords, r)

    def sample(nn, y, x, px, py):
        x0, y0 = int(np.floor(px)), int(np.floor(py))
        acc = 0.0
        for (xx, yy, w) in ((x0, y0, (1 - (px - x0)) * (1 - (py - y0))), (x0 + 1, y0, (px - x0) * (1 - (py - y0))),
                            (x0, y0 + 1, (1 - (px - x0)) * (py - y0)), (x0 + 1, y0 + 1, (px - x0) * (py - y0))):
            if 0 <= xx < w2 and 0 <= yy < h2:
                acc += w * vol[nn, y, x, yy, xx]
        return acc
    for nn in range(n):
        for y in range(h1):
            for x in range(w1):
                for i in range(2 * r + 1):
                    for j in range(2 * r + 1):
                        ref = sample(nn, y, x, coords[nn, 0, y, x] + i - r, coords[nn, 1, y, x] + j - r)
                        assert abs(out[nn, i, j, y, x] - ref) < 1e-4


def test_altcorr_equals_volume_path(rng):
    """A4 == A2+A3 in fp32 (same maths, different order)"""
    E, C, H, W, r = 1, 16, 6, 8, 2
    f = rng.normal(size=(2, C, H, W)).astype(np.float32)
    f1, f2 = f[0:1], f[1:2]
    vol = np.einsum("ecm,ecn->emn", (f1 / 4).reshape(E, C, -1), (f2 / 4).reshape(E, C, -1)).reshape(E, H, W, H, W).astype(np.float32)
    coords = np.stack([rng.uniform(0, W, (E, H, W)), rng.uniform(0, H, (E, H, W))], 1).astype(np.float32)
    a = corr.corr_index_forward(vol, coords, r).reshape(E, -1, H, W)
    b = corr.altcorr_forward((f1 / 4).transpose(0, 2, 3, 1), (f2 / 4).transpose(0, 2, 3, 1),
                             coords.transpose(0, 2, 3, 1)[:, None], r)[:, 0]
    assert np.allclose(a, b, atol=1e-4)


def test_pyramid_shapes_and_floor():
    rng = np.random.default_rng(0)
    f = rng.normal(size=(1, 8, 5, 7)).astype(np.float16)
    pyr = corr.corr_volume_pyramid(f, f)
    assert [p.shape for p in pyr] == [(1, 5, 7, 5, 7), (1, 5, 7, 2, 3), (1, 5, 7, 1, 1), (1, 5, 7, 0, 0)]


def test_ba_hessian_structure(rng):
    poses, disps, intr, ii, jj = make_window(rng, 5, 8, 10, extra_edges=2)
    target, weight = make_targets(rng, poses, disps, intr, ii, jj)
    ext = np.array([0, 0, 0, 0, 0, 0, 1], np.float32)
    eta = rng.uniform(1e-3, 1e-1, (5, 8, 10)).astype(np.float32)
    r = ba.reduced_camera_matrix(poses, disps, intr, ext, np.zeros_like(disps), target, weight, eta, ii, jj, 0, 5)
    H, A, S = r["H"], r["A"], r["S"]
    assert np.allclose(A, A.T, atol=1e-9) and np.allclose(S, S.T, atol=1e-9)
    assert np.linalg.eigvalsh(A).min() > -1e-8
    assert np.linalg.eigvalsh(H + 1e-6 * np.eye(30)).min() > -1e-6   # Schur complement stays PSD
    # Schur complement == elimination of the depth block of the full system
    dx, L = ba.dense_solve(H, r["v"], prior_idx=0, prior_err=np.zeros(6), prior_info=1e8)
    res = (H + np.pad(1e8 * np.eye(6), ((0, 24), (0, 24)))) @ dx.reshape(-1) - r["v"]
    assert np.abs(res).max() < 1e-6 * max(1.0, np.abs(r["v"]).max())


def test_ba_gauss_newton_reduces_residual(rng):
    """two BA iterations pull the reprojection towards the (noisy) targets"""
    poses, disps, intr, ii, jj = make_window(rng, 5, 8, 10, extra_edges=2)
    target, weight = make_targets(rng, poses, disps, intr, ii, jj, noise=0.05)
    # perturb the state
    poses2 = poses.copy()
    poses2[1:, :3] += rng.normal(0, 0.01, (4, 3)).astype(np.float32)
    ext = np.array([0, 0, 0, 0, 0, 0, 1], np.float32)
    eta = np.full((5, 8, 10), 1e-4, np.float32)

    def cost(p, d):
        c, _ = geom.reproject(p, d, intr, ii, jj)
        return (weight.transpose(0, 2, 3, 1) * (c - target.transpose(0, 2, 3, 1)) ** 2).sum()
    wTb = np.stack([np.concatenate(se3.inv_se3(p[:3].astype(np.float64), p[3:].astype(np.float64))) for p in poses2])
    cam = poses2.astype(np.float64)
    d = disps.astype(np.float64)
    c0 = cost(cam, d)
    for _ in range(3):
        r = ba.reduced_camera_matrix(cam, d, intr, ext, np.zeros_like(disps), target, weight, eta, ii, jj, 0, 5)
        err = ba.pose_prior_error(wTb[0], wTb[0])
        dx, _ = ba.dense_solve(r["H"], r["v"], 0, err, 1e8)
        wTb, cam = ba.gtsam_retract(wTb, ext, dx, 0)
        d, _ = ba.solve_depth(dx, d, r["Q"], r["E"], r["w"], ii, jj, 0, 5)
        d = np.maximum(d, 1e-3)
    assert cost(cam, d) < 0.5 * c0


def test_graph_builder_matches_bruteforce(rng):
    from nerf_slam_b200.ba_graph import BAGraphHost
    ii = np.array([3, 4, 5, 6, 4, 5, 2, 6, 3], np.int64)
    jj = np.array([4, 3, 4, 5, 6, 3, 3, 2, 5], np.int64)
    kf0, kf1 = 3, 7
    g = BAGraphHost(ii, jj, kf0, kf1)
    t = g.tables
    assert list(t["kx"]) == [2, 3, 4, 5, 6] and g.K == 5 and g.P == 4
    # every edge appears once under its source
    for k in range(g.K):
        es = t["src_edges"][t["src_ptr"][k]:t["src_ptr"][k + 1]]
        assert all(ii[e] == t["kx"][k] for e in es)
    assert sorted(t["src_edges"]) == list(range(len(ii)))
    # rows: self rows for frames in the window + edges with target in window
    nrows = sum(1 for f in t["kx"] if kf0 <= f < kf1) + sum(1 for j in jj if kf0 <= j < kf1)
    assert g.NR == nrows
    # dense assembly tables reproduce the brute-force block pattern
    pat = np.zeros((g.P, g.P), int)
    for a, b in zip(ii - kf0, jj - kf0):
        for (x, y) in ((a, a), (a, b), (b, a), (b, b)):
            if 0 <= x < g.P and 0 <= y < g.P:
                pat[x, y] += 1
    cnt = np.diff(t["hc_ptr"]).reshape(g.P, g.P)
    hs_cnt = np.zeros_like(pat)
    for blk in range(g.P * g.P):
        ids = t["hc_idx"][t["hc_ptr"][blk]:t["hc_ptr"][blk + 1]]
        hs_cnt[blk // g.P, blk % g.P] = (ids >= 0).sum()
    assert (hs_cnt == pat).all() and (cnt >= pat).all()


def test_cvx_upsample_partition_of_unity(rng):
    K, ht, wd = 2, 5, 6
    mask = rng.normal(size=(K, 576, ht, wd)).astype(np.float32)
    ones = geom.cvx_upsample(np.ones((K, ht, wd)), mask)
    assert np.allclose(ones, 1.0, atol=1e-12)   # convex combination of valid neighbours only


def test_frame_distance_f32_tree_close_to_f64(rng):
    poses, disps, intr, ii, jj = make_window(rng, 4, 16, 20, extra_edges=0)
    a = geom.frame_distance(poses, disps, intr, ii, jj, 0.3, np.float32)
    b = geom.frame_distance(poses, disps, intr, ii, jj, 0.3, np.float64)
    assert np.allclose(a, b, rtol=1e-4)


def test_row_pair_tiling_of_the_correlation_pyramid(rng):
    """index/rounding emulation of csrc/corr_volume_rows.cu (tile = two full target rows; level 1 from the tile,
    levels 2/3 from register carries across 2/4 consecutive row pairs, column halves per warp) against the
    oracle pyramid: locks the offsets and the carry logic of that kernel on the CPU."""
    from oracle import corr as ocorr
    H, W, C, E = 12, 16, 8, 2                      # W % 16 == 0, H even; H/2 = 6 row pairs (last quad incomplete)
    f1 = rng.normal(0, 1, (E, C, H, W)).astype(np.float16)
    f2 = rng.normal(0, 1, (E, C, H, W)).astype(np.float16)
    ref = ocorr.corr_volume_pyramid(f1, f2)
    HW = H * W
    out = [np.full((E, HW, (H >> l) * (W >> l)), np.nan, np.float16) for l in range(4)]
    f16 = np.float16
    pool = lambda a, b, c, d: f16((((np.float32(a) + np.float32(b)) + np.float32(c)) + np.float32(d)) * np.float32(0.25))
    WH = W // 2
    for e in range(E):
        a = f1[e].reshape(C, HW).T.astype(np.float32)          # [HW, C]
        b = f2[e].reshape(C, HW).T.astype(np.float32)
        vol = (a @ b.T) * np.float32(0.0625)                   # fp32 accumulate, scaled
        for m in range(HW):                                    # thread = source pixel
            for hsel in range(2):                              # column half of the warp
                c0 = hsel * WH
                l1p = l2p = None
                for rp in range(H // 2):
                    h0 = vol[m, 2 * rp * W + c0:2 * rp * W + c0 + WH].astype(f16)
                    h1 = vol[m, (2 * rp + 1) * W + c0:(2 * rp + 1) * W + c0 + WH].astype(f16)
                    out[0][e, m, 2 * rp * W + c0:2 * rp * W + c0 + WH] = h0
                    out[0][e, m, (2 * rp + 1) * W + c0:(2 * rp + 1) * W + c0 + WH] = h1
                    l1 = np.array([pool(h0[2 * x], h0[2 * x + 1], h1[2 * x], h1[2 * x + 1]) for x in range(WH // 2)], f16)
                    out[1][e, m, rp * (W // 2) + c0 // 2:rp * (W // 2) + c0 // 2 + WH // 2] = l1
                    if rp & 1:
                        l2 = np.array([pool(l1p[2 * y], l1p[2 * y + 1], l1[2 * y], l1[2 * y + 1]) for y in range(WH // 4)], f16)
                        q2 = rp >> 1
                        if q2 < (H >> 2):
                            out[2][e, m, q2 * (W // 4) + c0 // 4:q2 * (W // 4) + c0 // 4 + WH // 4] = l2
                        if (rp & 3) == 3:
                            q3 = rp >> 2
                            if q3 < (H >> 3):
                                l3 = np.array([pool(l2p[2 * z], l2p[2 * z + 1], l2[2 * z], l2[2 * z + 1]) for z in range(WH // 8)], f16)
                                out[3][e, m, q3 * (W // 8) + c0 // 8:q3 * (W // 8) + c0 // 8 + WH // 8] = l3
                        else:
                            l2p = l2
                    else:
                        l1p = l1
    for l in range(4):
        got = out[l].reshape(E, H, W, H >> l, W >> l).astype(np.float32)
        assert not np.isnan(got).any(), l
        assert np.abs(got - ref[l].astype(np.float32)).max() <= 2e-2, l       # same tolerance as the GPU parity test


def test_warp_chunked_march_equals_serial_lattice_march(rng):
    """the kernel tests 32 lattice points per iteration (ballot), truncates at max_n with the n-th set bit and replays
    the kept masks; emulated here in Python against the serial oracle (oracle/ngp.py::march_lattice)"""
    from oracle import ngp as ongp
    cascades = 3
    ncell = ongp.GRID ** 3 * cascades
    for trial in range(6):
        bits = (rng.random(ncell // 8) < [0.02, 0.3, 1.0][trial % 3]).astype(np.uint8) * rng.integers(1, 256, ncell // 8).astype(np.uint8)
        o = rng.uniform(0.3, 0.7, 3).astype(np.float32)
        d = rng.normal(size=3).astype(np.float32); d /= np.linalg.norm(d)
        max_n = [1024, 37, 5][trial % 3]
        jitter = float(rng.random())
        ref = ongp.march_lattice(o, d, -1.5, 2.5, 0.05, 1.0 / 256, cascades, bits, jitter, max_n)
        # ---- warp emulation
        inv = np.float32(1.0) / d
        tmin, tmax = np.float32(-1e30), np.float32(1e30)
        for a in range(3):
            t0, t1 = (np.float32(-1.5) - o[a]) * inv[a], (np.float32(2.5) - o[a]) * inv[a]
            tmin = max(tmin, min(t0, t1)); tmax = min(tmax, max(t0, t1))
        tbase = np.float32(max(tmin, np.float32(0.05)) + np.float32(1e-6))
        tbase = np.float32(tbase + ongp.calc_dt(tbase, 1.0 / 256) * np.float32(jitter))
        n, masks, bases = 0, [], []
        for ci in range(128):
            if not (tbase < tmax) or n >= max_n:
                break
            ts = []
            t = tbase
            for lane in range(32):
                ts.append(t); t = np.float32(t + ongp.calc_dt(t, 1.0 / 256))
            m = 0
            for lane in range(32):
                tl = ts[lane]
                if tl < tmax:
                    dt = ongp.calc_dt(tl, 1.0 / 256)
                    p = [np.float32(d[a] * tl + o[a]) for a in range(3)]
                    if ongp.occupied(p, ongp.mip_from_dt(dt, p, cascades), bits):
                        m |= 1 << lane
            c = bin(m).count("1")
            if n + c > max_n:                                   # keep the first max_n - n set bits (__fns)
                keep, mm, cnt = max_n - n, 0, 0
                for lane in range(32):
                    if (m >> lane) & 1:
                        if cnt == keep:
                            break
                        mm |= 1 << lane; cnt += 1
                m = mm
            n += bin(m).count("1")
            masks.append(m); bases.append(tbase)
            tbase = np.float32(ts[31] + ongp.calc_dt(ts[31], 1.0 / 256))
        got = []
        for m, tb in zip(masks, bases):
            t = tb
            for lane in range(32):
                if (m >> lane) & 1:
                    got.append((t, ongp.calc_dt(t, 1.0 / 256)))
                t = np.float32(t + ongp.calc_dt(t, 1.0 / 256))
        assert len(got) == len(ref) == n
        assert all(a[0] == b[0] and a[1] == b[1] for a, b in zip(got, ref))


def test_blocked_cholesky_solve_and_inverse_algorithm(rng):
    """numpy emulation of csrc/ba.cu::ba_solve_kernel: Cholesky blocked by the 6x6 pose blocks (diagonal block
    factored first, panel rows solved against it, rank-6 trailing update), blocked forward/back substitution and
    the column-wise triangular inverse (stored in the upper triangle) — against numpy.linalg"""
    P = 7
    n = 6 * P
    M = rng.normal(size=(n, n))
    H = M @ M.T + n * np.eye(n)
    v = rng.normal(size=n)
    A = H.copy()
    diag = np.zeros(n)
    for jb in range(P):
        j0 = 6 * jb
        Lb = np.tril(A[j0:j0 + 6, j0:j0 + 6]).copy()
        rd = np.zeros(6)
        for c in range(6):
            dd = Lb[c, c] - Lb[c, :c] @ Lb[c, :c]
            assert dd > 0
            l = np.sqrt(dd); rd[c] = 1.0 / l; Lb[c, c] = l
            for r in range(c + 1, 6):
                Lb[r, c] = (Lb[r, c] - Lb[r, :c] @ Lb[c, :c]) * rd[c]
        for i in range(j0 + 6, n):                               # panel: row_i(L) = row_i(A) Ljj^-T
            x = np.zeros(6)
            for c in range(6):
                x[c] = (A[i, j0 + c] - x[:c] @ Lb[c, :c]) * rd[c]
            A[i, j0:j0 + 6] = x
        A[j0:j0 + 6, j0:j0 + 6] = np.tril(Lb) + np.triu(A[j0:j0 + 6, j0:j0 + 6], 1)
        diag[j0:j0 + 6] = rd
        for i in range(j0 + 6, n):                               # trailing update (lower triangle)
            for c in range(j0 + 6, i + 1):
                A[i, c] -= A[i, j0:j0 + 6] @ A[c, j0:j0 + 6]
    L = np.tril(A)
    assert np.allclose(L, np.linalg.cholesky(H), rtol=1e-10, atol=1e-10)
    y = v.copy()
    for jb in range(P):                                          # L z = y
        j0 = 6 * jb
        for c in range(6):
            y[j0 + c] = (y[j0 + c] - A[j0 + c, j0:j0 + c] @ y[j0:j0 + c]) * diag[j0 + c]
        for i in range(j0 + 6, n):
            y[i] -= A[i, j0:j0 + 6] @ y[j0:j0 + 6]
    for jb in range(P - 1, -1, -1):                              # L^T x = z
        j0 = 6 * jb
        for c in range(5, -1, -1):
            y[j0 + c] = (y[j0 + c] - A[j0 + c + 1:j0 + 6, j0 + c] @ y[j0 + c + 1:j0 + 6]) * diag[j0 + c]
        for i in range(j0):
            y[i] -= A[j0:j0 + 6, i] @ y[j0:j0 + 6]
    assert np.allclose(y, np.linalg.solve(H, v), rtol=1e-9, atol=1e-10)
    Linv = np.zeros((n, n))
    for c in range(n):                                           # X = L^-1, column by column, X in the upper triangle
        xcc = diag[c]
        Linv[c, c] = xcc
        for i in range(c + 1, n):
            s = -A[i, c] * xcc
            for k in range(c + 1, i):
                s -= A[i, k] * A[c, k]                           # L[i][k] * X[k][c]
            x = s * diag[i]
            A[c, i] = x
            Linv[i, c] = x
    assert np.allclose(Linv, np.linalg.inv(np.linalg.cholesky(H)), rtol=1e-8, atol=1e-10)
