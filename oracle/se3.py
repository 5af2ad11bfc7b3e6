"""ORACLE (test infrastructure only — never imported by the product path).

SE3 / quaternion helpers restating the device functions of the reference's
src/droid_kernels.cu:66-188,994-1012 (actSO3, actSE3, adjSE3, relSE3, expSO3, expSE3, retrSE3),
vectorised in numpy.  The reference's lietorch dependency is absent from /root/reference
(empty submodule, princeton-vl/lietorch@e7df8655); the formulas duplicated in droid_kernels.cu are
the ones used here.  Pose layout [tx ty tz qx qy qz qw].
"""
import numpy as np


def act_so3(q, X):
    """src/droid_kernels.cu:66-76. q [...,4], X [...,3]"""
    qx, qy, qz, qw = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    x, y, z = X[..., 0], X[..., 1], X[..., 2]
    ux = 2.0 * (qy * z - qz * y)
    uy = 2.0 * (qz * x - qx * z)
    uz = 2.0 * (qx * y - qy * x)
    return np.stack([x + qw * ux + (qy * uz - qz * uy),
                     y + qw * uy + (qz * ux - qx * uz),
                     z + qw * uz + (qx * uy - qy * ux)], axis=-1)


def act_se3(t, q, X):
    """src/droid_kernels.cu:78-85. X [...,4] homogeneous (x,y,z,d)"""
    Y = act_so3(q, X[..., :3]) + X[..., 3:4] * t
    return np.concatenate([Y, X[..., 3:4]], axis=-1)


def adj_se3(t, q, X):
    """src/droid_kernels.cu:87-104: Y = Ad(G)^T X for twists ordered [tau, phi]. X [...,6]"""
    qinv = q * np.array([-1, -1, -1, 1], dtype=q.dtype)
    Ya = act_so3(qinv, X[..., 0:3])
    Yb = act_so3(qinv, X[..., 3:6])
    u = np.stack([t[..., 2] * X[..., 1] - t[..., 1] * X[..., 2],
                  t[..., 0] * X[..., 2] - t[..., 2] * X[..., 0],
                  t[..., 1] * X[..., 0] - t[..., 0] * X[..., 1]], axis=-1)
    return np.concatenate([Ya, Yb + act_so3(qinv, u)], axis=-1)


def rel_se3(ti, qi, tj, qj):
    """src/droid_kernels.cu:107-120: G_ij = G_j * G_i^-1"""
    q = np.stack([
        -qj[..., 3] * qi[..., 0] + qj[..., 0] * qi[..., 3] - qj[..., 1] * qi[..., 2] + qj[..., 2] * qi[..., 1],
        -qj[..., 3] * qi[..., 1] + qj[..., 1] * qi[..., 3] - qj[..., 2] * qi[..., 0] + qj[..., 0] * qi[..., 2],
        -qj[..., 3] * qi[..., 2] + qj[..., 2] * qi[..., 3] - qj[..., 0] * qi[..., 1] + qj[..., 1] * qi[..., 0],
        qj[..., 3] * qi[..., 3] + qj[..., 0] * qi[..., 0] + qj[..., 1] * qi[..., 1] + qj[..., 2] * qi[..., 2],
    ], axis=-1)
    t = tj - act_so3(q, ti)
    return t, q


def quat_mul(a, b):
    return np.stack([
        a[..., 3] * b[..., 0] + a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
        a[..., 3] * b[..., 1] + a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
        a[..., 3] * b[..., 2] + a[..., 2] * b[..., 3] + a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0],
        a[..., 3] * b[..., 3] - a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2],
    ], axis=-1)


def exp_so3(phi):
    """src/droid_kernels.cu:123-145"""
    th2 = (phi ** 2).sum(-1)
    th = np.sqrt(th2)
    small = th2 < 1e-8
    ths = np.where(small, 1.0, th)
    imag = np.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0, np.sin(0.5 * ths) / ths)
    real = np.where(small, 1.0 - th2 / 8.0 + th2 * th2 / 384.0, np.cos(0.5 * ths))
    return np.concatenate([imag[..., None] * phi, real[..., None]], axis=-1)


def exp_se3(xi):
    """src/droid_kernels.cu:160-188, xi = [tau, phi] -> (t, q)"""
    tau, phi = xi[..., :3], xi[..., 3:]
    q = exp_so3(phi)
    th2 = (phi ** 2).sum(-1)
    th = np.sqrt(th2)
    big = th > 1e-4
    th2s = np.where(big, th2, 1.0)
    ths = np.where(big, th, 1.0)
    a = np.where(big, (1 - np.cos(ths)) / th2s, 0.0)[..., None]
    b = np.where(big, (ths - np.sin(ths)) / (ths * th2s), 0.0)[..., None]
    c1 = np.cross(phi, tau)
    c2 = np.cross(phi, c1)
    return tau + a * c1 + b * c2, q


def retr_se3(xi, t, q):
    """src/droid_kernels.cu:994-1012: exp(xi) * (t, q)  (left retraction, DROID order)"""
    dt, dq = exp_se3(xi)
    return act_so3(dq, t) + dt, quat_mul(dq, q)


def inv_se3(t, q):
    qi = q * np.array([-1, -1, -1, 1], dtype=q.dtype)
    return -act_so3(qi, t), qi


def mul_se3(t1, q1, t2, q2):
    return act_so3(q1, t2) + t1, quat_mul(q1, q2)


def to_matrix(t, q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)
    T = np.zeros(q.shape[:-1] + (4, 4), dtype=q.dtype)
    T[..., :3, :3] = R
    T[..., :3, 3] = t
    T[..., 3, 3] = 1
    return T


def pose3_expmap(xi):
    """gtsam Pose3::Expmap, xi = [omega, v] -> (t, q). gtsam source is absent from /root/reference
    (ToniRV/gtsam-1@df2ac901); this is the published closed form t = V(omega) v."""
    w, v = xi[..., :3], xi[..., 3:]
    q = exp_so3(w)
    th2 = (w ** 2).sum(-1)
    th = np.sqrt(th2)
    big = th > 1e-10
    th2s = np.where(big, th2, 1.0)
    ths = np.where(big, th, 1.0)
    B = np.where(big, (1 - np.cos(ths)) / th2s, 0.5)[..., None]
    C = np.where(big, (ths - np.sin(ths)) / (ths * th2s), 1.0 / 6.0)[..., None]
    wv = np.cross(w, v)
    return v + B * wv + C * np.cross(w, wv), q


def pose3_retract(t, q, xi):
    """x.retract(xi) = x * Expmap(xi)  (right perturbation)"""
    dt, dq = pose3_expmap(xi)
    qn = quat_mul(q, dq)
    qn = qn / np.linalg.norm(qn, axis=-1, keepdims=True)
    return t + act_so3(q, dt), qn


def random_poses(rng, n, trans=0.1, rot_deg=5.0, dtype=np.float32):
    tau = rng.uniform(-1, 1, (n, 3))
    tau = tau / np.maximum(np.linalg.norm(tau, axis=-1, keepdims=True), 1e-9) * rng.uniform(0, trans, (n, 1))
    phi = rng.uniform(-1, 1, (n, 3))
    phi = phi / np.maximum(np.linalg.norm(phi, axis=-1, keepdims=True), 1e-9) * \
        rng.uniform(0, np.deg2rad(rot_deg), (n, 1))
    t, q = exp_se3(np.concatenate([tau, phi], -1))
    return np.concatenate([t, q], -1).astype(dtype)
