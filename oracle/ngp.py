"""ORACLE (test infrastructure only — never imported by the product path).

Path B: PARITY UNPINNED.  The reference's NeRF back-end is the `pyngp` module of the fork
ToniRV/instant-ngp@54aba7cfbeaf6a60f29469a9938485bebeba24c3 (branch feature/nerf_slam, with
tiny-cuda-nn as its submodule); that submodule is EMPTY under /root/reference/thirdparty, so no
source, config (configs/nerf/base.json) or golden output exists here.  What follows restates the
PUBLISHED instant-ngp algorithm (Mueller et al. 2022; tiny-cuda-nn GridEncoding / SphericalHarmonics
/ FullyFusedMLP; instant-ngp testbed_nerf.cu) in plain PyTorch fp32 with autograd; the CUDA product
path (csrc/ngp*.cu) is checked against THIS, plus convergence tests on synthetic scenes and the
call-site contract of fusion/nerf_fusion.py:57-101,285-289,296-300,388-424.

Conventions (upstream defaults for configs/nerf/base.json):
  hash grid: 16 levels x 2 features, T = 2^19 entries/level, base resolution 16,
             per-level scale = exp(ln(2048*aabb_scale/16)/15); dense indexing while res^3 <= T;
             hash(x,y,z) = x ^ y*2654435761 ^ z*805459861 (uint32) mod T; trilinear interpolation
             at pos*scale + 0.5, scale = 16*b^l - 1
  density MLP 32 -> 64 (ReLU) -> 16, sigma = exp(out[0]);  SH degree 4 on the direction;
  rgb MLP 32 (=16+16) -> 64 -> 64 (ReLU) -> 3 (sigmoid);
  compositing: alpha = 1 - exp(-sigma dt), T *= 1 - alpha, stop when T < 1e-4;
  loss: Huber(delta=0.1)/5 on RGB (+ background blend) + lambda_d * (d - d*)^2 / cov  (fork's
  depth supervision, fusion/nerf_fusion.py:99-101,285-289), averaged over rays.
"""
import math

import numpy as np
import torch

N_LEVELS, N_FEAT, LOG2_T, BASE_RES = 16, 2, 19, 16
PRIMES = (1, 2654435761, 805459861)


def per_level_scale(aabb_scale):
    return math.exp(math.log(2048.0 * aabb_scale / BASE_RES) / (N_LEVELS - 1))


def level_params(aabb_scale):
    """-> list of (scale, resolution, n_params, offset) ; n_params multiple of 8 capped at T"""
    b = per_level_scale(aabb_scale)
    T = 1 << LOG2_T
    out, off = [], 0
    for l in range(N_LEVELS):
        scale = BASE_RES * (b ** l) - 1.0
        res = int(math.ceil(scale)) + 1
        n = min(((res ** 3 + 7) // 8) * 8, T) if res ** 3 < (1 << 62) else T
        out.append((float(np.float32(scale)), res, n, off))
        off += n
    return out, off


def grid_index(ix, iy, iz, res, n_params):
    """ix,iy,iz int64 tensors. dense if res^3 <= n_params else hashed"""
    if res ** 3 <= n_params:
        return (ix + iy * res + iz * res * res) % n_params
    h = (ix * PRIMES[0]) ^ ((iy * PRIMES[1]) & 0xFFFFFFFF) ^ ((iz * PRIMES[2]) & 0xFFFFFFFF)
    return (h & 0xFFFFFFFF) % n_params


def hash_encode(x, grid, aabb_scale):
    """x [S,3] in [0,1]^3 (fp32), grid [total,2] fp32 -> [S,32]"""
    lv, _ = level_params(aabb_scale)
    feats = []
    for (scale, res, n, off) in lv:
        pos = x * scale + 0.5
        p0 = torch.floor(pos)
        w = pos - p0
        p0 = p0.long()
        acc = 0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    idx = grid_index(p0[:, 0] + dx, p0[:, 1] + dy, p0[:, 2] + dz, res, n)
                    wgt = (w[:, 0] if dx else 1 - w[:, 0]) * (w[:, 1] if dy else 1 - w[:, 1]) * \
                          (w[:, 2] if dz else 1 - w[:, 2])
                    acc = acc + wgt[:, None] * grid[off + idx]
        feats.append(acc)
    return torch.cat(feats, -1)


def sh4(d):
    """spherical harmonics degree 4 (16 coefficients), tiny-cuda-nn convention, d unit [S,3]"""
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([
        0.28209479177387814 * torch.ones_like(x),
        -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz, 0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2), 2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2), 0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2), 1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2)], -1)


def network(x01, d, P, aabb_scale):
    """P: dict(grid [total,2], W1 [32,64], W2 [64,16], W3 [32,64], W4 [64,64], W5 [64,16]) -> rgb [S,3], sigma [S]"""
    enc = hash_encode(x01, P["grid"], aabb_scale)
    h = torch.relu(enc @ P["W1"])
    o = h @ P["W2"]
    sigma = torch.exp(o[:, 0])
    inp = torch.cat([o, sh4(d)], -1)
    h = torch.relu(inp @ P["W3"])
    h = torch.relu(h @ P["W4"])
    rgb = torch.sigmoid((h @ P["W5"])[:, :3])
    return rgb, sigma


def composite_loss(rgb, sigma, dt, tdist, ray_ptr, target_rgb, target_depth, depth_cov, bg, lambda_d,
                   min_T=1e-4):
    """per-ray volume rendering + loss. ray_ptr [R+1] CSR over samples (host list).
    returns (loss, rgb_ray [R,3], depth_ray [R])"""
    R = len(ray_ptr) - 1
    tot = 0
    rgbs, deps = [], []
    for r in range(R):
        a, b = ray_ptr[r], ray_ptr[r + 1]
        T = torch.ones((), dtype=rgb.dtype)
        c = torch.zeros(3, dtype=rgb.dtype)
        dep = torch.zeros((), dtype=rgb.dtype)
        for s in range(a, b):
            if float(T) < min_T:
                break
            alpha = 1 - torch.exp(-sigma[s] * dt[s])
            w = alpha * T
            c = c + w * rgb[s]
            dep = dep + w * tdist[s]
            T = T * (1 - alpha)
        c = c + T * bg[r]
        diff = c - target_rgb[r]
        ad = diff.abs()
        hub = torch.where(ad < 0.1, 0.5 * diff * diff / 0.1, ad - 0.05) / 5.0
        l = hub.mean()
        if target_depth[r] > 0:
            l = l + lambda_d * (dep - target_depth[r]) ** 2 / depth_cov[r]
        tot = tot + l
        rgbs.append(c); deps.append(dep)
    return tot / R, torch.stack(rgbs), torch.stack(deps)


def init_params(aabb_scale, seed=0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    _, total = level_params(aabb_scale)
    P = {"grid": (torch.rand(total, 2, generator=g, dtype=dtype) * 2 - 1) * 1e-4}
    for name, (i, o) in dict(W1=(32, 64), W2=(64, 16), W3=(32, 64), W4=(64, 64), W5=(64, 16)).items():
        s = math.sqrt(6.0 / (i + o))       # xavier uniform (tcnn default for FullyFusedMLP)
        P[name] = (torch.rand(i, o, generator=g, dtype=dtype) * 2 - 1) * s
    return P


# --------------------------------------------------------------------------------------------------------------
# Ray march through the cascaded occupancy grid (the sampler of a training / render step).
# Restates nerf_slam_b200/csrc/ngp_train.cu::march_warp in its SERIAL meaning (fp32 arithmetic, numpy scalars):
#   step lattice t_{k+1} = t_k + dt(t_k), dt(t) = clamp(t * cone, MIN_STEP, MAX_STEP), independent of occupancy;
#   a sample is emitted at every lattice point that lies inside the box and in an occupied cell of its cascade
#   (cascade from the step size and the position, as in instant-ngp), up to max_n samples / 4096 lattice points.
# The published instant-ngp marcher skips ahead to the next voxel boundary when a cell is empty and then snaps to
# the same lattice; both visit the same lattice and differ only where the cascade changes inside a skipped voxel.
GRID = 128
MAX_STEPS = 1024
MIN_STEP = np.float32(np.float32(1.73205080757) / np.float32(MAX_STEPS))
MAX_STEP = np.float32(MIN_STEP * np.float32(128.0) * np.float32(8.0))
MAX_LATTICE = 4096


def calc_dt(t, cone):
    return np.float32(min(max(np.float32(t) * np.float32(cone), MIN_STEP), MAX_STEP))


def _mip_from_pos(p, cascades):
    m = max(abs(p[0] - np.float32(0.5)), abs(p[1] - np.float32(0.5)), abs(p[2] - np.float32(0.5)))
    e = math.frexp(float(m))[1] if m != 0 else 0
    return min(max(e + 1, 0), cascades - 1)


def mip_from_dt(dt, p, cascades):
    mip = _mip_from_pos(p, cascades)
    d = np.float32(dt) * np.float32(2 * GRID)
    if d < 1.0:
        return mip
    e = math.frexp(float(d))[1]
    return min(max(max(e, mip), 0), cascades - 1)


def occupied(p, mip, bits):
    """bits: uint8 bitfield [cascades * GRID^3 / 8]"""
    s = np.float32(2.0 ** (-mip))
    idx = [int(math.floor(float(((np.float32(p[a]) - np.float32(0.5)) * s + np.float32(0.5)) * np.float32(GRID)))) for a in range(3)]
    if min(idx) < 0 or max(idx) >= GRID:
        return False
    i = idx[0] + GRID * (idx[1] + GRID * idx[2]) + mip * GRID ** 3
    return bool((bits[i >> 3] >> (i & 7)) & 1)


def march_lattice(o, d, aabb_lo, aabb_hi, near, cone, cascades, bits, jitter, max_n):
    """-> list of (t, dt) of the emitted samples; o, d float32 [3] (d unit length)"""
    o = np.asarray(o, np.float32); d = np.asarray(d, np.float32)
    with np.errstate(divide="ignore"):
        inv = np.float32(1.0) / d
    tmin, tmax = np.float32(-1e30), np.float32(1e30)
    for a in range(3):
        t0, t1 = (np.float32(aabb_lo) - o[a]) * inv[a], (np.float32(aabb_hi) - o[a]) * inv[a]
        tmin = max(tmin, min(t0, t1)); tmax = min(tmax, max(t0, t1))
    if tmax <= max(tmin, np.float32(0.0)):
        return []
    t = np.float32(max(tmin, np.float32(near)) + np.float32(1e-6))
    t = np.float32(t + calc_dt(t, cone) * np.float32(jitter))
    out = []
    for _ in range(MAX_LATTICE):
        if not (t < tmax) or len(out) >= max_n:
            break
        dt = calc_dt(t, cone)
        p = [np.float32(np.float32(d[a]) * t + o[a]) for a in range(3)]     # fmaf(d, t, o) up to one rounding
        if occupied(p, mip_from_dt(dt, p, cascades), bits):
            out.append((t, dt))
        t = np.float32(t + dt)
    return out


def pcg(v):
    """counter-based hash of nerf_slam_b200/csrc/ngp_common.cuh::pcg (uint32 arithmetic)"""
    v &= 0xFFFFFFFF
    s_ = (v * 747796405 + 2891336453) & 0xFFFFFFFF
    w = (((s_ >> ((s_ >> 28) + 4)) ^ s_) * 277803737) & 0xFFFFFFFF
    return ((w >> 22) ^ w) & 0xFFFFFFFF


def rnd01(a, b, c):
    """uniform [0,1) of ngp_common.cuh::rnd01(a, b, c)"""
    return np.float32(np.float32(pcg(pcg(pcg(a) ^ (b & 0xFFFFFFFF)) ^ (c & 0xFFFFFFFF)) >> 8) * np.float32(1.0 / 16777216.0))


def render_view(P, c2w34, intr, width, height, aabb_scale, cascades, bits, near, bg, max_per_ray=1024, cone=1.0 / 256,
                min_T=1e-4):
    """B4 — Testbed.render of a pinhole view (fusion/nerf_fusion.py:411-424): per pixel centre the ray of the camera
    record, the occupancy-lattice march without jitter, the network, front-to-back compositing over `bg`.
    -> rgb [H,W,3] float32, z-depth [H,W] (sum_k w_k t_k * inv_len, the depth convention of the training loss)."""
    fx, fy, cx, cy = [np.float32(v) for v in intr]
    c2w = np.asarray(c2w34, np.float32).reshape(3, 4)
    lo, hi = np.float32(0.5 - 0.5 * aabb_scale), np.float32(0.5 + 0.5 * aabb_scale)
    rgb_out = np.zeros((height, width, 3), np.float32); dep_out = np.zeros((height, width), np.float32)
    for py in range(height):
        for px in range(width):
            dx, dy = (np.float32(px + 0.5) - cx) / fx, (np.float32(py + 0.5) - cy) / fy
            inv_len = np.float32(1.0) / np.sqrt(dx * dx + dy * dy + np.float32(1.0))
            dc = np.array([dx * inv_len, dy * inv_len, inv_len], np.float32)
            d = (c2w[:, :3] @ dc).astype(np.float32); o = c2w[:, 3].copy()
            samples = march_lattice(o, d, lo, hi, near, cone, cascades, bits, 0.0, max_per_ray)
            c = np.array(bg, np.float32).copy()
            if samples:
                t = torch.tensor([float(s[0]) for s in samples]); dt = torch.tensor([float(s[1]) for s in samples])
                x01 = (torch.from_numpy(o)[None] + t[:, None] * torch.from_numpy(d)[None] - float(lo)) / float(hi - lo)
                with torch.no_grad():
                    rgb, sigma = network(x01, torch.from_numpy(d)[None].repeat(len(samples), 1), P, aabb_scale)
                T, c, dep = 1.0, np.zeros(3, np.float32), 0.0
                for k in range(len(samples)):
                    if T < min_T:
                        break
                    a = 1.0 - math.exp(-float(sigma[k]) * float(dt[k]))
                    w = a * T
                    c = c + w * rgb[k].numpy(); dep += w * float(t[k]); T *= 1.0 - a
                c = c + T * np.array(bg, np.float32)
                dep_out[py, px] = dep * float(inv_len)
            rgb_out[py, px] = c
    return rgb_out, dep_out


# ------------------------------------------------------------------------------------------ B1 / B2
def srgb_to_linear(x):
    """utils/utils.py:136-139 (fp32 torch)"""
    return torch.where(x > 0.04045, torch.pow((x + 0.055) / 1.055, 2.4), x / 12.92)


def pose_tq_to_matrix(tq):
    """(t, q_xyzw) -> 4x4, fp64 (lietorch SE3.matrix(); formulas as in src/droid_kernels.cu:66-120)"""
    tq = np.asarray(tq, np.float64)
    out = np.zeros((len(tq), 4, 4))
    for k, v in enumerate(tq):
        x, y, z, w = v[3:]
        out[k, :3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        out[k, :3, 3] = v[:3]
        out[k, 3, 3] = 1.0
    return out


def process_slam_tuples(slam, mask_type="ours"):
    """What NerfFusion.process_slam + send_data hand to the trainer (fusion/nerf_fusion.py:140-289) for one SLAM packet:
    -> dict(ids, poses [n,3,4] world_T_cam (scale 1, offset 0, :167-170), images [n,H,W,4] fp32 linear colour
    premultiplied by alpha = 1 (:203-215), depths [n,H,W,1] = 1/idepth_up, depths_cov [n,H,W,1], scales (1, 1));
    mask types (:173-183): raw -> unit covariance; ours_w_thresh -> idepth = -1 where sqrt(cov) > median(cov) [sic: the
    quantile is taken of the covariance, not of its square root]; no_depth -> idepth = -1 everywhere."""
    idepths = slam["cam0_idepths_up"].clone().float()
    cov = slam["cam0_depths_cov_up"].clone().float()
    if mask_type == "raw":
        cov[...] = 1.0
    elif mask_type == "ours_w_thresh":
        idepths[cov.sqrt() > cov.quantile(0.50)] = -1.0
    elif mask_type == "no_depth":
        idepths[...] = -1.0
    elif mask_type != "ours":
        raise NotImplementedError(mask_type)
    img = slam["cam0_images"].permute(0, 2, 3, 1).float() / 255.0
    rgb = srgb_to_linear(img)                      # alpha = 255/255 = 1: premultiplication leaves the colour unchanged
    images = torch.cat([rgb, torch.ones_like(rgb[..., :1])], -1)
    w2c = pose_tq_to_matrix(slam["cam0_poses"].cpu().numpy())
    c2w = np.linalg.inv(w2c.astype(np.float32))    # the reference inverts the fp32 matrices (numpy)
    return {"ids": [int(v) for v in slam["viz_idx"].tolist()], "poses": c2w[:, :3, :4], "images": images.numpy(),
            "depths": (1.0 / idepths[..., None]).numpy(), "depths_cov": cov[..., None].numpy(), "scales": (1.0, 1.0)}
