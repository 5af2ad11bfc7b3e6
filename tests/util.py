"""shared synthetic-problem generators for the tests (seeded, SURVEY.md §8d micro-benchmarks)"""
import numpy as np

from oracle import se3


def make_window(rng, nframes=6, ht=12, wd=16, extra_edges=4, fixed_front=0):
    """sliding-window BA problem: radius-2 neighbourhood + random proximity edges"""
    poses = se3.random_poses(rng, nframes, 0.1, 5.0, np.float32)
    disps = rng.uniform(0.2, 2.0, (nframes, ht, wd)).astype(np.float32)
    intr = np.array([wd * 0.5, wd * 0.5, wd / 2 - 0.5, ht / 2 - 0.5], np.float32)
    es = [(i, j) for i in range(nframes) for j in range(nframes) if i != j and abs(i - j) <= 2]
    cand = [(i, j) for i in range(nframes) for j in range(nframes) if abs(i - j) > 2]
    if cand and extra_edges:
        for k in rng.choice(len(cand), min(extra_edges, len(cand)), replace=False):
            es.append(cand[k])
    ii = np.array([e[0] for e in es], np.int64)
    jj = np.array([e[1] for e in es], np.int64)
    return poses, disps, intr, ii, jj


def make_targets(rng, poses, disps, intr, ii, jj, noise=0.5):
    from oracle import geom
    coords, _ = geom.reproject(poses, disps, intr, ii, jj)
    target = coords + rng.normal(0, noise, coords.shape)
    target = np.ascontiguousarray(target.transpose(0, 3, 1, 2)).astype(np.float32)  # [E,2,ht,wd]
    weight = rng.uniform(0, 1, target.shape).astype(np.float32)
    return target, weight
