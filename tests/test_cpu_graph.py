"""edge selection (A18): the vectorised host logic must pick bit-identical edges, in the same order,
as the loop-for-loop restatement of the reference (oracle/graph.py)"""
import numpy as np
import pytest

from nerf_slam_b200.graph import proximity_edges, proximity_edges_numpy
from oracle.graph import add_proximity_factors_edges


@pytest.mark.parametrize("seed", range(40))
def test_proximity_edges_bit_exact(seed):
    rng = np.random.default_rng(seed)
    t = int(rng.integers(3, 26))
    kf0 = int(rng.integers(0, t)); kf1 = int(rng.integers(0, t))
    if seed % 3 == 0:                      # the frontend's steady-state call: kf0 = kf_idx - 4, kf1 = max(kf_idx + 1 - window, 0)
        kf0 = max(t - 5, 0); kf1 = max(t - 25, 0)
    rad = int(rng.integers(1, 4)); nms = int(rng.integers(0, 3))
    stereo = bool(rng.integers(0, 2)); max_factors = int(rng.integers(4, 80))
    thresh = float(rng.uniform(5, 40))
    n = (t - kf0) * (t - kf1)
    d = rng.uniform(0, 60, n).astype(np.float32)
    d[rng.random(n) < 0.05] = 150.0                                     # > 100 -> inf
    d[rng.random(n) < 0.1] = np.float32(7.5)                            # ties
    ne = int(rng.integers(0, 30))
    ii1 = rng.integers(0, t, ne); jj1 = rng.integers(0, t, ne)
    ix, jx = np.meshgrid(np.arange(kf0, t), np.arange(kf1, t), indexing="ij")
    ref = add_proximity_factors_edges(d, kf0, kf1, t, ii1, jj1, rad, nms, thresh, max_factors, stereo)
    for fn in (proximity_edges, proximity_edges_numpy):           # native host routine (the product), vectorised numpy twin
        got = fn(d.copy(), ix.reshape(-1), jx.reshape(-1), ii1, jj1, kf0, kf1, t, rad, nms, thresh, max_factors, stereo)
        assert got.shape == ref.shape and np.array_equal(got, ref), fn.__name__


def test_proximity_edges_empty_and_single():
    d = np.zeros(1, np.float32)
    got = proximity_edges(d.copy(), np.array([0]), np.array([0]), np.zeros(0, np.int64), np.zeros(0, np.int64), 0, 0, 1, 2, 2, 16.0, 48, False)
    ref = add_proximity_factors_edges(d, 0, 0, 1, [], [], 2, 2, 16.0, 48, False)
    assert np.array_equal(got, ref)


def test_native_and_vectorised_ba_graph_tables_equal_the_loop_formulation():
    """nerf_slam_b200.ba_graph.BAGraphHost — built by the native host routine nslam_ba_graph_build (csrc/ba_graph_host.cu,
    0.04 ms per keyframe candidate while the GPU waits) — and its vectorised numpy twin `from_numpy` must produce
    byte-identical int32 tables (and the same packed upload buffer) as the loop formulation on random windows: edges
    outside the window, fixed frames, duplicate and self (stereo) edges, empty edge lists, single-pose windows."""
    from nerf_slam_b200.ba_graph import BAGraphHost
    from tests.ba_graph_loops import BAGraphLoops
    rng = np.random.default_rng(0)
    for trial in range(600):
        P = int(rng.integers(1, 14)); kf0 = int(rng.integers(0, 6)); kf1 = kf0 + P
        E = int(rng.integers(0, 90))
        lo = max(0, kf0 - int(rng.integers(0, 5))); hi = kf1 + int(rng.integers(0, 3))
        ii = rng.integers(lo, hi, E); jj = rng.integers(lo, hi, E)
        if trial % 7 == 0:
            jj = ii.copy()
        a = BAGraphLoops(ii, jj, kf0, kf1)
        for b in (BAGraphHost(ii, jj, kf0, kf1), BAGraphHost.from_numpy(ii, jj, kf0, kf1)):
            for name in ("E", "P", "K", "kf0", "kf1", "NR", "NPAIR", "RMAX", "NHC", "NVC"):
                assert getattr(a, name) == getattr(b, name), (trial, name)
            assert list(a.tables) == list(b.tables)
            for k in a.tables:
                assert b.tables[k].dtype == np.int32 and np.array_equal(a.tables[k], b.tables[k]), (trial, k)
        fa, oa = BAGraphHost(ii, jj, kf0, kf1).packed()
        fb, ob = BAGraphHost.from_numpy(ii, jj, kf0, kf1).packed()
        assert oa == ob and np.array_equal(fa, fb) and all(o % 4 == 0 for o in oa.values()), trial


def test_native_proximity_selection_equals_numpy_on_many_windows():
    """nslam_proximity_edges (csrc/ba_graph_host.cu) vs proximity_edges_numpy: ties, > 100 cut-off, all nms radii, stereo,
    the max_factors break, negative (wrapping) flat indices; windows that the reference's own indexing rejects are skipped"""
    ok = 0
    for seed in range(1200):
        rng = np.random.default_rng(10_000 + seed)
        t = int(rng.integers(1, 30)); kf0 = int(rng.integers(0, t)); kf1 = int(rng.integers(0, t))
        if seed % 3 == 0:
            kf0 = max(t - 5, 0); kf1 = max(t - 25, 0)
        if seed % 5 == 0:
            kf0 = kf1 = 0
        rad = int(rng.integers(1, 4)); nms = int(rng.integers(0, 4)); stereo = bool(rng.integers(0, 2))
        mf = int(rng.integers(4, 120)); thresh = float(rng.uniform(5, 40))
        n = (t - kf0) * (t - kf1)
        d = rng.uniform(0, 60, n).astype(np.float32); d[rng.random(n) < 0.05] = 150.0; d[rng.random(n) < 0.1] = np.float32(7.5)
        ne = int(rng.integers(0, 40)); ii1 = rng.integers(0, t, ne); jj1 = rng.integers(0, t, ne)
        ix, jx = np.meshgrid(np.arange(kf0, t), np.arange(kf1, t), indexing="ij")
        try:
            a = proximity_edges_numpy(d.copy(), ix.reshape(-1), jx.reshape(-1), ii1, jj1, kf0, kf1, t, rad, nms, thresh, mf, stereo)
        except IndexError:
            continue
        b = proximity_edges(d.copy(), ix.reshape(-1), jx.reshape(-1), ii1, jj1, kf0, kf1, t, rad, nms, thresh, mf, stereo)
        assert a.shape == b.shape and np.array_equal(a, b), seed
        ok += 1
    assert ok > 1000
