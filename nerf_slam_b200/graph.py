"""Host side of the factor-graph management that is pure index logic:
the order-sensitive proximity edge selection of `add_proximity_factors`
(reference slam/visual_frontends/visual_frontend.py:711-775).  The distances come from the device
(`frame_distance` kernel); everything here runs on the host copy, like the reference, but with the
non-maximum suppression vectorised over edges instead of three nested Python loops per edge."""
import numpy as np


def _diamonds(nms):
    return [np.array([(di, dj) for di in range(-r, r + 1) for dj in range(-r, r + 1) if abs(di) + abs(dj) <= r],
                     dtype=np.int64).reshape(-1, 2) for r in range(nms + 1)]


def proximity_edges(d, ii, jj, ii1, jj1, kf0, kf1, t, rad, nms, thresh, max_factors, stereo):
    """The selection through the native host routine nslam_proximity_edges (csrc/ba_graph_host.cu); same arguments and
    results as proximity_edges_numpy below (the check in tests/test_cpu_graph.py)."""
    import ctypes
    from . import _lib
    lib = _lib.load(require_cuda=False)
    d = np.ascontiguousarray(d, dtype=np.float32)
    i1 = np.ascontiguousarray(ii1, dtype=np.int64).reshape(-1); j1 = np.ascontiguousarray(jj1, dtype=np.int64).reshape(-1)
    n_out = ctypes.c_int(0)
    # forced edges: <= 2 (rad + 1) + 1 per frame; the selection loop stops once the count exceeds max_factors
    cap = max(int(t) - int(kf0), 0) * (2 * int(rad) + 3) + max(int(max_factors), 0) + 8
    es = np.empty((cap, 2), np.int64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = lib.nslam_proximity_edges(P(d), int(kf0), int(kf1), int(t), P(i1), P(j1), int(i1.shape[0]), int(rad), int(nms),
                                   float(thresh), int(max_factors), int(bool(stereo)), P(es), cap, ctypes.byref(n_out))
    if rc != 0:
        raise RuntimeError(f"nslam_proximity_edges failed ({rc}; {n_out.value} edges, capacity {cap})")
    return es[:n_out.value].copy()


def proximity_edges_numpy(d, ii, jj, ii1, jj1, kf0, kf1, t, rad, nms, thresh, max_factors, stereo):
    """d: fp32 distances over the meshgrid (ii, jj) = [kf0,t) x [kf1,t) (row-major), MODIFIED in place;
    ii1/jj1: existing (active + bad + inactive) edges whose neighbourhoods are suppressed first.
    Returns the selected directed edges [n,2] in the reference's order."""
    W = t - kf1
    d[(ii - rad) < jj] = np.inf
    d[d > 100] = np.inf
    dia = _diamonds(nms)

    def suppress_many(ia, ja):
        ia = np.asarray(ia, np.int64); ja = np.asarray(ja, np.int64)
        if ia.size == 0:
            return
        rr = np.clip(np.abs(ia - ja) - 2, 0, nms)
        for r in range(nms + 1):
            sel = rr == r
            if not sel.any():
                continue
            i1 = (ia[sel][:, None] + dia[r][None, :, 0]).reshape(-1)
            j1 = (ja[sel][:, None] + dia[r][None, :, 1]).reshape(-1)
            ok = (i1 >= kf0) & (i1 < t) & (j1 >= kf1) & (j1 < t)
            d[(i1[ok] - kf0) * W + (j1[ok] - kf1)] = np.inf

    suppress_many(ii1, jj1)
    es = []
    for i in range(kf0, t):
        if stereo:
            es.append((i, i))
            d[(i - kf0) * W + (i - kf1)] = np.inf
        for j in range(max(i - rad - 1, 0), i):
            es.append((i, j)); es.append((j, i))
            d[(i - kf0) * W + (j - kf1)] = np.inf          # may be a negative (wrapping) index, as in the reference
    # torch.argsort (unstable) on the reference side; ties are resolved by index order here (stable),
    # which is what torch's CPU sort yields for equal fp32 keys
    for k in np.argsort(d, kind="stable"):
        if d[k] > thresh:
            continue
        if len(es) > max_factors:
            break
        i, j = int(ii[k]), int(jj[k])
        es.append((i, j)); es.append((j, i))
        suppress_many([i], [j])
    return np.asarray(es, dtype=np.int64).reshape(-1, 2)
