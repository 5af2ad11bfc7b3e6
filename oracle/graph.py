"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): loop-for-loop restatement of the reference's
proximity edge selection, slam/visual_frontends/visual_frontend.py:711-775, on numpy arrays.
`d` plays the role of the torch tensor returned by self.distance(ii, jj) (:721)."""
import numpy as np


def add_proximity_factors_edges(d, kf0, kf1, t, ii1, jj1, rad=2, nms=2, thresh=16.0, max_factors=48, stereo=False):
    d = np.array(d, dtype=np.float32, copy=True)
    ix = np.arange(kf0, t); jx = np.arange(kf1, t)
    ii, jj = np.meshgrid(ix, jx, indexing="ij")                       # torch.meshgrid default = 'ij' (:716)
    ii = ii.reshape(-1); jj = jj.reshape(-1)
    d[(ii - rad) < jj] = np.inf                                         # :722
    d[d > 100] = np.inf                                                 # :723
    for i, j in zip(ii1, jj1):                                          # :727-736
        for di in range(-nms, nms + 1):
            for dj in range(-nms, nms + 1):
                if abs(di) + abs(dj) <= max(min(abs(i - j) - 2, nms), 0):
                    i1 = i + di
                    j1 = j + dj
                    if (kf0 <= i1 < t) and (kf1 <= j1 < t):
                        d[(i1 - kf0) * (t - kf1) + (j1 - kf1)] = np.inf
    es = []
    for i in range(kf0, t):                                             # :738-747
        if stereo:
            es.append((i, i))
            d[(i - kf0) * (t - kf1) + (i - kf1)] = np.inf
        for j in range(max(i - rad - 1, 0), i):
            es.append((i, j))
            es.append((j, i))
            d[(i - kf0) * (t - kf1) + (j - kf1)] = np.inf
    order = np.argsort(d, kind="stable")                                # :749
    for k in order:                                                     # :750-771
        if d[k] > thresh:
            continue
        if len(es) > max_factors:
            break
        i = int(ii[k]); j = int(jj[k])
        es.append((i, j))
        es.append((j, i))
        for di in range(-nms, nms + 1):
            for dj in range(-nms, nms + 1):
                if abs(di) + abs(dj) <= max(min(abs(i - j) - 2, nms), 0):
                    i1 = i + di
                    j1 = j + dj
                    if (kf0 <= i1 < t) and (kf1 <= j1 < t):
                        d[(i1 - kf0) * (t - kf1) + (j1 - kf1)] = np.inf
    return np.asarray(es, dtype=np.int64).reshape(-1, 2)
