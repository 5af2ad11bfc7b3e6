"""Loop formulation of the BA graph tables (the first implementation of nerf_slam_b200/ba_graph.py, one Python loop per depth
map): kept as the reference the vectorised builder must match table for table (tests/test_cpu_graph.py)."""
from collections import OrderedDict

import numpy as np


class BAGraphLoops:
    """numpy tables; see include/nslam_ba.h::nslam_ba_graph for the meaning of each field."""

    def __init__(self, ii, jj, kf0, kf1):
        ii = np.asarray(ii, dtype=np.int64).reshape(-1)
        jj = np.asarray(jj, dtype=np.int64).reshape(-1)
        E = int(ii.shape[0])
        P = int(kf1 - kf0)
        if P <= 0:
            raise ValueError("empty BA window")
        ts = np.arange(kf0, kf1, dtype=np.int64)
        kx = np.unique(np.concatenate([ts, ii]))
        K = int(kx.shape[0])
        kk = np.searchsorted(kx, ii)

        order = np.argsort(kk, kind="stable")
        src_ptr = np.zeros(K + 1, dtype=np.int64)
        np.add.at(src_ptr, kk + 1, 1)
        src_ptr = np.cumsum(src_ptr)

        # Schur rows per depth map: self row (if the frame is in the window) + edges whose target
        # pose is in the window (schur_block keeps `j >= kf0 && j <= kf1`, :1368)
        row_pose, row_erow, row_k = [], [], []
        in_win = (kx >= kf0) & (kx < kf1)
        jw = (jj >= kf0) & (jj < kf1)
        for k in range(K):
            if in_win[k]:
                row_pose.append(int(kx[k] - kf0)); row_erow.append(int(kx[k] - kf0)); row_k.append(k)
            for e in order[src_ptr[k]:src_ptr[k + 1]]:
                if jw[e]:
                    row_pose.append(int(jj[e] - kf0)); row_erow.append(P + int(e)); row_k.append(k)
        row_pose = np.asarray(row_pose, dtype=np.int64)
        row_erow = np.asarray(row_erow, dtype=np.int64)
        row_k = np.asarray(row_k, dtype=np.int64)
        NR = int(row_pose.shape[0])
        row_ptr = np.zeros(K + 1, dtype=np.int64)
        np.add.at(row_ptr, row_k + 1, 1)
        row_ptr = np.cumsum(row_ptr)
        R = np.diff(row_ptr)
        pair_off = np.concatenate([[0], np.cumsum(R * R)])
        NPAIR = int(pair_off[-1])
        RMAX = int(R.max()) if K else 0

        # dense assembly: contributions to block (a, b) of H and to segment a of v
        a = ii - kf0
        b = jj - kf0
        av = (a >= 0) & (a < P)
        bv = (b >= 0) & (b < P)
        e_idx = np.arange(E, dtype=np.int64)
        keys = [a * P + a, a * P + b, b * P + a, b * P + b]
        oks = [av, av & bv, av & bv, bv]
        hk = [keys[w][oks[w]] for w in range(4)]
        hv = [(w * E + e_idx)[oks[w]] for w in range(4)]
        # Schur blocks
        sk, sv = [], []
        for k in range(K):
            r0, r1 = int(row_ptr[k]), int(row_ptr[k + 1])
            if r1 == r0:
                continue
            pp = row_pose[r0:r1]
            Rk = r1 - r0
            blk = pair_off[k] + np.arange(Rk * Rk)
            sk.append((pp[:, None] * P + pp[None, :]).reshape(-1))
            sv.append(-(blk + 1))
        hkeys = np.concatenate(hk + sk) if (E or sk) else np.zeros(0, np.int64)
        hvals = np.concatenate(hv + sv) if (E or sv) else np.zeros(0, np.int64)
        o = np.argsort(hkeys, kind="stable")
        hc_idx = hvals[o]
        hc_ptr = np.zeros(P * P + 1, dtype=np.int64)
        np.add.at(hc_ptr, hkeys + 1, 1)
        hc_ptr = np.cumsum(hc_ptr)

        vkeys = np.concatenate([a[av], b[bv], row_pose])
        vvals = np.concatenate([e_idx[av], (E + e_idx)[bv], -(np.arange(NR) + 1)])
        o = np.argsort(vkeys, kind="stable")
        vc_idx = vvals[o]
        vc_ptr = np.zeros(P + 1, dtype=np.int64)
        np.add.at(vc_ptr, vkeys + 1, 1)
        vc_ptr = np.cumsum(vc_ptr)

        self.E, self.P, self.K, self.kf0, self.kf1 = E, P, K, int(kf0), int(kf1)
        self.NR, self.NPAIR, self.RMAX = NR, NPAIR, RMAX
        self.NHC, self.NVC = int(hc_idx.shape[0]), int(vc_idx.shape[0])
        self.kk = kk
        self.tables = OrderedDict(
            ii=ii, jj=jj, kx=kx, src_ptr=src_ptr, src_edges=order, row_ptr=row_ptr,
            row_pose=row_pose, row_erow=row_erow, pair_off=pair_off, hc_ptr=hc_ptr,
            hc_idx=hc_idx, vc_ptr=vc_ptr, vc_idx=vc_idx)
        for k, v in self.tables.items():
            self.tables[k] = np.ascontiguousarray(v, dtype=np.int32)

