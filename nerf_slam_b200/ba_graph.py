"""Host-side structure of one bundle-adjustment window -> int32 tables for the BA kernels.

This is the part the reference does on the CPU inside `reduced_camera_matrix_cuda`
(`torch::_unique`, src/droid_kernels.cu:1697-1706), `accum_cuda` (argsort + CSR, :1065-1115) and
`schur_block` (O(P^2 deg^2) enumeration of co-visible (i,j,k) triples, :1349-1399) on EVERY call,
with device->host copies of ii/jj.  Here it is computed once per distinct graph from the host
copy of the edge list (the front-end owns the edges on the host anyway), vectorised in numpy,
cached, and uploaded as a single int32 buffer.
"""
import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib


_TABLES = ("ii", "jj", "kx", "src_ptr", "src_edges", "row_ptr", "row_pose", "row_erow", "pair_off", "hc_ptr", "hc_idx",
           "vc_ptr", "vc_idx")


class BAGraphHost:
    """Tables of one BA window; see include/nslam_ba.h::nslam_ba_graph for the meaning of each field.

    Built by the native host routine `nslam_ba_graph_build` (csrc/ba_graph_host.cu) directly into the packed int32 buffer
    that is uploaded — the reference does this part in C++ on the CPU too (src/droid_kernels.cu:1065-1115, 1349-1399,
    1697-1706), on every call.  `BAGraphHost.from_numpy` is the same construction in vectorised numpy, kept as the
    table-for-table check of the native routine (tests/test_cpu_graph.py)."""

    def __init__(self, ii, jj, kf0, kf1):
        ii = np.ascontiguousarray(np.asarray(ii, dtype=np.int64).reshape(-1))
        jj = np.ascontiguousarray(np.asarray(jj, dtype=np.int64).reshape(-1))
        E = int(ii.shape[0])
        if int(kf1 - kf0) <= 0:
            raise ValueError("empty BA window")
        lib = _lib.load(require_cuda=False)
        meta = np.zeros(36, np.int32)
        cap = 4096 + 64 * E
        for _ in range(2):
            flat = np.empty(cap, np.int32)
            rc = lib.nslam_ba_graph_build(ii.ctypes.data_as(ctypes.c_void_p), jj.ctypes.data_as(ctypes.c_void_p), E, int(kf0), int(kf1),
                                          flat.ctypes.data_as(ctypes.c_void_p), cap, meta.ctypes.data_as(ctypes.c_void_p))
            if rc != 1:
                break
            cap = int(meta[9])
        if rc != 0:
            raise RuntimeError(f"nslam_ba_graph_build failed ({rc})")
        self.E, self.P, self.K, self.kf0 = (int(v) for v in meta[:4])
        self.kf1 = int(kf1)
        self.NR, self.NPAIR, self.RMAX, self.NHC, self.NVC = (int(v) for v in meta[4:9])
        self._flat = flat[:int(meta[9])]
        self._offs = {name: int(meta[10 + t]) for t, name in enumerate(_TABLES)}
        self.tables = OrderedDict((name, self._flat[self._offs[name]:self._offs[name] + int(meta[23 + t])])
                                  for t, name in enumerate(_TABLES))

    @classmethod
    def from_numpy(cls, ii, jj, kf0, kf1):
        self = cls.__new__(cls)
        self._flat = None
        self._init_numpy(ii, jj, kf0, kf1)
        return self

    def _init_numpy(self, ii, jj, kf0, kf1):
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
        src_ptr = np.concatenate([[0], np.cumsum(np.bincount(kk, minlength=K))]).astype(np.int64)

        # Schur rows per depth map: self row (if the frame is in the window) + edges whose target
        # pose is in the window (schur_block keeps `j >= kf0 && j <= kf1`, :1368).
        # Order inside a depth map: self row first, then its edges in `order` (stable by edge id); built without Python
        # loops: self rows and (already k-sorted) edge rows are concatenated and stably sorted by k.
        in_win = (kx >= kf0) & (kx < kf1)
        k_self = np.nonzero(in_win)[0]
        e_rows = order[(jj[order] >= kf0) & (jj[order] < kf1)]
        cand_k = np.concatenate([k_self, kk[e_rows]])
        cand_pose = np.concatenate([kx[k_self] - kf0, jj[e_rows] - kf0])
        cand_erow = np.concatenate([kx[k_self] - kf0, P + e_rows])
        o = np.argsort(cand_k, kind="stable")
        row_k, row_pose, row_erow = cand_k[o], cand_pose[o].astype(np.int64), cand_erow[o].astype(np.int64)
        NR = int(row_pose.shape[0])
        R = np.bincount(row_k, minlength=K).astype(np.int64)
        row_ptr = np.concatenate([[0], np.cumsum(R)]).astype(np.int64)
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
        # Schur blocks: for every depth map all ordered pairs (ra, rb) of its rows, ra-major; block id = running index
        rep = R[row_k]                                              # pairs that start at each row
        ia = np.repeat(np.arange(NR, dtype=np.int64), rep)
        ib = row_ptr[row_k[ia]] + (np.arange(NPAIR, dtype=np.int64) - np.repeat(np.cumsum(rep) - rep, rep))
        sk = [row_pose[ia] * P + row_pose[ib]] if NPAIR else []
        sv = [-(np.arange(NPAIR, dtype=np.int64) + 1)] if NPAIR else []
        hkeys = np.concatenate(hk + sk) if (E or sk) else np.zeros(0, np.int64)
        hvals = np.concatenate(hv + sv) if (E or sv) else np.zeros(0, np.int64)
        o = np.argsort(hkeys, kind="stable")
        hc_idx = hvals[o]
        hc_ptr = np.concatenate([[0], np.cumsum(np.bincount(hkeys, minlength=P * P))]).astype(np.int64)

        vkeys = np.concatenate([a[av], b[bv], row_pose])
        vvals = np.concatenate([e_idx[av], (E + e_idx)[bv], -(np.arange(NR) + 1)])
        o = np.argsort(vkeys, kind="stable")
        vc_idx = vvals[o]
        vc_ptr = np.concatenate([[0], np.cumsum(np.bincount(vkeys, minlength=P))]).astype(np.int64)

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

    # ------------------------------------------------------------------ device side
    def packed(self):
        """one int32 vector + offsets (each table 16-byte aligned)"""
        if self._flat is not None:
            return self._flat, self._offs
        offs, chunks, pos = {}, [], 0
        for k, v in self.tables.items():
            offs[k] = pos
            n = v.shape[0]
            pad = (-n) % 4
            chunks.append(v)
            if pad:
                chunks.append(np.zeros(pad, dtype=np.int32))
            pos += n + pad
        if pos == 0:
            chunks.append(np.zeros(4, dtype=np.int32))
        return np.concatenate(chunks), offs

    def to_device(self, device):
        import torch
        flat, offs = self.packed()
        buf = torch.from_numpy(flat).pin_memory().to(device, non_blocking=True)
        g = _lib.BAGraph()
        for name in ("E", "P", "K", "kf0", "NR", "NPAIR", "RMAX", "NHC", "NVC"):
            setattr(g, name, getattr(self, name))
        base = buf.data_ptr()
        for name, off in offs.items():
            setattr(g, name, ctypes.c_void_p(base + 4 * off))
        return g, buf


_CACHE = OrderedDict()


def get_graph(ii_host, jj_host, kf0, kf1, device):
    """cached (BAGraphHost, BAGraph struct, device buffer) for an edge list"""
    ii_host = np.ascontiguousarray(ii_host, dtype=np.int64)
    jj_host = np.ascontiguousarray(jj_host, dtype=np.int64)
    key = (ii_host.tobytes(), jj_host.tobytes(), int(kf0), int(kf1), str(device))
    hit = _CACHE.get(key)
    if hit is not None:
        _CACHE.move_to_end(key)
        return hit
    gh = BAGraphHost(ii_host, jj_host, kf0, kf1)
    g, buf = gh.to_device(device)
    _CACHE[key] = (gh, g, buf)
    while len(_CACHE) > 16:
        _CACHE.popitem(last=False)
    return gh, g, buf
