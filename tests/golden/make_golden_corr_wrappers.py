"""Record the outputs of the REFERENCE's Python correlation wrappers — `CorrBlock.__call__` (per-level coordinate scaling,
channel concatenation; networks/modules/corr.py:40-50) and `AltCorrBlock` (feature pyramid by average pooling, frame
indexing, per-level sampling, output layout; :92-140) — with the two CUDA kernels they call replaced by the CPU oracle's
restatements (oracle/corr.py: corr_index_forward, altcorr_forward; these are pinned against the reference's compiled
kernels on the GPU, tests/test_gpu_vs_reference.py).  Build container only.

  python tests/golden/make_golden_corr_wrappers.py        ->  tests/golden/ref_corr_wrappers.npz"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("NSLAM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def inputs(seed=17):
    rng = np.random.default_rng(seed)
    n, C, h, w = 5, 16, 16, 24
    fmaps = rng.normal(0, 1, (1, n, C, h, w)).astype(np.float32)
    ii = np.array([0, 1, 2, 3, 4, 2]); jj = np.array([1, 0, 4, 2, 3, 2])
    y, x = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    coords = np.stack([x, y], -1)[None, None] + rng.uniform(-5, 5, (1, len(ii), h, w, 2)).astype(np.float32)
    return fmaps, ii, jj, coords


def main():
    sys.path.insert(0, ROOT)
    from oracle import corr as ocorr
    dbk = types.ModuleType("droid_backends")
    dbk.corr_index_forward = lambda vol, coords, r: (torch.from_numpy(ocorr.corr_index_forward(vol.numpy(), coords.numpy(), r)),)
    dbk.altcorr_forward = lambda f1, f2, coords, r: (torch.from_numpy(ocorr.altcorr_forward(f1.numpy(), f2.numpy(), coords.numpy(), r)),)
    sys.modules["droid_backends"] = dbk
    sys.path.insert(0, REF)
    from networks.modules.corr import AltCorrBlock, CorrBlock
    fmaps, ii, jj, coords = inputs()
    f = torch.from_numpy(fmaps)
    vol = CorrBlock(f[:, ii], f[:, jj])
    out_vol = vol(torch.from_numpy(coords))                           # [1,E,196,h,w]
    alt = AltCorrBlock(f)
    out_alt = alt(torch.from_numpy(coords), torch.from_numpy(ii), torch.from_numpy(jj))   # [1,E,196,h,w]
    np.savez_compressed(os.path.join(HERE, "ref_corr_wrappers.npz"), fmaps=fmaps, ii=ii, jj=jj, coords=coords,
                        volume_lookup=out_vol.numpy(), altcorr=out_alt.numpy())
    print("volume path", tuple(out_vol.shape), "alt path", tuple(out_alt.shape),
          "max |vol - alt|", float((out_vol - out_alt).abs().max()))


if __name__ == "__main__":
    main()
