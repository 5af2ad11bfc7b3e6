"""Runs tools/probes/umma_row_shift_probe.cu on the GPU box (compiles it with the box's nvcc, a few seconds).

  python tools/probes/run_umma_probe.py > gpurun_out/umma_probe.log

For every (shift s, base_offset, group stride SBO) prints whether the MMA read rows s, s+1, ... with the right
swizzle phase.  SBO = 1024: the M rows are 128 consecutive smem rows; SBO = 2048: 16 groups of 8 rows, 16 rows
apart (the 8-px-wide tile inside a 16-px-wide halo box)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "tools", "probes", "umma_row_shift_probe.cu")
SO = "/tmp/umma_probe.so"


def main():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "nerf_slam_b200", "csrc"),
           "-shared", "-Xcompiler", "-fPIC", SRC, "-o", SO, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout + r.stderr)
        sys.exit(1)
    lib = ctypes.CDLL(SO)
    lib.umma_row_shift_probe.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    out = np.zeros((128, 4), np.float32)
    # SBO 1024: 128 consecutive rows; 2048: 8-px tile in a 16-px-wide box; 1280: 8-px-wide tile in a 10-px-wide halo box
    # (16h x 8w tile, ONE box {64c,10w,18h} for all nine taps: start row = dy*10 + dx); 2304: 16-px tile rows in an 18-px box
    for sbo in (1024, 1280, 2048, 2304):
        for s in (0, 1, 2, 3, 5, 7, 8, 9, 10, 11, 12, 20, 21, 22):
            for bo in sorted({0, s % 8}):
                rc = lib.umma_row_shift_probe(s, bo, sbo, out.ctypes.data)
                if rc != 0:
                    print(f"sbo {sbo} shift {s} base_offset {bo}: CUDA error {rc}")
                    continue
                m = np.arange(128)
                want = s + (m // 8) * (sbo // 128) + (m % 8)
                rows_ok = bool((out[:, 0] == want).all() and (out[:, 2] == want).all())
                phase_ok = bool((out[:, 1] == 0).all() and (out[:, 3] == 1).all())
                print(f"sbo {sbo} shift {s} base_offset {bo}: rows {'OK' if rows_ok else 'WRONG'} swizzle {'OK' if phase_ok else 'WRONG'}"
                      + ("" if rows_ok and phase_ok else f"   first rows read {out[:10, 0].astype(int).tolist()} chunk ids {out[:10, 1].astype(int).tolist()} / {out[:10, 3].astype(int).tolist()}"))


if __name__ == "__main__":
    main()
