"""Runs tools/probes/umma_rate_probe.cu on the GPU box:  python tools/probes/run_umma_rate_probe.py > gpurun_out/umma_rate.log

Prints cycles per tcgen05.mma (M128 x N x K16, SS operands) for aligned and row-shifted A descriptors, on one SM and on
all 148 at once, and the L2 -> shared-memory bulk-copy rate per SM with all SMs streaming."""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "tools", "probes", "umma_rate_probe.cu")
SO = "/tmp/umma_rate_probe.so"


def main():
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "nerf_slam_b200", "csrc"),
           "-shared", "-Xcompiler", "-fPIC", SRC, "-o", SO, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout + r.stderr)
        sys.exit(1)
    lib = ctypes.CDLL(SO)
    lib.umma_rate_probe.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    lib.l2_fill_probe.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p]
    lib.l2_fill_box_probe.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p]
    lib.tma_multi_warp_probe.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p]
    if "--multi-only" in sys.argv:
        multi(lib)
        return
    if "--boxes-only" in sys.argv:
        boxes(lib)
        return
    reps = 4096
    for ctas in (1, 148):
        out = np.zeros(ctas, np.int64)
        for N in (16, 64, 128, 256):
            for shift, sbo in ((0, 1024), (1, 1024), (0, 2304), (19, 2304), (38, 2304)):
                rc = lib.umma_rate_probe(N, ctas, reps, shift, sbo, out.ctypes.data)
                if rc:
                    print(f"N {N} ctas {ctas} shift {shift} sbo {sbo}: CUDA error {rc}")
                    continue
                print(f"mma N {N:3d} ctas {ctas:3d} start row {shift:2d} SBO {sbo}: {out.max() / reps:7.1f} cycles/MMA (ideal {N / 2:.0f})")
    out = np.zeros(148, np.int64)
    for shared, span in ((1, 1 << 20), (1, 1 << 18), (0, 1 << 18), (0, 1 << 20)):
        for ctas in (1, 148):
            rc = lib.l2_fill_probe(ctas, 2048, span, shared, out.ctypes.data)
            if rc:
                print(f"fill shared {shared} span {span} ctas {ctas}: CUDA error {rc}")
                continue
            cyc = out[:ctas].max()
            print(f"fill {'same' if shared else 'own '} {span >> 10:5d} KB per CTA, ctas {ctas:3d}: {2048 * 16384 / cyc:6.1f} B/cycle/SM")
    boxes(lib)


def multi(lib):
    """do TMA operations issued by different warps of one SM overlap?"""
    reps = 1024
    for ctas in (1, 148):
        out = np.zeros(ctas * 8, np.int64)
        for nbytes in (4096, 16384, 20480):
            for warps in (1, 2):
                rc = lib.tma_multi_warp_probe(ctas, warps, reps, nbytes, 0, out.ctypes.data)
                if rc:
                    print(f"multi bytes {nbytes} warps {warps} ctas {ctas}: error {rc}"); continue
                cyc = out.reshape(ctas, 8)[:, :warps].max()
                print(f"bulk copies of {nbytes:5d} B, {warps} issuing warp(s), ctas {ctas:3d}: {cyc / reps:6.0f} cycles per copy per warp, "
                      f"{warps * reps * nbytes / cyc:6.1f} B/cycle/SM")
        rc = lib.tma_multi_warp_probe(ctas, 2, reps, 16384, 1, out.ctypes.data)
        if rc == 0:
            c = out.reshape(ctas, 8).max(0)
            print(f"mixed, ctas {ctas:3d}: warp 0 boxes 64c x 16w x 10h: {c[0] / reps:6.0f} cycles per box; warp 1 bulk 16 KB: {c[1] / reps:6.0f} cycles per copy; "
                  f"together {(reps * 20480 + reps * 16384) / max(c[0], c[1]):6.1f} B/cycle/SM")


def boxes(lib):
    """TMA tensor loads of the convolutions' activation boxes (128-byte rows, one global segment each)"""
    out = np.zeros(148, np.int64)
    for C in (128, 448):
        for bw, bh, stages in ((16, 10, 4), (18, 18, 2), (16, 8, 4)):
            for ctas in (1, 148):
                reps = 1024
                rc = lib.l2_fill_box_probe(ctas, reps, C if C % 64 == 0 else 128, bw, bh, stages, out.ctypes.data)
                if rc:
                    print(f"box C {C} {bw}x{bh} ctas {ctas}: CUDA error {rc}")
                    continue
                print(f"box fill C {C:3d} box 64c x {bw:2d}w x {bh:2d}h ({128 * bw * bh / 1024:.1f} KB, {stages} stages) ctas {ctas:3d}: "
                      f"{reps * 128 * bw * bh / out[:ctas].max():6.1f} B/cycle/SM")


if __name__ == "__main__":
    main()
