"""The C-ABI shared library loads on a CPU-only box and exports every symbol the headers declare."""
import os
import re

from nerf_slam_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    inc = os.path.join(ROOT, "include")
    for f in os.listdir(inc):
        txt = open(os.path.join(inc, f)).read()
        names |= set(re.findall(r"^int\s+(nslam_\w+)\s*\(", txt, flags=re.M))
    return names


def test_library_loads_and_exports_all_declared_symbols(lib):
    declared = _declared()
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    # and the python binding table covers exactly the declared set
    assert set(_lib.exported_symbols()) == declared


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from nerf_slam_b200 import droid_backends
    with pytest.raises(RuntimeError):
        droid_backends.frame_distance(torch.zeros(2, 7), torch.zeros(2, 4, 4), torch.zeros(4),
                                      torch.zeros(1, dtype=torch.long), torch.ones(1, dtype=torch.long), 0.3)


def test_product_code_never_touches_the_oracle():
    """oracle/ is test infrastructure: the package must not import, load or reference it"""
    import re
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nerf_slam_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.\.?oracle)|oracle/_ref|oracle\._ref|build_ref", re.M)
    bad = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):                      # (CUDA sources only mention the oracle in comments)
                txt = open(os.path.join(root, f), errors="ignore").read()
                if pat.search(txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad
