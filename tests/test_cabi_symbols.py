"""The C-ABI library loads on a GPU-less host and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

from conftest import ROOT


def _declared():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_the_ten_pointnet2_entries():
    d = _declared()
    for n in ["rt_ball_query", "rt_group_points", "rt_group_points_grad", "rt_gather_points",
              "rt_gather_points_grad", "rt_furthest_point_sampling", "rt_knn", "rt_three_nn",
              "rt_three_interpolate", "rt_three_interpolate_grad"]:
        assert n in d


def test_library_exports_every_declared_symbol():
    from ratrack_b200 import _cabi, build

    build.build()
    lib = ctypes.CDLL(_cabi.SO_PATH)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    lib.rt_abi_version.restype = ctypes.c_int
    assert lib.rt_abi_version() >= 1


def test_ctypes_signatures_cover_the_header():
    from ratrack_b200 import _cabi

    decl = set(_declared()) - {"rt_abi_version", "rt_last_error"}
    bound = set(_cabi.SIGNATURES) | set(_cabi.OTHER)
    assert decl == bound, decl ^ bound


def test_argument_errors_do_not_need_a_gpu():
    """Argument validation happens before any launch: the reference exits the process on errors
    (lib/src/ball_query_gpu.cu:62-65); this library returns a code + message."""
    from ratrack_b200 import _cabi

    L = _cabi.lib()
    rc = L.rt_knn(1, 4, 4, 201, 16, 16, 16, 16, None)   # k > 200 (interpolate_gpu.cu:30-31)
    assert rc < 0 and b"200" in L.rt_last_error()
    rc = L.rt_ball_query(1, 4, 4, 1.0, 4, None, None, None, None)
    assert rc < 0


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (test infrastructure only)."""
    bad = []
    for f in glob.glob(os.path.join(ROOT, "ratrack_b200", "**", "*.py"), recursive=True):
        for line in open(f):
            if re.match(r"\s*(from|import)\s+(\.+)?oracle\b", line) or "import oracle" in line:
                bad.append((f, line.strip()))
    assert not bad, bad
