"""Load the UNMODIFIED reference extension built by oracle/build_ref.py (oracle/_ref/pointnet2_cuda.so).

TEST INFRASTRUCTURE ONLY.  Needs a CUDA device to run anything.  Returns the pybind module with
the reference's ten entry points (/root/reference/src/lib/src/pointnet2_api.cpp:11-24) or None.
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "pointnet2_cuda.so")
_mod = None


def load():
    global _mod
    if _mod is not None:
        return _mod
    if not os.path.exists(SO):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location("pointnet2_cuda", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _mod = mod
    return mod
