"""Import the UNMODIFIED reference Python model in the GPU-less dev container.

TEST INFRASTRUCTURE ONLY, and usable only where /root/reference exists (this
container).  Used by oracle/gen_golden.py to produce tests/golden/*.npz and by
tests that pin oracle/backbone_oracle.py against the reference's own modules.

What is substituted (and nothing else):
  * absent third-party roots open3d / matplotlib / k3d / glob2 -> MagicMock stubs;
  * the compiled CUDA module `pointnet2_cuda` -> oracle/cpu_pointnet2_cuda.py
    (the C restatement of lib/src/*.cu);
  * torch.Tensor.cuda / nn.Module.cuda -> identity, torch.cuda.FloatTensor /
    IntTensor -> the CPU constructors (the reference hard-codes them at
    utils/model_utils/model_utils.py:38,114,295 and throughout lib/pointnet2_utils.py).
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REF_SRC = "/root/reference/src"
_installed = False


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = {"open3d", "matplotlib", "k3d", "glob2"}

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        m = MagicMock()
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__all__ = []
        return m

    def exec_module(self, module):
        pass


def available() -> bool:
    return os.path.isdir(REF_SRC)


def install():
    """Make `import models`, `from utils.model_utils import *`, `from lib import ...` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present (only the dev container has /root/reference)")
    import torch
    from . import cpu_pointnet2_cuda

    sys.modules["pointnet2_cuda"] = cpu_pointnet2_cuda
    sys.meta_path.append(_StubFinder())
    sys.path.insert(0, REF_SRC)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.cuda.IntTensor = torch.IntTensor
    import models  # noqa: F401  (must precede utils.model_utils: circular import, model_utils.py:7)
    _installed = True


def make_track4d(npoints=512, seed=1234):
    install()
    import torch
    from models.track4d import Track4D
    from utils.parser_util import EasyDict

    torch.manual_seed(seed)
    net = Track4D(EasyDict(dict(npoints=npoints, num_points=256, rigid_thres=0.15, min_obj_points=2)))
    return net
