"""Top-level alias so `import pointnet2_cuda` (reference: src/lib/pointnet2_utils.py:7) resolves to
the sm_100a kernels: add this directory to sys.path instead of installing the reference extension."""
from ratrack_b200.pointnet2_cuda import *  # noqa: F401,F403
