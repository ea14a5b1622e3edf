"""ratrack_b200 -- B200-native (sm_100a) implementation of RaTrack's per-frame point-cloud hot path.

Layout:
  csrc/                 hand-written CUDA kernels + the C ABI (include/ratrack_b200.h)
  _cabi.py              ctypes binding of libratrack_b200.so (no fallback: raises if missing)
  pointnet2_cuda.py     Seam A -- drop-in for the reference's compiled module `pointnet2_cuda`
  lib/                  Seam B -- pointnet2_utils / pointnet2_modules / pytorch_utils mirrors
  model_utils.py        Seam C -- PNHead, FeatureCorrelator, FlowDecoder, ... + Track4D backbone
  engine.py             fused inference engine (eval-mode Track4D.backbone in CUDA kernels)
  track4d.py            Track4D drop-in: backbone + DBSCAN clustering + affinity + Sinkhorn association
  losses.py, metrics.py, train.py, sharding.py   loss terms, device metrics, training step, multi-GPU sharding
  data_io.py, main_utils.py   radar .bin reader / batchers / result lines; frame order of the dataset class + evaluation loop
  synthetic.py          deterministic VoD-shaped synthetic frame pairs
"""
__version__ = "0.1.0"
