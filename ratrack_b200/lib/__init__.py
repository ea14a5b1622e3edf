"""Seam B: mirrors of the reference's `lib` package (pointnet2_utils, pointnet2_modules, pytorch_utils)."""
