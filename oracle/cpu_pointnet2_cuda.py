"""CPU stand-in for the compiled module `pointnet2_cuda` (Seam A), backed by the C oracle.

TEST INFRASTRUCTURE ONLY.  Same ten entry points, argument order and in-place
output convention as /root/reference/src/lib/src/pointnet2_api.cpp:11-24, but
operating on CPU torch tensors.  oracle/ref_harness.py installs it as
sys.modules['pointnet2_cuda'] so the UNMODIFIED reference Python
(lib/pointnet2_utils.py, lib/pointnet2_modules.py, utils/model_utils) runs in the
GPU-less dev container to produce tests/golden/*.
"""
import ctypes

import torch

from . import pointnet2_oracle as _o

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def _pf(t):
    assert t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), _F)


def _pi(t):
    assert t.dtype == torch.int32 and t.is_contiguous() and t.device.type == "cpu"
    return ctypes.cast(t.data_ptr(), _I)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    _o.lib().orc_ball_query(b, n, m, ctypes.c_float(radius), nsample, _pf(new_xyz), _pf(xyz), _pi(idx))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    _o.lib().orc_group_points(b, c, n, npoints, nsample, _pf(points), _pi(idx), _pf(out))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _o.lib().orc_group_points_grad(b, c, n, npoints, nsample, _pf(grad_out), _pi(idx), _pf(grad_points))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    _o.lib().orc_gather_points(b, c, n, npoints, _pf(points), _pi(idx), _pf(out))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    _o.lib().orc_gather_points_grad(b, c, n, npoints, _pf(grad_out), _pi(idx), _pf(grad_points))
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    _o.lib().orc_furthest_point_sampling(b, n, m, _pf(points), _pf(temp), _pi(idx))
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    rc = _o.lib().orc_knn(b, n, m, k, _pf(unknown), _pf(known), _pf(dist2), _pi(idx))
    if rc != 0:
        raise ValueError("k > 200")


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    _o.lib().orc_three_nn(b, n, m, _pf(unknown), _pf(known), _pf(dist2), _pi(idx))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    _o.lib().orc_three_interpolate(b, c, m, n, _pf(points), _pi(idx), _pf(weight), _pf(out))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    _o.lib().orc_three_interpolate_grad(b, c, n, m, _pf(grad_out), _pi(idx), _pf(weight), _pf(grad_points))
