"""Seam A: drop-in for the reference's compiled module `pointnet2_cuda`.

Same ten entry points, positional arguments, in-place output convention and return values as
the pybind module of the reference (src/lib/src/pointnet2_api.cpp:11-24): tensors are CUDA,
contiguous, fp32 / int32; outputs are pre-allocated by the caller.  Work is enqueued on torch's
current CUDA stream of the tensors' device, no synchronisation.

Put `ratrack_b200/compat` on sys.path (it holds a one-line `pointnet2_cuda.py` re-export) and the
unmodified reference Python (`import pointnet2_cuda as pointnet2`, src/lib/pointnet2_utils.py:7)
runs on these kernels.  Errors raise RuntimeError (the reference prints and calls exit(-1)).
"""
import torch

from . import _cabi


def _chk(t, dtype, name):
    # the reference checks is_cuda + contiguity only in ball_query (ball_query.cpp:14-21) and lets
    # .data<T>() throw on dtype mismatch; here every op checks all three.
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDAtensor ")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous ")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _f(t, name):
    return _chk(t, torch.float32, name)


def _i(t, name):
    return _chk(t, torch.int32, name)


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    with torch.cuda.device_of(xyz):
        _cabi.call("rt_ball_query", b, n, m, float(radius), nsample, _f(new_xyz, "new_xyz"), _f(xyz, "xyz"),
                   _i(idx, "idx"), _stream(xyz))
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    with torch.cuda.device_of(points):
        _cabi.call("rt_group_points", b, c, n, npoints, nsample, _f(points, "points"), _i(idx, "idx"),
                   _f(out, "out"), _stream(points))
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    with torch.cuda.device_of(grad_out):
        _cabi.call("rt_group_points_grad", b, c, n, npoints, nsample, _f(grad_out, "grad_out"), _i(idx, "idx"),
                   _f(grad_points, "grad_points"), _stream(grad_out))
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    with torch.cuda.device_of(points):
        _cabi.call("rt_gather_points", b, c, n, npoints, _f(points, "points"), _i(idx, "idx"), _f(out, "out"),
                   _stream(points))
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    with torch.cuda.device_of(grad_out):
        _cabi.call("rt_gather_points_grad", b, c, n, npoints, _f(grad_out, "grad_out"), _i(idx, "idx"),
                   _f(grad_points, "grad_points"), _stream(grad_out))
    return 1


def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    with torch.cuda.device_of(points):
        _cabi.call("rt_furthest_point_sampling", b, n, m, _f(points, "points"), _f(temp, "temp"), _i(idx, "idx"),
                   _stream(points))
    return 1


def knn_wrapper(b, n, m, k, unknown, known, dist2, idx):
    with torch.cuda.device_of(unknown):
        _cabi.call("rt_knn", b, n, m, k, _f(unknown, "unknown"), _f(known, "known"), _f(dist2, "dist2"),
                   _i(idx, "idx"), _stream(unknown))


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    with torch.cuda.device_of(unknown):
        _cabi.call("rt_three_nn", b, n, m, _f(unknown, "unknown"), _f(known, "known"), _f(dist2, "dist2"),
                   _i(idx, "idx"), _stream(unknown))


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    with torch.cuda.device_of(points):
        _cabi.call("rt_three_interpolate", b, c, m, n, _f(points, "points"), _i(idx, "idx"), _f(weight, "weight"),
                   _f(out, "out"), _stream(points))


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    with torch.cuda.device_of(grad_out):
        _cabi.call("rt_three_interpolate_grad", b, c, n, m, _f(grad_out, "grad_out"), _i(idx, "idx"),
                   _f(weight, "weight"), _f(grad_points, "grad_points"), _stream(grad_out))
