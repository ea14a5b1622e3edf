"""ctypes front end of the CPU oracle (oracle/pointnet2_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  ratrack_b200/ never imports it.

Functions take and return numpy arrays (fp32 / int32, C-contiguous) with the
reference's layouts; each mirrors one Python entry of
/root/reference/src/lib/pointnet2_utils.py *including* the Python-side glue
(1e10 temp fill :26, zero idx fill :246, sqrt :97,:126).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpointnet2_oracle.so")
_lib = None

_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_knn.restype = ctypes.c_int
        _lib.orc_opt_n_threads.restype = ctypes.c_int
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(_F)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(_I)


def opt_n_threads(n: int) -> int:
    return lib().orc_opt_n_threads(int(n))


def furthest_point_sample(xyz, npoint, return_temp=False):
    """xyz (B,N,3) -> idx (B,npoint) int32 [, temp (B,N): the running minimum the kernel leaves behind].  pointnet2_utils.py:10-36"""
    xyz, px = _f(xyz)
    B, N, _ = xyz.shape
    temp, pt = _f(np.full((B, N), 1e10, dtype=np.float32))
    idx = np.zeros((B, npoint), dtype=np.int32)  # the reference leaves it uninitialised; kernel writes all
    lib().orc_furthest_point_sampling(B, N, int(npoint), px, pt, idx.ctypes.data_as(_I))
    return (idx, temp) if return_temp else idx


def gather_operation(features, idx):
    """features (B,C,N), idx (B,M) -> (B,C,M).  pointnet2_utils.py:39-73"""
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), dtype=np.float32)
    lib().orc_gather_points(B, C, N, M, pf, pi, out.ctypes.data_as(_F))
    return out


def gather_operation_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, M = grad_out.shape
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().orc_gather_points_grad(B, C, int(N), M, pg, pi, out.ctypes.data_as(_F))
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """xyz (B,N,3), new_xyz (B,M,3) -> idx (B,M,nsample).  pointnet2_utils.py:228-256"""
    xyz, px = _f(xyz)
    new_xyz, pn = _f(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), dtype=np.int32)
    lib().orc_ball_query(B, N, M, ctypes.c_float(radius), int(nsample), pn, px, idx.ctypes.data_as(_I))
    return idx


def grouping_operation(features, idx):
    """features (B,C,N), idx (B,P,S) -> (B,C,P,S).  pointnet2_utils.py:184-225"""
    features, pf = _f(features)
    idx, pi = _i(idx)
    B, C, N = features.shape
    _, P, S = idx.shape
    out = np.empty((B, C, P, S), dtype=np.float32)
    lib().orc_group_points(B, C, N, P, S, pf, pi, out.ctypes.data_as(_F))
    return out


def grouping_operation_grad(grad_out, idx, N):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    B, C, P, S = grad_out.shape
    out = np.zeros((B, C, N), dtype=np.float32)
    lib().orc_group_points_grad(B, C, int(N), P, S, pg, pi, out.ctypes.data_as(_F))
    return out


def three_nn_raw(unknown, known):
    """-> (dist2 (B,n,3) f32 *squared*, idx (B,n,3) i32) exactly as the kernel writes them."""
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((B, n, 3), dtype=np.float32)
    idx = np.empty((B, n, 3), dtype=np.int32)
    lib().orc_three_nn(B, n, m, pu, pk, d2.ctypes.data_as(_F), idx.ctypes.data_as(_I))
    return d2, idx


def three_nn(unknown, known):
    """-> (sqrt(dist2), idx).  pointnet2_utils.py:104-133"""
    d2, idx = three_nn_raw(unknown, known)
    with np.errstate(invalid="ignore"):
        return np.sqrt(d2), idx


def knn_raw(k, unknown, known):
    unknown, pu = _f(unknown)
    known, pk = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((B, n, k), dtype=np.float32)
    idx = np.empty((B, n, k), dtype=np.int32)
    rc = lib().orc_knn(B, n, m, int(k), pu, pk, d2.ctypes.data_as(_F), idx.ctypes.data_as(_I))
    if rc != 0:
        raise ValueError("knn: k must be in [0, 200] (interpolate_gpu.cu:30-31)")
    return d2, idx


def knn(k, unknown, known):
    """-> (sqrt(dist2), idx).  pointnet2_utils.py:75-102"""
    d2, idx = knn_raw(k, unknown, known)
    return np.sqrt(d2), idx


def three_interpolate(features, idx, weight):
    """features (B,C,M), idx (B,n,3), weight (B,n,3) -> (B,C,n).  pointnet2_utils.py:136-181"""
    features, pf = _f(features)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, M = features.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), dtype=np.float32)
    lib().orc_three_interpolate(B, C, M, n, pf, pi, pw, out.ctypes.data_as(_F))
    return out


def three_interpolate_grad(grad_out, idx, weight, M):
    grad_out, pg = _f(grad_out)
    idx, pi = _i(idx)
    weight, pw = _f(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, M), dtype=np.float32)
    lib().orc_three_interpolate_grad(B, C, n, int(M), pg, pi, pw, out.ctypes.data_as(_F))
    return out
