"""Seam B operators on the sm_100a kernels of libratrack_b200.so.

Public surface = the names the reference's modules import from src/lib/pointnet2_utils.py, with the same
argument order, return shapes / dtypes and differentiability:

    furthest_point_sample(xyz (B,N,3), npoint)              -> (B,npoint) int32          reference :10-36
    gather_operation(features (B,C,N), idx (B,M))           -> (B,C,M)     grad -> features        :39-73
    knn(k, unknown (B,n,3), known (B,m,3))                  -> (dist (B,n,k), idx int32)            :75-102
    three_nn(unknown (B,n,3), known (B,m,3))                -> (dist (B,n,3), idx int32)            :104-133
    three_interpolate(features (B,C,m), idx, weight)        -> (B,C,n)     grad -> features        :136-181
    grouping_operation(features (B,C,N), idx (B,P,S))       -> (B,C,P,S)   grad -> features        :184-225
    ball_query(radius, nsample, xyz (B,N,3), new_xyz)       -> (B,npoint,nsample) int32            :228-256
    QueryAndGroup, GroupAll                                                                         :259-318

Index-producing ops carry no gradient in the reference (their backward returns None), so here they are plain
functions evaluated without autograd; the two copy-type ops share one autograd Function (a gather and its
scatter-add), three_interpolate has its own.  `dist` outputs are square roots of the kernels' squared distances,
as in the reference (:97, :126).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import pointnet2_cuda as pointnet2


def _contig(t, what):
    if not t.is_contiguous():
        raise AssertionError(f"{what} must be contiguous")
    return t


def _new(shape, like, dtype):
    return torch.empty(shape, dtype=dtype, device=like.device)


# ---- ops without gradients ---------------------------------------------------------------------------------
@torch.no_grad()
def furthest_point_sample(xyz, npoint):
    B, N, _ = _contig(xyz, "xyz").shape
    idx = _new((B, npoint), xyz, torch.int32)
    running_min = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)   # reference :26
    pointnet2.furthest_point_sampling_wrapper(B, N, npoint, xyz, running_min, idx)
    return idx


@torch.no_grad()
def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = _contig(xyz, "xyz").shape
    M = _contig(new_xyz, "new_xyz").shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)   # rows without a hit stay 0 (:246)
    pointnet2.ball_query_wrapper(B, N, M, radius, nsample, new_xyz, xyz, idx)
    return idx


def _nearest(wrapper, k, unknown, known):
    B, n, _ = _contig(unknown, "unknown").shape
    m = _contig(known, "known").shape[1]
    d2 = _new((B, n, k), unknown, torch.float32)
    idx = _new((B, n, k), unknown, torch.int32)
    if wrapper is pointnet2.knn_wrapper:
        wrapper(B, n, m, k, unknown, known, d2, idx)
    else:
        wrapper(B, n, m, unknown, known, d2, idx)
    return torch.sqrt(d2), idx


@torch.no_grad()
def three_nn(unknown, known):
    return _nearest(pointnet2.three_nn_wrapper, 3, unknown, known)


@torch.no_grad()
def knn(k, unknown, known):
    return _nearest(pointnet2.knn_wrapper, k, unknown, known)


# ---- copy-type ops with gradients ----------------------------------------------------------------------------
class _IndexedCopy(Function):
    """out[b,c,...] = features[b,c,idx[b,...]]; backward scatter-adds into a zero tensor (atomicAdd, as the reference)."""

    @staticmethod
    def forward(ctx, features, idx):
        _contig(features, "features")
        idx = _contig(idx, "idx").int()
        B, C, N = features.shape
        ctx.n_src = N
        ctx.idx = idx
        if idx.dim() == 2:
            out = _new((B, C, idx.shape[1]), features, torch.float32)
            pointnet2.gather_points_wrapper(B, C, N, idx.shape[1], features, idx, out)
        else:
            P, S = idx.shape[1:]
            out = _new((B, C, P, S), features, torch.float32)
            pointnet2.group_points_wrapper(B, C, N, P, S, features, idx, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, g = ctx.idx, grad_out.contiguous()
        B, C = g.shape[:2]
        grad = torch.zeros((B, C, ctx.n_src), dtype=torch.float32, device=g.device)
        if idx.dim() == 2:
            pointnet2.gather_points_grad_wrapper(B, C, ctx.n_src, idx.shape[1], g, idx, grad)
        else:
            pointnet2.group_points_grad_wrapper(B, C, ctx.n_src, idx.shape[1], idx.shape[2], g, idx, grad)
        return grad, None


def gather_operation(features, idx):
    assert idx.dim() == 2
    return _IndexedCopy.apply(features, idx)


def grouping_operation(features, idx):
    assert idx.dim() == 3
    return _IndexedCopy.apply(features, idx)


class _Interpolate3(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        for t, nm in ((features, "features"), (idx, "idx"), (weight, "weight")):
            _contig(t, nm)
        B, C, m = features.shape
        n = idx.shape[1]
        ctx.saved = (idx, weight, m)
        out = _new((B, C, n), features, torch.float32)
        pointnet2.three_interpolate_wrapper(B, C, m, n, features, idx, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, m = ctx.saved
        g = grad_out.contiguous()
        B, C, n = g.shape
        grad = torch.zeros((B, C, m), dtype=torch.float32, device=g.device)
        pointnet2.three_interpolate_grad_wrapper(B, C, n, m, g, idx, weight, grad)
        return grad, None, None


three_interpolate = _Interpolate3.apply

# the reference's autograd class names, for code that spells `<Class>.apply`
GatherOperation = GroupingOperation = _IndexedCopy
ThreeInterpolate = _Interpolate3


class _GroupRows(Function):
    """QueryAndGroup's [grouped_xyz - new_xyz ; grouped_features] produced directly with the channels innermost:
    (B,npoint,nsample,3+C) rows, the layout the dense layers consume -- one kernel instead of two channel-major gathers, a
    subtraction, a cat and the transposing copy in front of the first 1x1 convolution.  Gradient w.r.t. the features only
    (the reference's xyz never requires grad): a bit-repeatable segmented sum."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, features, idx):
        from .. import _cabi

        B, C, N = features.shape
        S, ns = idx.shape[1:]
        feat_rows = features.transpose(1, 2).contiguous()                       # (B,N,C): small next to the grouped tensor
        out = _new((B, S, ns, 3 + C), features, torch.float32)
        with torch.cuda.device_of(features):
            _cabi.call("rt_group_rows", B, C, N, S, ns, _contig(xyz, "xyz").data_ptr(), _contig(new_xyz, "new_xyz").data_ptr(),
                       feat_rows.data_ptr(), _contig(idx, "idx").data_ptr(), out.data_ptr(),
                       torch.cuda.current_stream(features.device).cuda_stream)
        ctx.idx = idx
        ctx.shape = (B, C, N, S, ns)
        return out

    @staticmethod
    def backward(ctx, grad_rows):
        from .. import _cabi

        B, C, N, S, ns = ctx.shape
        g = grad_rows.contiguous()
        grad = _new((B, N, C), g, torch.float32)
        with torch.cuda.device_of(g):
            _cabi.call("rt_group_rows_grad", B, C, N, S, ns, g.data_ptr(), ctx.idx.data_ptr(), grad.data_ptr(),
                       torch.cuda.current_stream(g.device).cuda_stream)
        return None, None, grad.transpose(1, 2), None


# ---- grouping modules ------------------------------------------------------------------------------------------
class QueryAndGroup(nn.Module):
    """ball_query -> [grouped xyz - centre ; grouped features] with the xyz channels first (reference :259-292)."""

    rows_layout = True   # class-wide switch: False = the reference's op-by-op chain (channel-major result)

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        if (self.rows_layout and features is not None and self.use_xyz and features.is_cuda and features.dtype == torch.float32
                and not xyz.requires_grad and not new_xyz.requires_grad):
            # same (B,3+C,npoint,nsample) result, stored channels-innermost (torch.channels_last strides)
            return _GroupRows.apply(xyz, new_xyz, features.contiguous(), idx).permute(0, 3, 1, 2)
        rel = grouping_operation(xyz.transpose(1, 2).contiguous(), idx) - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return rel
        grouped = grouping_operation(features, idx)
        return torch.cat([rel, grouped], dim=1) if self.use_xyz else grouped


class GroupAll(nn.Module):
    """One group holding every point (reference :295-318)."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        everything = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return everything
        feats = features.unsqueeze(2)
        return torch.cat([everything, feats], dim=1) if self.use_xyz else feats
