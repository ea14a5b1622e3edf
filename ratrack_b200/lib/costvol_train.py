"""Fused cost-volume ops of the training path (csrc/costvol_train.cu), channels innermost, as autograd Functions.

FeatureCorrelator.forward (reference: src/utils/model_utils/model_utils.py:193-250) builds (B,515,16,N) and (B,256,16,N)
tensors and walks them with one torch op per step.  Here the same function is

    P1, P2 = per-point projections of the first convolution          (dense_tc GEMMs)
    x1     = cv_layer1(P1, P2, xyz1, xyz2, knn, Wd, b0)               LeakyReLU(P2[nbr] + P1 + Wd.dir + b0), rows (B,N,K,C)
    x2, x3 = dense_tc.linear_act(x, W, b, "leaky")                    activation in the GEMM epilogue
    cost   = weighted_sum(x3, h2, W3, b3)                             sum_k ReLU(W3.h2 + b3) * x3   (WeightNet's last layer in registers)
    out    = weighted_sum(cost, h2', W3', b3', idx=knn11)             the patch-to-patch step gathers its rows on the fly

with hand-written backward kernels (segmented sums over the stable inverse index, fixed-order partial sums: bit-repeatable).
"""
import torch
from torch.autograd import Function

from .. import _cabi
from . import dense_tc


def _st(t):
    return torch.cuda.current_stream(t.device).cuda_stream


class _CvLayer1(Function):
    @staticmethod
    def forward(ctx, p1, p2, xyz1, xyz2, idx, wd, bias):
        B, N1, C = p1.shape
        N2, K = p2.shape[1], idx.shape[2]
        p1, p2, wd = p1.contiguous(), p2.contiguous(), wd.contiguous()
        out = torch.empty(B, N1, K, C, dtype=torch.float32, device=p1.device)
        direction = torch.empty(B, N1, K, 3, dtype=torch.float32, device=p1.device)
        with torch.cuda.device_of(p1):
            _cabi.call("rt_cv1_forward", B, N1, N2, K, C, p1.data_ptr(), p2.data_ptr(), xyz1.data_ptr(), xyz2.data_ptr(), idx.data_ptr(),
                       wd.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(), direction.data_ptr(), _st(p1))
        ctx.save_for_backward(out, direction, idx)
        ctx.dims = (B, N1, N2, K, C)
        ctx.has_bias = bias is not None
        ctx.mark_non_differentiable(direction)
        return out, direction

    @staticmethod
    def backward(ctx, dout, _ddir):
        out, direction, idx = ctx.saved_tensors
        B, N1, N2, K, C = ctx.dims
        dout = dout.contiguous()
        dp1 = torch.empty(B, N1, C, dtype=torch.float32, device=dout.device)
        dp2 = torch.empty(B, N2, C, dtype=torch.float32, device=dout.device)
        dwd = torch.empty(C, 3, dtype=torch.float32, device=dout.device)
        db = torch.empty(C, dtype=torch.float32, device=dout.device)
        with torch.cuda.device_of(dout):
            _cabi.call("rt_cv1_backward", B, N1, N2, K, C, dout.data_ptr(), out.data_ptr(), direction.data_ptr(), idx.data_ptr(),
                       dp1.data_ptr(), dp2.data_ptr(), dwd.data_ptr(), db.data_ptr(), _st(dout))
        return dp1, dp2, None, None, None, dwd, (db if ctx.has_bias else None)


def cv_layer1(p1, p2, xyz1, xyz2, idx, wd, bias):
    """-> (x1 (B,N1,K,C), direction (B,N1,K,3)); p1 (B,N1,C), p2 (B,N2,C) rows, xyz (B,N,3), idx (B,N1,K) int32, wd (C,3)."""
    return _CvLayer1.apply(p1, p2, xyz1.contiguous(), xyz2.contiguous(), idx.contiguous(), wd, bias)


_ACT = {"relu": 1, "leaky": 2}


class _LinearActTC(Function):
    """act(x.W^T + b) with the activation in the GEMM epilogue; backward = one pass g = dy * act'(y) that also yields max |g|
    and the bias gradient, then the dgrad / wgrad GEMMs on g."""

    @staticmethod
    def forward(ctx, x2, weight, bias, act):
        w = weight.contiguous()
        n, k = w.shape
        ldx = x2.stride(0) if x2.shape[0] > 1 else k
        x_amax, w_amax = dense_tc.absmax(x2), dense_tc.absmax(w)
        y = dense_tc.forward_raw(x2, ldx, w, k, 1, k, n, bias.contiguous() if bias is not None else None, x_amax, w_amax, act)
        ctx.save_for_backward(x2, w, x_amax, w_amax, y)
        ctx.ldx, ctx.act, ctx.has_bias = ldx, act, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w, x_amax, w_amax, y = ctx.saved_tensors
        n, k = w.shape
        dy = dy.contiguous()
        rows = dy.shape[0]
        g = torch.empty_like(dy)
        amax = torch.empty(1, dtype=torch.float32, device=dy.device)
        db = torch.empty(n, dtype=torch.float32, device=dy.device)
        with torch.cuda.device_of(dy):
            _cabi.call("rt_act_grad", rows, n, ctx.act, y.data_ptr(), dy.data_ptr(), g.data_ptr(), amax.data_ptr(), db.data_ptr(), _st(dy))
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = dense_tc.forward_raw(g, n, w, 1, k, n, k, None, amax, w_amax)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(n, k, dtype=torch.float32, device=w.device)
            with torch.cuda.device_of(w):
                _cabi.call("rt_dense_tc_wgrad", rows, n, k, g.data_ptr(), n, x2.data_ptr(), ctx.ldx, amax.data_ptr(), x_amax.data_ptr(),
                           dw.data_ptr(), _st(w))
        return dx, dw, (db if ctx.has_bias else None), None


def linear_act(x, weight, bias, act):
    """act(linear(x, weight, bias)) over the last axis of a contiguous (..., k) tensor; act in {"relu", "leaky"} (0.1)."""
    k = x.shape[-1]
    y = _LinearActTC.apply(x.reshape(-1, k), weight, bias, _ACT[act])
    return y.view(*x.shape[:-1], weight.shape[0])


class _WeightedSum(Function):
    @staticmethod
    def forward(ctx, x, idx, h2, w3, b3):
        B, N, K, H = h2.shape
        C = w3.shape[0]
        assert H == 8 and w3.shape[1] == 8
        x, h2, w3, b3 = x.contiguous(), h2.contiguous(), w3.contiguous(), b3.contiguous()
        out = torch.empty(B, N, C, dtype=torch.float32, device=x.device)
        with torch.cuda.device_of(x):
            _cabi.call("rt_wsum_forward", B, N, K, C, x.data_ptr(), idx.data_ptr() if idx is not None else None, h2.data_ptr(),
                       w3.data_ptr(), b3.data_ptr(), out.data_ptr(), _st(x))
        ctx.save_for_backward(x, h2, w3, b3, *([idx] if idx is not None else []))
        ctx.dims = (B, N, K, C)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, h2, w3, b3, *rest = ctx.saved_tensors
        idx = rest[0] if rest else None
        B, N, K, C = ctx.dims
        dout = dout.contiguous()
        dx = torch.empty_like(x)
        dh2 = torch.empty_like(h2)
        dw3 = torch.empty_like(w3)
        db3 = torch.empty_like(b3)
        with torch.cuda.device_of(x):
            _cabi.call("rt_wsum_backward", B, N, K, C, x.data_ptr(), idx.data_ptr() if idx is not None else None, h2.data_ptr(),
                       w3.data_ptr(), b3.data_ptr(), dout.data_ptr(), dx.data_ptr(), dh2.data_ptr(), dw3.data_ptr(), db3.data_ptr(), _st(x))
        return dx, None, dh2, dw3, db3


def weighted_sum(x, h2, w3, b3, idx=None):
    """sum_k ReLU(w3.h2[.,k] + b3) * X[.,k]: x (B,N,K,C) rows, or x (B,N,C) per-point rows gathered through idx (B,N,K) int32;
    h2 (B,N,K,8), w3 (C,8), b3 (C) -> (B,N,C)."""
    return _WeightedSum.apply(x, idx.contiguous() if idx is not None else None, h2, w3, b3)
