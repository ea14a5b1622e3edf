"""Dense layers of the training path on the tcgen05 tensor cores (csrc/dense_tc.cu).

`linear(x, weight, bias)` has the semantics of `torch.nn.functional.linear` for fp32 CUDA tensors whose last axis is
contiguous, with its three products -- forward X.W^T, dgrad dY.W, wgrad dY^T.X -- evaluated by the library's split-fp16
tensor-core kernels (fp32-class accuracy, bit-repeatable) instead of cuBLAS' fp32 SIMT sgemm, which is what the reference's
1x1 convolutions and Linear layers run on when TF32 is off (reference: src/lib/pytorch_utils.py:35-101,
src/utils/model_utils/model_utils.py:223-231, 308-357, 393-424).

Layers that are not GEMM-shaped (fewer than 16 input or output channels: WeightNet, the 3-channel heads) or tiny stay on
torch's library GEMM -- a plain library call, as the task allows; there is no silent fallback for the covered shapes:
a failed launch raises.
"""
import torch

from .. import _cabi

enabled = True          # module-wide switch (tests / A-B runs); False = torch.nn.functional.linear everywhere
MIN_ROWS = 4096         # below this a launch is latency-bound either way
MIN_CH = 16


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _rows2d(x):
    """(..., C) with unit stride on the last axis -> (rows, C) view + leading dimension, or a contiguous copy."""
    c = x.shape[-1]
    x2 = x.reshape(-1, c)
    if x2.stride(1) != 1 or (x2.shape[0] > 1 and x2.stride(0) < c):
        x2 = x2.contiguous()
    return x2, (x2.stride(0) if x2.shape[0] > 1 else c)


def absmax(x):
    """max |x| as a 1-element device tensor (no host synchronisation); the kernels read it on the device and scale the operand
    by a power of two into the fp16 planes' range, so any finite fp32 input is valid."""
    if not x.is_contiguous():
        return torch.linalg.vector_norm(x.detach(), float("inf")).reshape(1).float()   # strided view: one torch reduction
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    with torch.cuda.device_of(x):
        _cabi.call("rt_absmax", x.data_ptr(), x.numel(), out.data_ptr(), _stream(x))
    return out


def forward_raw(x2, ldx, w, w_sn, w_sk, k, n, bias=None, amax=None, w_amax=None, act=0):
    """y (rows, n) = act(x2 (rows, k; leading dimension ldx) . W'^T + bias), W'[i][j] = w.data_ptr()[i * w_sn + j * w_sk];
    amax / w_amax: device scalars max |x2| / max |w| (None: unscaled x2, 2^10 w -- range-limited, see the header)."""
    rows = x2.shape[0]
    y = torch.empty(rows, n, dtype=torch.float32, device=x2.device)
    with torch.cuda.device_of(x2):
        _cabi.call("rt_dense_tc_forward", rows, k, n, x2.data_ptr(), ldx, w.data_ptr(), w_sn, w_sk,
                   bias.data_ptr() if bias is not None else None, amax.data_ptr() if amax is not None else None,
                   w_amax.data_ptr() if w_amax is not None else None, act, y.data_ptr(), n, _stream(x2))
    return y


class _LinearTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, weight, bias):
        """x2 (rows, k) with unit column stride -> a fresh (rows, n) tensor (never a view: the callers apply in-place
        activations to views of it)."""
        ldx = x2.stride(0) if x2.shape[0] > 1 else x2.shape[1]
        w = weight.contiguous()
        n, k = w.shape
        x_amax, w_amax = absmax(x2), absmax(w)          # one extra read of x: no range restriction on the activations
        y = forward_raw(x2, ldx, w, k, 1, k, n, bias.contiguous() if bias is not None else None, x_amax, w_amax)
        ctx.save_for_backward(x2, w, x_amax, w_amax)
        ctx.ldx = ldx
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, w, x_amax, w_amax = ctx.saved_tensors
        n, k = w.shape
        dy2, lddy = _rows2d(dy)
        rows = dy2.shape[0]
        amax = absmax(dy2)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            # dX = dY . W: the same kernel on dY with the transposed view of W (W'[i][j] = W[j][i])
            dx = forward_raw(dy2, lddy, w, 1, k, n, k, None, amax, w_amax)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(n, k, dtype=torch.float32, device=w.device)
            with torch.cuda.device_of(w):
                _cabi.call("rt_dense_tc_wgrad", rows, n, k, dy2.data_ptr(), lddy, x2.data_ptr(), ctx.ldx, amax.data_ptr(),
                           x_amax.data_ptr(), dw.data_ptr(), _stream(w))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db


def covered(x, weight):
    n, k = weight.shape
    return (enabled and x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and k >= MIN_CH and n >= MIN_CH
            and x.numel() // max(k, 1) >= MIN_ROWS and x.stride(-1) == 1)


def linear(x, weight, bias=None):
    """torch.nn.functional.linear(x, weight, bias) on the tensor-core kernels where the shape is covered."""
    if not covered(x, weight):
        return torch.nn.functional.linear(x, weight, bias)
    x2, _ = _rows2d(x)
    return _LinearTC.apply(x2, weight, bias).view(*x.shape[:-1], weight.shape[0])
