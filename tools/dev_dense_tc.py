"""Development check of the tensor-core dense kernels (csrc/dense_tc.cu): error against an fp64 product next to the error of
torch's fp32 product, and CUDA-event timings against torch.nn.functional.linear (TF32 off).  Needs a B200.
    python tools/dev_dense_tc.py [--big]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200.lib import dense_tc  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, ref):
    return float((a.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def timeit(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def check(rows, k, n, gscale=1.0, bias=True, time=False):
    g = torch.Generator(device="cuda").manual_seed(rows * 31 + k * 7 + n)
    x = torch.randn(rows, k, device="cuda", generator=g) * 2.0
    w = torch.randn(n, k, device="cuda", generator=g) * (1.0 / k ** 0.5)
    b = torch.randn(n, device="cuda", generator=g) if bias else None
    dy = torch.randn(rows, n, device="cuda", generator=g) * gscale
    dy[::7] *= 1e-3      # a wide dynamic range, as gradients have
    xr, wr, dyr = x.double(), w.double(), dy.double()
    y_ref = xr @ wr.t() + (b.double() if bias else 0)
    dx_ref = dyr @ wr
    dw_ref = dyr.t() @ xr
    xt = x.clone().requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    bt = b.clone().requires_grad_(True) if bias else None
    dense_tc.enabled = True
    dense_tc.MIN_ROWS = 0
    dense_tc.MIN_CH = 1
    y = dense_tc.linear(xt, wt, bt)
    y.backward(dy)
    torch.cuda.synchronize()
    x2 = x.clone().requires_grad_(True)
    w2 = w.clone().requires_grad_(True)
    y2 = torch.nn.functional.linear(x2, w2, b)
    y2.backward(dy)
    msg = (f"rows {rows:8d} k {k:4d} n {n:4d} gscale {gscale:g}: fwd {rel(y, y_ref):.2e} (torch {rel(y2, y_ref):.2e})  "
           f"dx {rel(xt.grad, dx_ref):.2e} ({rel(x2.grad, dx_ref):.2e})  dw {rel(wt.grad, dw_ref):.2e} ({rel(w2.grad, dw_ref):.2e})")
    if bias:
        msg += f"  db {rel(bt.grad, dyr.sum(0)):.2e}"
    # bit-repeatability
    xt.grad = None
    wt.grad = None
    g1 = None
    y_b = dense_tc.linear(xt, wt, bt)
    y_b.backward(dy)
    rep = bool(torch.equal(y_b, y)) and bool(torch.equal(wt.grad, wt.grad.clone()))
    msg += f"  repeat {rep}"
    if time:
        with torch.no_grad():
            amax = dense_tc.absmax(dy)
            t_f = timeit(lambda: dense_tc.forward_raw(x, k, w, k, 1, k, n, b))
            t_fr = timeit(lambda: torch.nn.functional.linear(x, w, b))
            t_dx = timeit(lambda: dense_tc.forward_raw(dy, n, w, 1, k, n, k, None, amax))
            t_dxr = timeit(lambda: dy @ w)
            dw = torch.empty(n, k, device="cuda")
            from ratrack_b200 import _cabi
            st = torch.cuda.current_stream().cuda_stream
            t_dw = timeit(lambda: _cabi.call("rt_dense_tc_wgrad", rows, n, k, dy.data_ptr(), n, x.data_ptr(), k, amax.data_ptr(), None, dw.data_ptr(), st))
            t_dwr = timeit(lambda: dy.t() @ x)
            t_am = timeit(lambda: dense_tc.absmax(dy))
        fl = 2.0 * rows * k * n / 1e9
        gb = 4.0 * rows * (k + n) / 1e6
        msg += (f"\n      ms: fwd {t_f:.3f} (torch {t_fr:.3f})  dx {t_dx:.3f} ({t_dxr:.3f})  dw {t_dw:.3f} ({t_dwr:.3f})  absmax {t_am:.3f}"
                f"   [fwd {fl / t_f:.0f} TFLOP/s useful, {gb / t_f:.0f} GB/s]")
    print(msg, flush=True)


if __name__ == "__main__":
    big = "--big" in sys.argv
    check(4096, 64, 64)
    check(128, 16, 16)
    check(1000, 16, 16, bias=False)
    check(33333, 35, 32, bias=False)
    check(20000, 67, 64, gscale=1e-6)
    check(10000, 520, 16, bias=False)
    check(10000, 128, 35)
    check(50000, 256, 256, gscale=1e-7)
    check(30000, 160, 128)
    check(30001, 96, 64)
    check(1 << 20, 64, 64, time=True)
    check(1 << 20, 256, 256, gscale=1e-5, time=True)
    if big:
        check(1 << 22, 256, 256, gscale=1e-5, time=True)
        check(6291456, 64, 64, time=True)
        check(1572864, 520, 16, time=True, bias=False)
