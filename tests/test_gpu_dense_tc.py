"""Tensor-core dense layers of the training path (csrc/dense_tc.cu, lib/dense_tc.py) against an fp64 product of the same
operands, next to torch's own fp32 product (cuBLAS sgemm, TF32 off) -- the arithmetic the reference trains on
(reference: src/lib/pytorch_utils.py:35-101, src/utils/model_utils/model_utils.py:223-231).

Tolerances (max |err| / max |ref|): forward and dgrad 2e-6 -- the level of the fp32 sgemm itself, which the test prints and
asserts to be within a factor 3; wgrad 2e-4 relative for up to 2^22 rows (the tensor core truncates its fp32 accumulator
at every K step, and a wgrad chain is thousands of steps long: the error grows linearly with the rows per CTA, measured
1.9e-5 at 2^20 rows)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, ref):
    return float((a.detach().double() - ref).abs().max() / ref.abs().max().clamp_min(1e-300))


@pytest.fixture(autouse=True)
def _fp32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


def _case(rows, k, n, gscale=1.0, bias=True, seed=0):
    from ratrack_b200.lib import dense_tc

    g = torch.Generator(device="cuda").manual_seed(seed + rows + 131 * k + 17 * n)
    x = torch.randn(rows, k, device="cuda", generator=g) * 2.0
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g) if bias else None
    dy = torch.randn(rows, n, device="cuda", generator=g) * gscale
    dy[::7] *= 1e-3                                   # gradients span orders of magnitude
    xt, wt = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    bt = b.clone().requires_grad_(True) if bias else None
    dense_tc.MIN_ROWS, dense_tc.MIN_CH = 0, 1
    y = dense_tc.linear(xt, wt, bt)
    dense_tc.MIN_ROWS, dense_tc.MIN_CH = 4096, 16
    y.backward(dy)
    x2, w2 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    y2 = torch.nn.functional.linear(x2, w2, b)
    y2.backward(dy)
    xr, wr, dyr = x.double(), w.double(), dy.double()
    ref = {"fwd": xr @ wr.t() + (b.double() if bias else 0), "dx": dyr @ wr, "dw": dyr.t() @ xr}
    ours = {"fwd": y, "dx": xt.grad, "dw": wt.grad}
    torch_ = {"fwd": y2, "dx": x2.grad, "dw": w2.grad}
    errs = {key: (_rel(ours[key], ref[key]), _rel(torch_[key], ref[key])) for key in ref}
    if bias:
        assert _rel(bt.grad, dyr.sum(0)) < 2e-6
    return errs, (y, xt.grad, wt.grad)


@pytest.mark.parametrize("rows,k,n,gscale,bias", [
    (4096, 64, 64, 1.0, True),
    (128, 16, 16, 1.0, True),          # one tile, one K step
    (1000, 16, 16, 1.0, False),        # ragged last tile
    (33333, 35, 32, 1.0, False),       # SA2 first layer: unaligned rows (scalar loads), K padded 35 -> 48
    (20000, 67, 64, 1e-6, True),       # SA3 first layer, tiny gradients (scaled by their absolute maximum)
    (10000, 520, 16, 1.0, False),      # mse SA1 first layer: 9 K chunks, two x tiles in wgrad
    (10000, 128, 35, 1.0, True),       # N not a multiple of 16: padded columns, bounded stores
    (50000, 256, 256, 1e-7, True),     # cost-volume layer: two N tiles, two dy tiles in wgrad
    (30001, 96, 64, 1.0, True),
    (30000, 160, 128, 1.0, True),
])
def test_forward_dgrad_wgrad_match_fp64(rows, k, n, gscale, bias):
    errs, _ = _case(rows, k, n, gscale, bias)
    print(rows, k, n, {key: f"{a:.2e} (torch {b:.2e})" for key, (a, b) in errs.items()})
    for key in ("fwd", "dx"):
        ours, th = errs[key]
        assert ours < 2e-6 and ours < 3 * th + 1e-7, (key, ours, th)
    ours, th = errs["dw"]
    assert ours < 2e-4 and ours < 10 * th + 2e-6, ("dw", ours, th)


def test_long_reduction_and_bit_repeatability():
    errs, out1 = _case(1 << 20, 64, 64, 1e-4, True, seed=3)
    _, out2 = _case(1 << 20, 64, 64, 1e-4, True, seed=3)
    assert errs["dw"][0] < 2e-4 and errs["fwd"][0] < 2e-6 and errs["dx"][0] < 2e-6, errs
    for a, b in zip(out1, out2):
        assert torch.equal(a, b)      # fixed split of the rows, fixed order of the partial sums: no atomics anywhere


def test_strided_rows_and_dispatch():
    """The module-level entry takes any (..., C) tensor with unit stride on the last axis (the channels-innermost views
    PointwiseConv2d hands over) and sends narrow / tiny layers to torch's library GEMM."""
    from ratrack_b200.lib import dense_tc

    g = torch.Generator(device="cuda").manual_seed(5)
    big = torch.randn(6000, 96, device="cuda", generator=g)
    x = big[:, 4:68]                                    # leading dimension 96, 64 columns, 16-byte aligned only
    w = torch.randn(32, 64, device="cuda", generator=g) / 8
    assert dense_tc.covered(x, w)
    y = dense_tc.linear(x, w)
    assert _rel(y, x.double() @ w.double().t()) < 2e-6
    x4 = torch.randn(8, 32, 20, 64, device="cuda", generator=g)
    y4 = dense_tc.linear(x4, w)
    assert y4.shape == (8, 32, 20, 32) and _rel(y4, x4.double() @ w.double().t()) < 2e-6
    assert not dense_tc.covered(torch.randn(100000, 3, device="cuda"), torch.randn(8, 3, device="cuda"))       # WeightNet
    assert not dense_tc.covered(torch.randn(64, 64, device="cuda"), torch.randn(64, 64, device="cuda"))        # tiny


def test_out_of_range_operand_is_loud():
    from ratrack_b200.lib import dense_tc

    x = torch.ones(4096, 64, device="cuda")
    x[5, 3] = 1.0e5                                      # beyond fp16: the hi plane holds inf
    w = torch.ones(64, 64, device="cuda") / 64
    y = dense_tc.forward_raw(x, 64, w, 64, 1, 64, 64)    # raw call without the operands' maxima: range-limited
    assert not bool(torch.isfinite(y[5]).all())          # never a silently saturated value
    assert bool(torch.isfinite(y[6]).all())
    amax = dense_tc.absmax(x)
    assert float(amax) == 1.0e5
    y2 = dense_tc.forward_raw(x, 64, w, 64, 1, 64, 64, None, amax, dense_tc.absmax(w))   # scaled operands: any finite input
    assert _rel(y2, x.double() @ w.double().t()) < 2e-6
    # the module-level entry always scales: activations ~1e5 and weights ~1e3 behave like fp32
    xb, wb = x * 3.0, w * 6.4e4
    assert _rel(dense_tc.linear(xb, wb), xb.double() @ wb.double().t()) < 2e-6
