"""Fused cost-volume kernels of the training path (csrc/costvol_train.cu, lib/costvol_train.py) against the op-by-op chain
of torch ops they replace -- FeatureCorrelator.forward as the reference evaluates it (reference:
src/utils/model_utils/model_utils.py:193-250), here `FeatureCorrelator.fused_rows = False` with every dense layer on torch's
fp32 GEMM -- forward and every gradient (inputs, all convolution / WeightNet parameters).

Tolerances: forward 2e-6 of the output's scale; gradients 2e-5 of each tensor's scale (both sides are fp32 sums of a few
thousand terms in different orders); two runs of the fused path are bit-identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(fused, B, N, seed, dense_tc_on):
    from ratrack_b200 import synthetic
    from ratrack_b200.lib import dense_tc
    from ratrack_b200.model_utils import FeatureCorrelator

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(seed)
    fc = FeatureCorrelator(16, in_channel=256 * 2 + 3, mlp=[256, 256, 256]).cuda()
    with torch.no_grad():
        for p in fc.parameters():                      # biases too: the default init leaves them at zero
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    d = synthetic.make_batch(B, N, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(seed)
    pc1, pc2 = torch.from_numpy(d["pc1"]).cuda(), torch.from_numpy(d["pc2"]).cuda()
    f1 = torch.randn(B, 256, N, device="cuda", generator=g).requires_grad_(True)
    f2 = torch.randn(B, 256, N, device="cuda", generator=g).requires_grad_(True)
    dout = torch.randn(B, 256, N, device="cuda", generator=g)
    old = (FeatureCorrelator.fused_rows, dense_tc.enabled)
    FeatureCorrelator.fused_rows, dense_tc.enabled = fused, dense_tc_on
    try:
        out = fc(pc1, pc2, f1, f2)
        out.backward(dout)
    finally:
        FeatureCorrelator.fused_rows, dense_tc.enabled = old
    grads = {"f1": f1.grad, "f2": f2.grad}
    grads.update({k: p.grad for k, p in fc.named_parameters() if p.grad is not None})
    return out.detach(), {k: v.detach().clone() for k, v in grads.items()}


@pytest.mark.parametrize("B,N", [(2, 256), (3, 320), (1, 1024)])
def test_fused_rows_path_equals_the_op_chain(B, N):
    out_f, g_f = _run(True, B, N, seed=B * 7 + N, dense_tc_on=True)
    out_r, g_r = _run(False, B, N, seed=B * 7 + N, dense_tc_on=False)
    scale = float(out_r.abs().max())
    assert float((out_f - out_r).abs().max()) <= 2e-6 * scale, (float((out_f - out_r).abs().max()), scale)
    assert set(g_f) == set(g_r) and len(g_f) >= 14, sorted(g_f)     # 3 convs + 2 WeightNets (w, b) + the two feature inputs
    worst = {}
    for k in g_r:
        s = float(g_r[k].abs().max())
        worst[k] = float((g_f[k] - g_r[k]).abs().max()) / max(s, 1e-30)
    print({k: f"{v:.1e}" for k, v in worst.items()})
    assert max(worst.values()) <= 2e-5, worst


def test_fused_rows_path_is_bit_repeatable():
    a = _run(True, 2, 256, seed=5, dense_tc_on=True)
    b = _run(True, 2, 256, seed=5, dense_tc_on=True)
    assert torch.equal(a[0], b[0])
    for k in a[1]:
        assert torch.equal(a[1][k], b[1][k]), k


def test_act_grad_and_weighted_sum_kernels_directly():
    """rt_act_grad and the gathered weighted sum against plain torch expressions on odd sizes."""
    from ratrack_b200.lib import costvol_train as cvt

    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(5000, 96, device="cuda", generator=g).requires_grad_(True)
    w = (torch.randn(35, 96, device="cuda", generator=g) / 10).requires_grad_(True)
    b = torch.randn(35, device="cuda", generator=g).requires_grad_(True)
    dy = torch.randn(5000, 35, device="cuda", generator=g)
    y = cvt.linear_act(x, w, b, "leaky")
    y.backward(dy)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    y2 = torch.nn.functional.leaky_relu(torch.nn.functional.linear(x2, w2, b2), 0.1)
    y2.backward(dy)
    for got, ref in ((y, y2), (x.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        assert float((got - ref).abs().max()) <= 3e-6 * float(ref.abs().max())
    # gathered weighted sum, C = 64, K = 5
    B, N, K, C = 2, 77, 5, 64
    xp = torch.randn(B, N, C, device="cuda", generator=g).requires_grad_(True)
    idx = torch.randint(0, N, (B, N, K), device="cuda", generator=g, dtype=torch.int32)
    h2 = torch.rand(B, N, K, 8, device="cuda", generator=g).requires_grad_(True)
    w3 = torch.randn(C, 8, device="cuda", generator=g).requires_grad_(True)
    b3 = torch.randn(C, device="cuda", generator=g).requires_grad_(True)
    dout = torch.randn(B, N, C, device="cuda", generator=g)
    out = cvt.weighted_sum(xp, h2, w3, b3, idx=idx)
    out.backward(dout)
    ref_in = [t.detach().clone().requires_grad_(True) for t in (xp, h2, w3, b3)]
    xg = torch.gather(ref_in[0].unsqueeze(2).expand(B, N, K, C), 1, idx.long().unsqueeze(-1).expand(B, N, K, C))
    ref = (torch.relu(torch.nn.functional.linear(ref_in[1], ref_in[2], ref_in[3])) * xg).sum(2)
    ref.backward(dout)
    assert float((out - ref).abs().max()) <= 3e-6 * float(ref.abs().max())
    for got, r in zip((xp, h2, w3, b3), ref_in):
        assert float((got.grad - r.grad).abs().max()) <= 1e-5 * float(r.grad.abs().max())
