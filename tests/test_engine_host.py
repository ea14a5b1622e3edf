"""Host logic of the fused engine that needs no GPU (ratrack_b200/engine.py): BatchNorm folding, the project-then-gather
split of first-layer weights, the fp16 hi/lo operand packing and the size / order of the weight table the C ABI expects."""
import torch
import torch.nn as nn

from ratrack_b200 import engine, synthetic
from ratrack_b200.model_utils import Track4DBackbone


class Args:
    npoints = 512


def test_fold_equals_conv_then_eval_batchnorm():
    torch.manual_seed(0)
    conv = nn.Conv2d(7, 5, 1, bias=False)
    bn = nn.BatchNorm2d(5)
    bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(); bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0)
    bn.eval()
    x = torch.randn(3, 7, 4, 2)
    w, b = engine._fold(conv.weight, bn)
    got = torch.einsum("oc,bchw->bohw", w, x) + b.view(1, -1, 1, 1)
    assert float((got - bn(conv(x))).abs().max()) <= 1e-5


def test_weight_table_size_and_project_then_gather_identity():
    net = Track4DBackbone(Args())
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    net.eval()
    ws = engine.build_weight_table(net)
    assert len(ws) == 157                                              # struct EngineW, csrc/engine.cu
    # SA1 scale 0 of pn_head: W'.[dxyz; f] == WX.dxyz + WF.f   (WF rows of both scales are stacked: first 16 = scale 0)
    layer = net.pn_head.sa1.mlps[0][0]
    w_full, b_full = engine._fold(layer.conv.weight, layer.bn.bn)
    wf_ft, wx, b1 = ws[0], ws[4], ws[5]
    dxyz, f = torch.randn(3), torch.randn(2)
    assert float((w_full @ torch.cat([dxyz, f]) + b_full - (wx @ dxyz + wf_ft[:16] @ f + b1)).abs().max()) <= 1e-5
    # cost-volume first conv: the five column blocks tile the 515 input channels
    w1 = net.fc_layer.mlp_convs[0].weight.detach().flatten(1)
    base = 2 * 56
    assert torch.equal(torch.cat(ws[base:base + 5], dim=1), w1)


def test_umma_planes_hold_the_weights_to_22_bits_in_core_matrix_order():
    torch.manual_seed(1)
    w = torch.randn(16, 32)
    p = engine._umma_planes(w, 32)                                     # [chunk 1][hi, lo][kc 4][row group 2][8][8]
    assert p.shape == (1, 2, 4, 2, 8, 8) and p.dtype == torch.float16
    hi, lo = p[0, 0].float(), p[0, 1].float()
    for n, k in ((0, 0), (3, 9), (15, 31), (8, 16)):
        v = hi[k // 8, n // 8, n % 8, k % 8] + lo[k // 8, n // 8, n % 8, k % 8] / engine.LO_SCALE   # lo plane is stored x 2^11
        assert abs(float(v) - float(w[n, k]) * engine.W_SCALE) <= abs(float(w[n, k])) * engine.W_SCALE * 2 ** -20
    wc = engine.pack_weightnet_last(torch.randn(256, 8))
    assert wc.shape == (2, 2, 32, 8, 8)                                # K padded 8 -> 16
