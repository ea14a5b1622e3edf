"""SharedMLP and its 1x1-conv building block -- interface mirror of the reference's
src/lib/pytorch_utils.py (SharedMLP :5-32, _ConvBase :35-101, BatchNorm2d :120-123, Conv2d :163-197).

Only the state_dict surface and the call semantics are kept: a SharedMLP is a Sequential of
`layer{i}` blocks, each a Sequential with children `conv` (nn.Conv2d 1x1, bias only when there is
no BN, Kaiming-normal init), `bn.bn` (nn.BatchNorm2d, weight 1 / bias 0) and `activation`, so
reference checkpoints load key-for-key (`mlps.0.layer0.conv.weight`, `...layer0.bn.bn.running_mean`).
The fused engine reads these parameters directly and folds BN in eval mode.
"""
from typing import List

import torch
import torch.nn as nn

from . import dense_tc


class PointwiseConv2d(nn.Conv2d):
    """nn.Conv2d whose 1x1 / stride-1 case is evaluated as what it is: ONE GEMM over the channel axis, rows = all
    B*H*W positions (same parameters, same state_dict keys, same fp32 result up to summation order).  With TF32 off --
    which parity with the fp32 reference needs -- cuDNN serves 1x1 convolutions with its generic SIMT convolution
    engines; torch.profiler on the training step showed 53 % of the device time in those forward / wgrad / dgrad
    kernels.  Here the activation is kept channels-innermost (torch.channels_last; the cost volume's concatenated
    tensor already is), so forward, dX and dW are each a single large cuBLAS sgemm and BatchNorm / ReLU / max-pool run
    their NHWC kernels on the result without a layout copy."""

    use_gemm = True   # class-wide switch: False = plain nn.Conv2d (cuDNN), what the reference's own modules do

    def forward(self, x):
        if (not PointwiseConv2d.use_gemm or self.kernel_size != (1, 1) or self.stride != (1, 1) or self.padding != (0, 0) or self.groups != 1 or
                self.dilation != (1, 1) or x.dim() != 4):
            return super().forward(x)
        dims = [0, 2, 3]
        if x.stride(1) == 1:
            dims.sort(key=lambda d: -x.stride(d))                       # memory order of the non-channel axes
        rows = x.permute(*dims, 1)
        if not rows.is_contiguous():                                   # e.g. NCHW from a gather kernel: one copy, then
            x = x.contiguous(memory_format=torch.channels_last)        # every later layer stays channels-innermost
            dims = [0, 2, 3]
            rows = x.permute(0, 2, 3, 1)
        # forward / dgrad / wgrad on the tcgen05 kernels (dense_tc) for GEMM-shaped layers, torch's library GEMM for the rest
        y = dense_tc.linear(rows, self.weight.view(self.out_channels, self.in_channels), self.bias)
        src = dims + [1]                                               # y's axes in terms of (B,C,H,W) axes
        return y.permute(*[src.index(d) for d in range(4)])            # (B,Cout,H,W) view, channels still innermost


class BatchNorm2d(nn.Sequential):
    def __init__(self, in_size: int, name: str = ""):
        super().__init__()
        self.add_module(name + "bn", nn.BatchNorm2d(in_size))
        nn.init.constant_(self[0].weight, 1.0)
        nn.init.constant_(self[0].bias, 0)


class Conv2d(nn.Sequential):
    def __init__(self, in_size: int, out_size: int, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=nn.ReLU(inplace=True), bn: bool = False, init=nn.init.kaiming_normal_,
                 bias: bool = True, preact: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        bias = bias and (not bn)
        conv = PointwiseConv2d(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
        init(conv.weight)
        if bias:
            nn.init.constant_(conv.bias, 0)
        norm_width = in_size if preact else out_size
        post = []
        if bn:
            post.append((name + "bn", BatchNorm2d(norm_width)))
        if activation is not None:
            post.append((name + "activation", activation))
        if not bn and instance_norm:
            post.append((name + "in", nn.InstanceNorm2d(norm_width, affine=False, track_running_stats=False)))
        if preact:
            for k, m in post:
                self.add_module(k, m)
        self.add_module(name + "conv", conv)
        if not preact:
            for k, m in post:
                self.add_module(k, m)


class SharedMLP(nn.Sequential):
    def __init__(self, args: List[int], *, bn: bool = False, activation=nn.ReLU(inplace=True), preact: bool = False,
                 first: bool = False, name: str = "", instance_norm: bool = False):
        super().__init__()
        for i in range(len(args) - 1):
            plain = (not first) or (not preact) or (i != 0)
            self.add_module(
                name + "layer{}".format(i),
                Conv2d(args[i], args[i + 1], bn=plain and bn, activation=activation if plain else None,
                       preact=preact, instance_norm=instance_norm))
