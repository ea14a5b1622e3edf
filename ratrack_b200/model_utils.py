"""Seam C: the model blocks of RaTrack's hot path with the reference's names, constructor arguments,
forward signatures, return shapes and state_dict keys
(reference: src/utils/model_utils/model_utils.py:17-424 and src/models/track4d.py:13-106).

Two execution paths share one set of parameters:
  * modular path (this file): nn.Modules whose native ops are the sm_100a kernels behind
    ratrack_b200.lib.pointnet2_utils -- supports train and eval, autograd, any batch size;
  * fused path (ratrack_b200/engine.py): `Track4DBackbone.backbone` in eval mode without grad is
    served by the fused CUDA engine (BN folded, grouping+MLP+max fused, cost volume fused).

Differences from the reference that do not change results: no hard-coded `.cuda()` (tensors are
created on the inputs' device), `h=None` creates zeros for the actual batch size
(the reference hard-codes batch 1: model_utils.py:295), and the dead sub-module
`FlowDecoder.pnnGru` (constructed but never called, model_utils.py:278) is not instantiated --
reference checkpoints load with strict=False exactly as the reference itself loads them
(src/models/model.py:24,37).
"""
import contextlib

import torch
import torch.nn as nn
import torch.nn.functional as F

from .lib import dense_tc, pointnet2_utils
from .lib.pointnet2_modules import PointnetFPModule, PointnetSAModuleMSG
from .lib.pytorch_utils import PointwiseConv2d


@contextlib.contextmanager
def reference_dataflow():
    """Inside this context the modular path evaluates the network op by op exactly as the reference's own modules do --
    1x1 convolutions through nn.Conv2d / cuDNN, Linear through torch, QueryAndGroup and the
    cost volume as chains of channel-major ops -- with only the ten pointnet2 kernels native.  Used where the REFERENCE's
    arithmetic is the thing measured (bench.py `ref_gpu`, tests/test_gpu_parity_floor.py), never on the product path."""
    saved = (PointwiseConv2d.use_gemm, dense_tc.enabled, pointnet2_utils.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows)
    PointwiseConv2d.use_gemm = dense_tc.enabled = False
    pointnet2_utils.QueryAndGroup.rows_layout = FeatureCorrelator.fused_rows = False
    try:
        yield
    finally:
        PointwiseConv2d.use_gemm, dense_tc.enabled, pointnet2_utils.QueryAndGroup.rows_layout, FeatureCorrelator.fused_rows = saved


def square_distance(src, dst):
    """Expanded-form squared distances, clamped at 0: (B,N,C),(B,M,C) -> (B,N,M).  reference :17-39"""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return torch.clamp_min(dist, 0.0)


def index_points(points, idx):
    """points (B,N,C), idx (B,S[,K]) -> (B,S[,K],C).  reference :42-59"""
    B = points.shape[0]
    batch = torch.arange(B, dtype=torch.long, device=points.device).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[batch, idx, :]


def knn_point(nsample, xyz, new_xyz):
    """(B,S,nsample) int64 indices of the nsample nearest points of xyz (B,N,3) to each new_xyz (B,S,3).
    reference :85-99 (square_distance + torch.topk).  Runs the sm_100a kernel rt_knn_expanded: same expanded-form
    fp32 distances in torch's rounding order, never materialising the (B,S,N) matrix; ties -> lower index
    (torch.topk leaves them undefined)."""
    from . import _cabi

    q = new_xyz.contiguous().float()
    s = xyz.contiguous().float()
    B, S, _ = q.shape
    if nsample > 32 or not q.is_cuda:
        raise _cabi.RatrackError("knn_point: needs CUDA tensors and nsample <= 32 (no CPU / torch fallback)")
    idx = torch.empty(B, S, nsample, dtype=torch.int32, device=q.device)
    with torch.cuda.device_of(q):
        _cabi.call("rt_knn_expanded", B, S, s.shape[1], nsample, q.data_ptr(), s.data_ptr(), idx.data_ptr(),
                   torch.cuda.current_stream(q.device).cuda_stream)
    return idx.long()


class WeightNet(nn.Module):
    """conv 3->8->8->out + ReLU each; the BN modules exist (state_dict) but are unused when bn=False.  reference :359-390"""

    def __init__(self, in_channel, out_channel, hidden_unit=[8, 8], bn=False):
        super().__init__()
        self.bn = bn
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        widths = [in_channel] + list(hidden_unit or []) + [out_channel]
        for cin, cout in zip(widths[:-1], widths[1:]):
            self.mlp_convs.append(PointwiseConv2d(cin, cout, 1))
            self.mlp_bns.append(nn.BatchNorm2d(cout))

    def forward(self, localized_xyz):
        weights = localized_xyz
        for i, conv in enumerate(self.mlp_convs):
            weights = conv(weights)
            if self.bn:
                weights = self.mlp_bns[i](weights)
            weights = F.relu(weights)
        return weights


class FeatureCorrelator(nn.Module):
    """Cost volume: point-to-patch (kNN in pc2) then patch-to-patch (kNN in pc1).  reference :166-250"""

    def __init__(self, nsample, in_channel, mlp, bn=False, use_leaky=True):
        super().__init__()
        self.nsample = nsample
        self.bn = bn
        self.mlp_convs = nn.ModuleList()
        self.mlp2 = nn.ModuleList()
        if bn:
            self.mlp_bns = nn.ModuleList()
            self.mlp_bns2 = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(PointwiseConv2d(last_channel, out_channel, 1))
            if bn:
                self.mlp_bns.append(nn.BatchNorm2d(out_channel))
            last_channel = out_channel
        self.cls_mlp = nn.Linear(16, 1)  # unused by forward; kept for checkpoint compatibility
        self.weightnet1 = WeightNet(3, last_channel)
        self.weightnet2 = WeightNet(3, last_channel)
        self.relu = nn.ReLU(inplace=True) if not use_leaky else nn.LeakyReLU(0.1, inplace=True)
        self.sig = nn.Sigmoid()

    def forward(self, pc1, pc2, feature1, feature2):
        """pc (B,3,N), feature (B,D,N) -> (B,mlp[-1],N1).  reference :193-250.

        Same function, different evaluation: the reference concatenates [f1 (repeated 16x) ; f2[knn] ; dxyz] into a
        (B,515,16,N) tensor and runs the first 1x1 convolution over it.  That convolution is linear in the three blocks,
        W.[f1; f2_j; d_j] = Wa.f1 + Wb.f2_j + Wc.d_j, so Wa.f1 and Wb.f2 are evaluated once per POINT and gathered: the
        515-channel tensor never exists (8.6 GB at batch 256) and the layer costs 1/16 of the FLOPs, forward and backward.
        Every gather is the package's own CUDA op (deterministic segmented-sum backward), nothing goes through
        torch's atomics-based index_put."""
        B, C, N1 = pc1.shape
        K = self.nsample
        if self.bn:
            return self._forward_dense(pc1, pc2, feature1, feature2)
        if self._rows_path_ok(pc1, feature1):
            return self._forward_rows(pc1, pc2, feature1, feature2)
        xyz1, xyz2 = pc1.permute(0, 2, 1), pc2.permute(0, 2, 1)
        idx = knn_point(K, xyz2, xyz1).int().contiguous()                                  # (B,N1,K)
        direction = pointnet2_utils.grouping_operation(pc2.contiguous(), idx) - pc1.unsqueeze(-1)      # (B,3,N1,K)
        conv0 = self.mlp_convs[0]
        d1, d2 = feature1.shape[1], feature2.shape[1]
        w = conv0.weight.view(conv0.out_channels, conv0.in_channels)
        p1 = torch.matmul(w[:, :d1], feature1)                                             # (B,Cout,N1)
        p2 = torch.matmul(w[:, d1:d1 + d2], feature2).contiguous()                         # (B,Cout,N2)
        x = pointnet2_utils.grouping_operation(p2, idx)                                    # (B,Cout,N1,K)
        x = x + p1.unsqueeze(-1) + torch.einsum("oc,bcnk->bonk", w[:, d1 + d2:], direction)
        if conv0.bias is not None:
            x = x + conv0.bias.view(1, -1, 1, 1)
        x = self.relu(x)
        for conv in self.mlp_convs[1:]:
            x = self.relu(conv(x))
        weights = self.weightnet1(direction)
        x = torch.sum(weights * x, dim=3)                                                  # (B,C,N1)

        idx = knn_point(K, xyz1, xyz1).int().contiguous()
        direction = pointnet2_utils.grouping_operation(pc1.contiguous(), idx) - pc1.unsqueeze(-1)
        weights = self.weightnet2(direction)
        x = pointnet2_utils.grouping_operation(x.contiguous(), idx)
        return torch.sum(weights * x, dim=3).contiguous()

    fused_rows = True   # class-wide switch: False = the op-by-op chain above (A/B runs, parity tests)

    def _rows_path_ok(self, pc1, feature1):
        wn = self.weightnet1
        C = self.mlp_convs[-1].out_channels
        return (FeatureCorrelator.fused_rows and pc1.is_cuda and feature1.dtype == torch.float32 and self.nsample <= 32
                and not wn.bn and len(wn.mlp_convs) == 3 and wn.mlp_convs[-1].in_channels == 8 and C % 32 == 0 and C <= 1024
                and all(c.out_channels == C for c in self.mlp_convs) and isinstance(self.relu, nn.LeakyReLU)
                and abs(self.relu.negative_slope - 0.1) < 1e-12)

    @staticmethod
    def _wn_hidden(wn, direction_rows):
        """WeightNet's first two layers on (B,N,K,3) direction rows -> (B,N,K,8) hidden rows (its last layer is evaluated
        inside the weighted-sum kernel)."""
        h = direction_rows
        for conv in wn.mlp_convs[:-1]:
            h = F.relu(F.linear(h, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias))
        return h

    def _forward_rows(self, pc1, pc2, feature1, feature2):
        """The same function on the fused channels-innermost kernels (lib/costvol_train.py, csrc/costvol_train.cu): the
        (B,C,N1,K) tensors of the chain below exist only as the three (B,N1,K,C) activations the backward pass needs."""
        from .lib import costvol_train as cvt

        K = self.nsample
        xyz1, xyz2 = pc1.permute(0, 2, 1).contiguous(), pc2.permute(0, 2, 1).contiguous()
        idx12 = knn_point(K, xyz2, xyz1).int().contiguous()
        conv0 = self.mlp_convs[0]
        d1, d2 = feature1.shape[1], feature2.shape[1]
        w = conv0.weight.view(conv0.out_channels, conv0.in_channels)
        p1 = dense_tc.linear(feature1.permute(0, 2, 1).contiguous(), w[:, :d1])               # (B,N1,C) rows
        p2 = dense_tc.linear(feature2.permute(0, 2, 1).contiguous(), w[:, d1:d1 + d2])
        x, direction = cvt.cv_layer1(p1, p2, xyz1, xyz2, idx12, w[:, d1 + d2:], conv0.bias)   # (B,N1,K,C), (B,N1,K,3)
        for conv in self.mlp_convs[1:]:
            x = cvt.linear_act(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias, "leaky")
        wn = self.weightnet1
        last = wn.mlp_convs[-1]
        x = cvt.weighted_sum(x, self._wn_hidden(wn, direction), last.weight.view(last.out_channels, 8), last.bias)   # (B,N1,C)

        idx11 = knn_point(K, xyz1, xyz1).int().contiguous()
        direction = index_points(xyz1, idx11.long()) - xyz1.unsqueeze(2)                      # (B,N1,K,3)
        wn = self.weightnet2
        last = wn.mlp_convs[-1]
        x = cvt.weighted_sum(x, self._wn_hidden(wn, direction), last.weight.view(last.out_channels, 8), last.bias, idx=idx11)
        return x.permute(0, 2, 1).contiguous()

    def _forward_dense(self, pc1, pc2, feature1, feature2):
        """The reference's literal dataflow (needed when the first layer is followed by a BatchNorm: bn=True)."""
        B, C, N1 = pc1.shape
        pc1 = pc1.permute(0, 2, 1)
        pc2 = pc2.permute(0, 2, 1)
        feature1 = feature1.permute(0, 2, 1)
        feature2 = feature2.permute(0, 2, 1)
        D1 = feature1.shape[2]

        knn_idx = knn_point(self.nsample, pc2, pc1)
        direction_xyz = index_points(pc2, knn_idx) - pc1.reshape(B, N1, 1, C)
        grouped_feature2 = index_points(feature2, knn_idx)
        grouped_feature1 = feature1.reshape(B, N1, 1, D1).expand(-1, -1, self.nsample, -1)
        x = torch.cat([grouped_feature1, grouped_feature2, direction_xyz], dim=-1).permute(0, 3, 2, 1)
        for i, conv in enumerate(self.mlp_convs):
            x = conv(x)
            if self.bn:
                x = self.mlp_bns[i](x)
            x = self.relu(x)
        weights = self.weightnet1(direction_xyz.permute(0, 3, 2, 1))
        x = torch.sum(weights * x, dim=2)  # (B,C,N)

        knn_idx = knn_point(self.nsample, pc1, pc1)
        direction_xyz = index_points(pc1, knn_idx) - pc1.reshape(B, N1, 1, C)
        weights = self.weightnet2(direction_xyz.permute(0, 3, 2, 1))
        x = index_points(x.permute(0, 2, 1), knn_idx)
        return torch.sum(weights * x.permute(0, 3, 2, 1), dim=2).contiguous()


class FlowPredictor(nn.Module):
    """[conv1x1(no bias)+BN+ReLU] x3 -> conv1x1 -> (B,3,N).  reference :308-329"""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.sf_mlp = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.sf_mlp.append(nn.Sequential(PointwiseConv2d(last_channel, out_channel, 1, bias=False),
                                             nn.BatchNorm2d(out_channel), nn.ReLU(inplace=False)))
            last_channel = out_channel
        self.conv2 = PointwiseConv2d(mlp[-1], 3, 1, bias=False)

    def forward(self, feat):
        feat = feat.unsqueeze(3)
        for block in self.sf_mlp:
            feat = block(feat)
        return self.conv2(feat).squeeze(3).contiguous()


class ClsPredictor(nn.Module):
    """FlowPredictor trunk -> Linear(3->1) -> sigmoid -> (B,N).  reference :332-357"""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.sf_mlp = nn.ModuleList()
        last_channel = in_channel
        for out_channel in mlp:
            self.sf_mlp.append(nn.Sequential(PointwiseConv2d(last_channel, out_channel, 1, bias=False),
                                             nn.BatchNorm2d(out_channel), nn.ReLU(inplace=False)))
            last_channel = out_channel
        self.conv2 = PointwiseConv2d(mlp[-1], 3, 1, bias=False)
        self.linear = nn.Linear(3, 1)
        self.sig = nn.Sigmoid()

    def forward(self, feat):
        feat = feat.unsqueeze(3)
        for block in self.sf_mlp:
            feat = block(feat)
        out = self.conv2(feat)
        out = self.linear(out.squeeze(3).permute(0, 2, 1))
        return self.sig(out).squeeze(2)


class PNHead(nn.Module):
    """3 x SA-MSG + 3 x Linear + 3 x FP.  pc (B,N,3), features (B,C,N) -> (l3_xyz (B,S,3), (B,128,N)).  reference :393-424"""

    def __init__(self, sample_point_num, in_channels):
        super().__init__()
        S = sample_point_num
        self.sa1 = PointnetSAModuleMSG(npoint=S, radii=[2, 4], nsamples=[4, 8],
                                       mlps=[[in_channels, 16, 16, 32], [in_channels, 16, 16, 32]])
        self.sa2 = PointnetSAModuleMSG(npoint=S, radii=[4, 8], nsamples=[8, 16],
                                       mlps=[[3 + 32, 32, 32], [3 + 32, 32, 64]])
        self.sa3 = PointnetSAModuleMSG(npoint=S, radii=[8, 16], nsamples=[16, 32],
                                       mlps=[[3 + 64, 64, 64], [3 + 64, 64, 64]])
        self.fp3 = PointnetFPModule(mlp=[128, 128])
        self.fp2 = PointnetFPModule(mlp=[160, 128])
        self.fp1 = PointnetFPModule(mlp=[128, 128])
        self.linear1 = nn.Linear(64, 32)
        self.linear2 = nn.Linear(96, 64)
        self.linear3 = nn.Linear(128, 64)

    @staticmethod
    def _lin(layer, x):
        return dense_tc.linear(x.permute(0, 2, 1).contiguous(), layer.weight, layer.bias).permute(0, 2, 1).contiguous()

    def forward(self, pc, features):
        l0_points = features.contiguous()
        l0_xyz = pc.contiguous()
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points)
        l1_points = self._lin(self.linear1, l1_points)
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points)
        l2_points = self._lin(self.linear2, l2_points)
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l3_points = self._lin(self.linear3, l3_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(l0_xyz, l1_xyz, None, l1_points)
        return l3_xyz, l0_points


class FlowDecoder(nn.Module):
    """cls head, PNHead over [ft, feat, cost] embeddings, 5-layer GRU on the global feature, flow head.  reference :253-305"""

    def __init__(self, fc_inch, args):
        super().__init__()
        ep_inch = fc_inch * 2 + 5
        ep_mlp2s = [int(fc_inch / 8)] * 3
        num_eps = 4
        self.mse = PNHead(args.npoints, ep_inch)
        sf_inch = num_eps * ep_mlp2s[-1] * 2
        sf_mlps = [int(sf_inch / 2), int(sf_inch / 4), int(sf_inch / 8)]
        self.fp = FlowPredictor(in_channel=sf_inch, mlp=sf_mlps)
        self.cp = ClsPredictor(in_channel=sf_inch, mlp=sf_mlps)
        self.mlp2 = nn.ModuleList([nn.Linear(3, 1)])                    # unused by forward (checkpoint compat)
        self.gru2 = nn.GRU(input_size=fc_inch, hidden_size=fc_inch)     # unused by forward (checkpoint compat)
        self.sig = nn.Sigmoid()
        self.torchGRU = nn.GRU(fc_inch // 2, fc_inch // 2, 5)

    def forward(self, pc1, feature1, pc1_features, cor_features, h):
        cls = self.cp(cor_features)
        if feature1 is not None:
            embeddings = torch.cat((feature1, pc1_features, cor_features), dim=1)
        else:
            embeddings = torch.cat((pc1_features, cor_features), dim=1)
        _, prop_features = self.mse(pc1.permute(0, 2, 1).contiguous(), embeddings)
        gfeat = torch.max(prop_features, -1)[0].unsqueeze(2)
        if h is None:
            h = torch.zeros(5, pc1.size(0), 128, device=pc1.device, dtype=pc1.dtype)
        gfeat, h = self.torchGRU(gfeat.permute(2, 0, 1), h)
        gfeat = gfeat.permute(1, 2, 0).expand(prop_features.size(0), prop_features.size(1), pc1.size(2))
        output = self.fp(torch.cat((prop_features, gfeat), dim=1))
        return output, h, prop_features, cls


class Track4DBackbone(nn.Module):
    """The hot path of `Track4D` (reference: src/models/track4d.py:13-106): pn_head, fc_layer, fd_layer and
    the methods backbone / flow_head / feature_extraction_head with the reference's signatures and
    parameter names (`pn_head.*`, `fc_layer.*`, `fd_layer.*`), so a reference checkpoint loads with
    strict=False.  Clustering / association (track4d.py:108-223) are out of this path's scope.

    `args` needs `.npoints` (FPS sample count of all three SA levels, configs.yaml:25).
    """

    def __init__(self, args):
        super().__init__()
        self.npoints_cfg = args.npoints
        self.pn_head = PNHead(args.npoints, 5)
        fc_inch = 2 * 128
        self.fc_layer = FeatureCorrelator(16, in_channel=fc_inch * 2 + 3, mlp=[fc_inch, fc_inch, fc_inch])
        self.fd_layer = FlowDecoder(fc_inch=fc_inch, args=args)
        self.use_fused = True   # eval + no_grad -> fused engine (ratrack_b200/engine.py)
        self.capture_knn = False  # fused path: also return the cost volume's neighbour sets (tie-aware parity checks)
        self.checked_forward = False   # True: synchronise after every fused forward and re-run it on the fp32 SIMT kernels if the fp16-range guard fired
        self._engine = None

    # -- reference-shaped methods ------------------------------------------------------------
    def feature_extraction_head(self, feature1, feature2, pc1, pc2):
        xyz1_new, pc1_features = self.pn_head(pc1.permute(0, 2, 1).contiguous(), feature1)
        xyz2_new, pc2_features = self.pn_head(pc2.permute(0, 2, 1).contiguous(), feature2)
        return feature1, pc1, pc1_features, pc2, pc2_features, xyz1_new, xyz2_new

    def flow_head(self, feature1, h, pc1, pc1_features, pc2, pc2_features, xyz1_new, xyz2_new):
        gfeat_1 = torch.max(pc1_features, -1)[0].unsqueeze(2).expand(-1, -1, pc1.size(2))
        gfeat_2 = torch.max(pc2_features, -1)[0].unsqueeze(2).expand(-1, -1, pc2.size(2))
        pc1_features = torch.cat((pc1_features, gfeat_1), dim=1)
        pc2_features = torch.cat((pc2_features, gfeat_2), dim=1)
        cor_features = self.fc_layer(pc1, pc2, pc1_features, pc2_features)
        output, h, prop_features, cls = self.fd_layer(pc1, feature1, pc1_features, cor_features, h)
        return cor_features, h, output, cls, pc1_features, pc2_features, prop_features

    def backbone_modular(self, pc1, pc2, feature1, feature2, h):
        feature1, pc1, pc1_features, pc2, pc2_features, xyz1_new, xyz2_new = \
            self.feature_extraction_head(feature1, feature2, pc1, pc2)
        cor_features, h, output, cls, pc1_features, pc2_features, prop_features = \
            self.flow_head(feature1, h, pc1, pc1_features, pc2, pc2_features, xyz1_new, xyz2_new)
        return output, h, cls, cor_features, pc1_features, pc2_features, prop_features

    def backbone(self, pc1, pc2, feature1, feature2, h, npts1=None, npts2=None):
        """pc (B,3,N), feature (B,2,N), h (5,B,128)|None ->
        (output (B,3,N), h, cls (B,N), cor_features (B,256,N), pc1_features, pc2_features (B,256,N), prop_features (B,128,N))
        npts1 / npts2 (optional, beyond the reference's signature): (B,) point counts of a zero-padded variable-size batch
        (data_io.PaddedBatcher) -- served by the fused engine only."""
        if npts1 is not None:
            if not (self.use_fused and not self.training and not torch.is_grad_enabled()):
                raise RuntimeError("variable-size batches (npts1 / npts2) are served by the fused inference engine only")
            eng = self._fused_engine()
            run = eng.run_checked if self.checked_forward else eng
            return run(pc1, pc2, feature1, feature2, h, want_knn=self.capture_knn, npts1=npts1, npts2=npts2)
        if self.use_fused and not self.training and not torch.is_grad_enabled():
            # fp16-range guard of the tensor-core kernels: a forward that trips it returns NaN in flow / cls / h (the device
            # overwrites them), reports through a status word polled without synchronisation at the next call, and the
            # engine then stays on its fp32 SIMT kernels.  `checked_forward = True` trades one stream synchronisation per
            # call for an immediate re-run of such a forward (what infer_host does).
            eng = self._fused_engine()
            run = eng.run_checked if self.checked_forward else eng
            return run(pc1, pc2, feature1, feature2, h, want_knn=self.capture_knn)
        return self.backbone_modular(pc1, pc2, feature1, feature2, h)

    def _fused_engine(self):
        if self._engine is None:
            from .engine import FusedBackbone
            self._engine = FusedBackbone(self)
        return self._engine

    # The engine snapshots BatchNorm-folded copies of the parameters: anything that can change a parameter or move the
    # module drops the snapshot (train()/eval(), load_state_dict, .to()/.cuda()/.half()...).  In-place edits of
    # parameters that bypass these (optimizer steps happen in train mode; manual `p.data.copy_`, EMA) need refresh_engine().
    def refresh_engine(self):
        self._engine = None

    def _apply(self, fn, *args, **kw):
        self._engine = None
        return super()._apply(fn, *args, **kw)

    def load_state_dict(self, *args, **kw):
        self._engine = None
        return super().load_state_dict(*args, **kw)

    def forward(self, pc1, pc2, feature1, feature2, h=None):
        return self.backbone(pc1, pc2, feature1, feature2, h)

    @torch.no_grad()
    def infer_host(self, pc1, pc2, feature1, feature2, h=None):
        """Host-buffer entry: (B,3,N)/(B,2,N) fp32 host tensors (pinned for async copies) -> (flow (B,3,N),
        cls (B,N)) as host tensors.  H2D copies, the backbone and the D2H result reads all happen here; one call is
        serial by construction -- `infer_host_stream` overlaps the copies of neighbouring batches with the compute."""
        dev = next(self.parameters()).device
        args = [x.to(dev, non_blocking=True) for x in (pc1, pc2, feature1, feature2)]
        if h is None:
            h = torch.zeros(5, pc1.size(0), 128, device=dev)
        fused = self.use_fused and not self.training
        out = (self._fused_engine().run_checked(args[0], args[1], args[2], args[3], h, want_knn=self.capture_knn) if fused
               else self.backbone(args[0], args[1], args[2], args[3], h))
        key = (tuple(out[0].shape), dev)
        if getattr(self, "_host_out_key", None) != key:
            self._host_out = (torch.empty(out[0].shape, dtype=torch.float32).pin_memory(),
                              torch.empty(out[2].shape, dtype=torch.float32).pin_memory())
            self._host_out_key = key
        self._host_out[0].copy_(out[0], non_blocking=True)
        self._host_out[1].copy_(out[2], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return self._host_out

    @torch.no_grad()
    def infer_host_stream(self, batches, depth=2):
        """Pipelined host-buffer entry for throughput: `batches` yields (pc1, pc2, feature1, feature2) pinned host tensors
        of one shape; this generator yields (flow (B,3,N), cls (B,N)) pinned host tensors in order, each valid until the
        second-next result is requested.  The H2D copy of batch i+1 (copy stream), the backbone of batch i
        (current stream) and the D2H read of batch i-1 (second copy stream) overlap; every batch still pays its own copies.
        No recurrent state is carried (h = 0 per batch, as in the reference's per-sequence reset, main_utils.py:94-98)."""
        from collections import deque

        dev = next(self.parameters()).device
        main = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        slots, pending, h = [], deque(), None
        fused = self.use_fused and not self.training

        def take(item):
            ev, slot, batch = item
            ev.synchronize()
            flow, cls = slot["host"]
            if fused and bool(torch.isnan(flow.view(-1)[0])):      # the fp16-range guard fired: that step was poisoned on the device
                out = self._fused_engine().run_checked(*[x.to(dev) for x in batch], h)
                flow.copy_(out[0])
                cls.copy_(out[2])
            return flow, cls

        for i, batch in enumerate(batches):
            if len(slots) < depth + 2:
                slots.append({"dev": [torch.empty(x.shape, dtype=torch.float32, device=dev) for x in batch], "host": None,
                              "free": None})
            slot = slots[i % (depth + 2)]
            if h is None:
                h = torch.zeros(5, batch[0].size(0), 128, device=dev)
            with torch.cuda.stream(s_in):
                if slot["free"] is not None:
                    s_in.wait_event(slot["free"])               # the forward that last read these device buffers is done
                for d, x in zip(slot["dev"], batch):
                    d.copy_(x, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            main.wait_event(ev_in)
            out = self.backbone(*slot["dev"], h)
            slot["free"] = torch.cuda.Event()
            slot["free"].record(main)
            if slot["host"] is None:
                slot["host"] = (torch.empty(out[0].shape, dtype=torch.float32).pin_memory(),
                                torch.empty(out[2].shape, dtype=torch.float32).pin_memory())
            with torch.cuda.stream(s_out):
                s_out.wait_event(slot["free"])
                slot["host"][0].copy_(out[0], non_blocking=True)
                slot["host"][1].copy_(out[2], non_blocking=True)
                out[0].record_stream(s_out)
                out[2].record_stream(s_out)
                ev_out = torch.cuda.Event()
                ev_out.record(s_out)
            pending.append((ev_out, slot, batch))
            if len(pending) > depth:
                yield take(pending.popleft())
        while pending:
            yield take(pending.popleft())

    @staticmethod
    def fused_available():
        try:
            from . import engine  # noqa: F401
            return True
        except ImportError:
            return False

    def cost_volume_neighbours(self, pc1, pc2):
        """The two 16-NN index sets the cost volume uses (pc1->pc2, pc1->pc1), (B,N,16) each --
        exposed so parity checks can be tie-aware (reference: model_utils.py:216,239)."""
        if self.use_fused and not self.training and self._engine is not None and self._engine.last_knn is not None:
            return tuple(k.long() for k in self._engine.last_knn)   # what the fused forward actually used
        a, b = pc1.permute(0, 2, 1), pc2.permute(0, 2, 1)
        return knn_point(self.fc_layer.nsample, b, a), knn_point(self.fc_layer.nsample, a, a)

    def train(self, mode: bool = True):
        self._engine = None  # parameters may change: re-fold on next eval call
        return super().train(mode)
