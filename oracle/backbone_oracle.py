"""CPU restatement of RaTrack's `Track4D.backbone` (PointNet++ heads + cost volume + flow decoder).

TEST INFRASTRUCTURE ONLY: used by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs as the checker / CPU baseline.  The product
(ratrack_b200/) never imports it.

Functional, driven by a reference-format state_dict (key names of SURVEY.md App. C):
fp32 torch-CPU ops for the dense layers (conv1x1 = matmul, BatchNorm, Linear, GRU) and
the C oracle (oracle/pointnet2_oracle.c) for the native pointnet2 ops.  Each function
cites the reference lines it restates.  Pinned against the UNMODIFIED reference Python
(run through oracle/ref_harness.py) by tests/test_oracle_golden.py::test_oracle_vs_live_reference_python in the dev
container and by tests/golden/backbone_*.npz (outputs of that reference run) everywhere.

BatchNorm: `training=False` uses running stats (net.eval()); `training=True` uses batch
statistics and returns nothing extra (running-stat updates are not modelled here).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pointnet2_oracle as P


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _bn(sd, prefix, x, training):
    """nn.BatchNorm2d defaults (eps 1e-5); lib/pytorch_utils.py:117-123."""
    return F.batch_norm(x, sd[prefix + "running_mean"].clone(), sd[prefix + "running_var"].clone(),
                        sd[prefix + "weight"], sd[prefix + "bias"], training=training, momentum=0.0, eps=1e-5)


def shared_mlp(sd, prefix, x, nlayers, training):
    """[conv1x1(no bias) -> BN2d -> ReLU] x L on (B,C,P,S).  lib/pytorch_utils.py:5-32,163-197."""
    for j in range(nlayers):
        x = F.conv2d(x, sd[f"{prefix}layer{j}.conv.weight"])
        x = _bn(sd, f"{prefix}layer{j}.bn.bn.", x, training)
        x = F.relu(x)
    return x


def query_and_group(radius, nsample, xyz, new_xyz, features):
    """lib/pointnet2_utils.py:269-292: ball_query -> group xyz - centre -> group features -> cat [xyz, feat]."""
    idx = P.ball_query(radius, nsample, xyz.numpy(), new_xyz.numpy())
    xyz_trans = xyz.transpose(1, 2).contiguous()
    grouped_xyz = _t(P.grouping_operation(xyz_trans.numpy(), idx))
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if features is None:
        return grouped_xyz
    grouped_features = _t(P.grouping_operation(features.contiguous().numpy(), idx))
    return torch.cat([grouped_xyz, grouped_features], dim=1)


def sa_module_msg(sd, prefix, xyz, features, npoint, radii, nsamples, nlayers, training):
    """lib/pointnet2_modules.py:19-55 (FPS -> gather -> per scale group+MLP+max -> concat)."""
    fps_idx = P.furthest_point_sample(xyz.numpy(), npoint)
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    new_xyz = _t(P.gather_operation(xyz_flipped.numpy(), fps_idx)).transpose(1, 2).contiguous()
    outs = []
    for i, (r, ns) in enumerate(zip(radii, nsamples)):
        g = query_and_group(r, ns, xyz, new_xyz, features)
        g = shared_mlp(sd, f"{prefix}mlps.{i}.", g, nlayers[i], training)
        outs.append(g.max(dim=3)[0])
    return new_xyz, torch.cat(outs, dim=1)


def fp_module(sd, prefix, unknown, known, unknow_feats, known_feats, training):
    """lib/pointnet2_modules.py:129-158 (three_nn -> inverse-distance weights -> interpolate -> cat -> 1-layer MLP)."""
    dist, idx = P.three_nn(unknown.numpy(), known.numpy())
    dist = _t(dist)
    dist_recip = 1.0 / (dist + 1e-8)
    norm = torch.sum(dist_recip, dim=2, keepdim=True)
    weight = dist_recip / norm
    interp = _t(P.three_interpolate(known_feats.contiguous().numpy(), idx, weight.numpy()))
    new = interp if unknow_feats is None else torch.cat([interp, unknow_feats], dim=1)
    new = shared_mlp(sd, f"{prefix}mlp.", new.unsqueeze(-1), 1, training)
    return new.squeeze(-1)


def pnhead(sd, prefix, pc, features, npoint, training=False):
    """utils/model_utils/model_utils.py:393-424.  pc (B,N,3), features (B,C,N) -> (l3_xyz, (B,128,N))."""
    l0_points = features.contiguous()
    l0_xyz = pc.contiguous()

    def lin(name, x):
        return F.linear(x.permute(0, 2, 1), sd[f"{prefix}{name}.weight"], sd[f"{prefix}{name}.bias"]).permute(0, 2, 1).contiguous()

    l1_xyz, l1_points = sa_module_msg(sd, f"{prefix}sa1.", l0_xyz, l0_points, npoint, [2, 4], [4, 8], [3, 3], training)
    l1_points = lin("linear1", l1_points)
    l2_xyz, l2_points = sa_module_msg(sd, f"{prefix}sa2.", l1_xyz, l1_points, npoint, [4, 8], [8, 16], [2, 2], training)
    l2_points = lin("linear2", l2_points)
    l3_xyz, l3_points = sa_module_msg(sd, f"{prefix}sa3.", l2_xyz, l2_points, npoint, [8, 16], [16, 32], [2, 2], training)
    l3_points = lin("linear3", l3_points)
    l2_points = fp_module(sd, f"{prefix}fp3.", l2_xyz, l3_xyz, l2_points, l3_points, training)
    l1_points = fp_module(sd, f"{prefix}fp2.", l1_xyz, l2_xyz, l1_points, l2_points, training)
    l0_points = fp_module(sd, f"{prefix}fp1.", l0_xyz, l1_xyz, None, l1_points, training)
    return l3_xyz, l0_points


def square_distance(src, dst):
    """utils/model_utils/model_utils.py:17-39 (expanded form, clamped at 0)."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return torch.maximum(dist, torch.zeros_like(dist))


def knn_point(nsample, xyz, new_xyz):
    """utils/model_utils/model_utils.py:85-99."""
    return torch.topk(square_distance(new_xyz, xyz), nsample, dim=-1, largest=False, sorted=False)[1]


def index_points(points, idx):
    """utils/model_utils/model_utils.py:42-59.  points (B,N,C), idx (B,S,K) -> (B,S,K,C)."""
    B = points.shape[0]
    bi = torch.arange(B).view(B, *([1] * (idx.dim() - 1))).expand_as(idx)
    return points[bi, idx, :]


def weightnet(sd, prefix, x):
    """utils/model_utils/model_utils.py:379-390 (bn=False: conv+bias -> ReLU, three times)."""
    for i in range(3):
        x = F.relu(F.conv2d(x, sd[f"{prefix}mlp_convs.{i}.weight"], sd[f"{prefix}mlp_convs.{i}.bias"]))
    return x


def feature_correlator(sd, prefix, pc1, pc2, feature1, feature2, nsample=16, knn_override=None):
    """utils/model_utils/model_utils.py:193-250.  pc (B,3,N), feature (B,D,N) -> (B,256,N).

    knn_override: optional (idx12, idx11) int64 tensors to substitute for the two knn_point
    results (tie-aware parity checks feed both sides the same neighbour sets).
    """
    B, C, N1 = pc1.shape
    pc1 = pc1.permute(0, 2, 1)
    pc2 = pc2.permute(0, 2, 1)
    feature1 = feature1.permute(0, 2, 1)
    feature2 = feature2.permute(0, 2, 1)
    D1 = feature1.shape[2]

    knn_idx = knn_point(nsample, pc2, pc1) if knn_override is None else knn_override[0]
    neighbor_xyz = index_points(pc2, knn_idx)
    direction_xyz = neighbor_xyz - pc1.reshape(B, N1, 1, C)
    grouped_feature2 = index_points(feature2, knn_idx)
    grouped_feature1 = feature1.reshape(B, N1, 1, D1).repeat(1, 1, nsample, 1)
    new_features = torch.cat([grouped_feature1, grouped_feature2, direction_xyz], dim=-1).permute(0, 3, 2, 1)
    for i in range(3):
        new_features = F.leaky_relu(
            F.conv2d(new_features, sd[f"{prefix}mlp_convs.{i}.weight"], sd[f"{prefix}mlp_convs.{i}.bias"]), 0.1)
    weights = weightnet(sd, f"{prefix}weightnet1.", direction_xyz.permute(0, 3, 2, 1))
    new_features = torch.sum(weights * new_features, dim=2)  # B C N

    knn_idx = knn_point(nsample, pc1, pc1) if knn_override is None else knn_override[1]
    neighbor_xyz = index_points(pc1, knn_idx)
    direction_xyz = neighbor_xyz - pc1.reshape(B, N1, 1, C)
    weights = weightnet(sd, f"{prefix}weightnet2.", direction_xyz.permute(0, 3, 2, 1))
    new_features = index_points(new_features.permute(0, 2, 1), knn_idx)
    new_features = weights * new_features.permute(0, 3, 2, 1)
    return torch.sum(new_features, dim=2)


def _predictor(sd, prefix, feat, training):
    """FlowPredictor / ClsPredictor trunk.  utils/model_utils/model_utils.py:321-329, 347-354."""
    feat = feat.unsqueeze(3)
    for i in range(3):
        feat = F.conv2d(feat, sd[f"{prefix}sf_mlp.{i}.0.weight"])
        feat = F.relu(_bn(sd, f"{prefix}sf_mlp.{i}.1.", feat, training))
    return F.conv2d(feat, sd[f"{prefix}conv2.weight"]).squeeze(3)


def flow_predictor(sd, prefix, feat, training=False):
    return _predictor(sd, prefix, feat, training)


def cls_predictor(sd, prefix, feat, training=False):
    """utils/model_utils/model_utils.py:347-357."""
    out = _predictor(sd, prefix, feat, training)
    out = F.linear(out.permute(0, 2, 1), sd[f"{prefix}linear.weight"], sd[f"{prefix}linear.bias"])
    return torch.sigmoid(out).squeeze(2)


def gru_step(sd, prefix, x, h, num_layers=5):
    """nn.GRU(128,128,5), seq_len 1: x (B,128), h (L,B,128).  model_utils.py:279,294-297 (torch GRU equations)."""
    h_out = []
    inp = x
    for l in range(num_layers):
        gi = F.linear(inp, sd[f"{prefix}weight_ih_l{l}"], sd[f"{prefix}bias_ih_l{l}"])
        gh = F.linear(h[l], sd[f"{prefix}weight_hh_l{l}"], sd[f"{prefix}bias_hh_l{l}"])
        i_r, i_z, i_n = gi.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        r = torch.sigmoid(i_r + h_r)
        z = torch.sigmoid(i_z + h_z)
        n = torch.tanh(i_n + r * h_n)
        hn = (1 - z) * n + z * h[l]
        h_out.append(hn)
        inp = hn
    return inp, torch.stack(h_out, 0)


def flow_decoder(sd, prefix, pc1, feature1, pc1_features, cor_features, h, npoint, training=False):
    """utils/model_utils/model_utils.py:281-305."""
    cls = cls_predictor(sd, f"{prefix}cp.", cor_features, training)
    embeddings = torch.cat((feature1, pc1_features, cor_features), dim=1)
    _, prop_features = pnhead(sd, f"{prefix}mse.", pc1.permute(0, 2, 1).contiguous(), embeddings, npoint, training)
    gfeat = torch.max(prop_features, -1)[0]  # (B,128)
    if h is None:
        h = torch.zeros(5, pc1.shape[0], 128)
    g, h = gru_step(sd, f"{prefix}torchGRU.", gfeat, h)
    g = g.unsqueeze(2).expand(prop_features.size(0), prop_features.size(1), pc1.size(2))
    new_features = torch.cat((prop_features, g), dim=1)
    output = flow_predictor(sd, f"{prefix}fp.", new_features, training)
    return output, h, prop_features, cls


def backbone(sd, pc1, pc2, feature1, feature2, h, npoint=512, training=False, knn_override=None):
    """models/track4d.py:67-106.  Returns the reference's 7-tuple
    (output, h, cls, cor_features, pc1_features, pc2_features, prop_features)."""
    sd = {k: v.detach().float().cpu() for k, v in sd.items()}
    with torch.no_grad():
        _, f1 = pnhead(sd, "pn_head.", pc1.permute(0, 2, 1).contiguous(), feature1, npoint, training)
        _, f2 = pnhead(sd, "pn_head.", pc2.permute(0, 2, 1).contiguous(), feature2, npoint, training)
        g1 = torch.max(f1, -1)[0].unsqueeze(2).expand(-1, -1, pc1.size(2))
        g2 = torch.max(f2, -1)[0].unsqueeze(2).expand(-1, -1, pc2.size(2))
        pc1_features = torch.cat((f1, g1), dim=1)
        pc2_features = torch.cat((f2, g2), dim=1)
        cor = feature_correlator(sd, "fc_layer.", pc1, pc2, pc1_features, pc2_features, 16, knn_override)
        output, h, prop, cls = flow_decoder(sd, "fd_layer.", pc1, feature1, pc1_features, cor, h, npoint, training)
    return output, h, cls, cor, pc1_features, pc2_features, prop
