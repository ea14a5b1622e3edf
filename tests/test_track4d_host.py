"""Host-side pieces of the Track4D drop-in that need no GPU: the batched object embeddings against the reference's
per-object expressions (src/models/track4d.py:202-216), the Affinity MLP on embedding differences, the state_dict surface."""
import torch

from ratrack_b200.track4d import Affinity, Track4D, object_embeddings


class Args:
    npoints = 512
    min_obj_points = 2


def _reference_embedding(o):
    """one object (1,139,k) -> (1,1,141), literally the reference's lines"""
    feat = torch.max(o[:, 11:(11 + 128), :], dim=2)[0].unsqueeze(1)
    flow = torch.mean(o[:, 6:9, :], dim=2).unsqueeze(1)
    pos = torch.mean(o[:, 3:6, :], dim=2).unsqueeze(1)
    rrv = torch.mean(o[:, 9:11, :], dim=2).unsqueeze(1)
    rrv_var = torch.var(o[:, 9:11, :], dim=2, unbiased=False).unsqueeze(1)
    var = torch.var(o[:, 3:6, :], dim=2, unbiased=False).unsqueeze(1)
    return torch.cat((pos, var, feat, flow, rrv, rrv_var), dim=2)


def test_object_embeddings_match_per_object_reference_expressions():
    g = torch.Generator().manual_seed(3)
    sizes = [1, 2, 7, 40, 3]
    objs = [torch.randn(1, 139, k, generator=g) * 5 + torch.tensor([50.0]) for k in sizes]      # positions far from 0
    pts = torch.cat([o[0] for o in objs], dim=1)
    seg = torch.cat([torch.full((k,), i, dtype=torch.long) for i, k in enumerate(sizes)])
    emb = object_embeddings(pts, seg, len(sizes))
    ref = torch.cat([_reference_embedding(o)[0] for o in objs], dim=0)
    assert emb.shape == (5, 141)
    assert float((emb - ref).abs().max()) <= 2e-5 * float(ref.abs().max())


def test_batched_affinity_equals_pairwise_calls():
    torch.manual_seed(0)
    aff = Affinity(141)
    e_prev, e_curr = torch.randn(4, 141), torch.randn(3, 141)
    batched = aff.affinity((e_curr[None] - e_prev[:, None]).reshape(12, 141)).reshape(4, 3)
    for i in range(4):
        for j in range(3):
            one = aff(e_curr[j].view(1, 1, 141), e_prev[i].view(1, 1, 141))      # the reference's call shape (track4d.py:217)
            assert abs(float(one) - float(batched[i, j])) <= 1e-6


def test_track4d_state_dict_surface():
    net = Track4D(Args())
    keys = set(net.state_dict().keys())
    for k in ["bin_score", "affinity.affinity.0.weight", "affinity.affinity.8.bias", "pn_head.sa1.mlps.0.layer0.conv.weight",
              "fc_layer.mlp_convs.0.weight", "fd_layer.torchGRU.weight_hh_l4", "fd_layer.cp.linear.weight"]:
        assert k in keys, k
    assert net.state_dict()["affinity.affinity.0.weight"].shape == (564, 141)
    assert net.affinity_module([], dict())[1].shape == (1, 0)          # nothing to associate: the reference's empty matrix
