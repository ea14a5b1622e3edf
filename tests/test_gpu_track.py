"""The association path of Track4D (clustering -> affinity -> Sinkhorn -> ids) on the device,
against two consecutive frames of the UNMODIFIED reference `Track4D.forward` (tests/golden/track4d_two_frames.npz,
generator oracle/gen_golden_track.py).  The golden's backbone outputs are fed in, so every later quantity is comparable
one to one (a 1e-5 difference in `cls` would move points across the 0.5 motion threshold)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from ratrack_b200 import synthetic
from ratrack_b200.track4d import Track4D

pytestmark = pytest.mark.gpu


class Args:
    npoints = 512
    min_obj_points = 2


def test_track_two_frames_vs_reference_forward():
    g = np.load(os.path.join(GOLDEN, "track4d_two_frames.npz"))
    net = Track4D(Args())
    missing = net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    net = net.cuda().eval()
    d = synthetic.make_batch(2, 512, seed=1234)
    prev = dict()
    with torch.no_grad():
        for fr in range(2):
            pc1 = torch.from_numpy(d["pc1"][fr:fr + 1]).cuda()
            ft1 = torch.from_numpy(d["ft1"][fr:fr + 1]).cuda()
            out = (torch.from_numpy(g[f"f{fr}_flow"]).cuda(), torch.from_numpy(g[f"f{fr}_h"]).cuda(),
                   torch.from_numpy(g[f"f{fr}_cls"]).cuda(), None, None, None, torch.from_numpy(g[f"f{fr}_prop"]).cuda())
            h, warp, cls, aff_list, aff_mat, idx1, confs, objects, _, objs_curr = net.track(pc1, ft1, out, prev)
            # clusters: same number, same points, same order
            assert len(objs_curr) == int(g[f"f{fr}_nobj"])
            feats = torch.cat((warp, pc1, out[0], ft1, out[6]), dim=1)[0]
            for i, o in enumerate(objs_curr):
                want = feats[:, torch.from_numpy(g[f"f{fr}_obj{i}"]).cuda()]
                assert tuple(o.shape) == (1, 139, want.shape[1]) and torch.equal(o[0], want), (fr, i)
            # affinities, matches, ids
            ref_aff = g[f"f{fr}_aff_mat"]
            assert tuple(aff_mat.shape) == tuple(ref_aff.shape)
            if ref_aff.size:
                assert np.abs(aff_mat.cpu().numpy() - ref_aff).max() <= 1e-5
                assert np.array_equal(idx1.cpu().numpy(), g[f"f{fr}_idx1"])
                assert aff_list.shape == (ref_aff.shape[1] * ref_aff.shape[2],)
            else:
                assert idx1 is None and aff_list == []
            assert list(objects.keys()) == g[f"f{fr}_ids"].tolist()
            assert np.abs(np.array([float(c) for c in confs], np.float32) - g[f"f{fr}_confs"]).max() <= 1e-5
            prev = {k: v.clone().detach() for k, v in objects.items()}
    assert net.max_id == int(g["f1_ids"].max()) + 1


def test_track4d_forward_runs_end_to_end_and_keeps_reference_keys():
    net = Track4D(Args())
    sd = synthetic.make_state_dict(net, seed=1234)
    assert "affinity.affinity.8.weight" in sd and "bin_score" in sd and "fd_layer.torchGRU.weight_ih_l0" in sd
    net.load_state_dict(sd, strict=False)
    net = net.cuda().eval()
    d = synthetic.make_batch(1, 512, seed=5)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    with torch.no_grad():
        r = net(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None, dict())
        r2 = net(t["pc1"], t["pc2"], t["ft1"], t["ft2"], None, {k: v.clone() for k, v in r[7].items()})   # same h: same frame
    assert len(r) == 10 and tuple(r[1].shape) == (1, 3, 512) and tuple(r[0].shape) == (5, 1, 128)
    # the same frame again against its own objects: same clusters, a square affinity matrix, ids either kept or fresh
    n = len(r2[9])
    assert n == len(r[9]) and r2[4].shape == (1, n, n)
    assert all((k in r[7]) or k >= n for k in r2[7].keys())


def _load_net():
    net = Track4D(Args())
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    return net.cuda().eval()


def test_track4d_sequence_on_its_own_backbone_vs_reference_forward():
    """Ten consecutive frames of the UNMODIFIED reference forward with the recurrent state carried (h, previous objects,
    id counter): tests/golden/track4d_sequence.npz, generator oracle/gen_golden_track_seq.py.  Here the product runs on
    ITS OWN backbone outputs.  Hard decisions (cls > 0.5, DBSCAN eps) must equal the reference's; a frame may differ only
    where the reference itself sits within the floating-point tolerance of the threshold (the golden stores both margins),
    and from such a frame on the two runs legitimately follow different object sets, so the comparison stops there."""
    from oracle.gen_golden_track_seq import FRAMES, POINTS, SEED

    g = np.load(os.path.join(GOLDEN, "track4d_sequence.npz"))
    net = _load_net()
    d = synthetic.make_batch(FRAMES, POINTS, seed=SEED)
    prev, h, matched = dict(), None, 0
    with torch.no_grad():
        for fr in range(FRAMES):
            a = {k: torch.from_numpy(v[fr:fr + 1]).cuda() for k, v in d.items()}
            h, warp, cls, aff_list, aff_mat, idx1, confs, objects, _, objs_curr = net(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h, prev)
            cls_ref = g[f"f{fr}_cls"]
            assert np.abs(cls.cpu().numpy() - cls_ref).max() <= 1e-4 and np.abs(warp.cpu().numpy() - g[f"f{fr}_warp"]).max() <= 1e-3
            differ = (cls.cpu().numpy() > 0.5) != (cls_ref > 0.5)
            if differ.any():
                assert (np.abs(cls_ref - 0.5)[differ] <= 1e-4).all(), ("motion mask differs away from the threshold", fr)
                break
            # the same clusters = the same point sets in the same order: every object's (warped xyz, xyz) columns must be the
            # columns of the reference's point indices (exact duplicates of a point make index recovery ambiguous, columns not)
            feats = torch.cat((warp, a["pc1"]), dim=1)[0]
            same = len(objs_curr) == int(g[f"f{fr}_nobj"]) and all(
                o.shape[2] == len(g[f"f{fr}_obj{i}"]) and torch.equal(o[0, :6], feats[:, torch.from_numpy(g[f"f{fr}_obj{i}"]).cuda()])
                for i, o in enumerate(objs_curr))
            if not same:
                assert float(g[f"f{fr}_margin_eps"]) <= 2e-3, ("clusters differ although no pair distance is near eps", fr)
                break
            if list(objects.keys()) != g[f"f{fr}_ids"].tolist():
                # same clusters, different association: only legitimate when two affinities are nearly tied
                ref_aff = g[f"f{fr}_aff_mat"]
                assert ref_aff.size and np.abs(aff_mat.cpu().numpy() - ref_aff).max() <= 1e-3, fr
                break
            matched += 1
            prev = {k: v.clone().detach() for k, v in objects.items()}
    assert matched >= 4, matched      # frames 0-3 have comfortable margins (printed by the generator)
    if matched == FRAMES:
        assert net.max_id == int(g["max_id"])


def test_forward_batch_equals_per_sequence_forward():
    """B sequences per call (forward_batch) against the same frames pushed one at a time through the batch-1 forward."""
    d = synthetic.make_batch(6, 320, seed=21)
    t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
    B = 3
    single = _load_net()
    ref_frames = [[], [], []]
    with torch.no_grad():
        for b in range(B):
            single.max_id = 0
            prev, h = dict(), None
            for step in range(2):
                i = 2 * b + step
                r = single(t["pc1"][i:i + 1], t["pc2"][i:i + 1], t["ft1"][i:i + 1], t["ft2"][i:i + 1], h, prev)
                h, prev = r[0], {k: v.clone() for k, v in r[7].items()}
                ref_frames[b].append(r)
        net = _load_net()
        prevs, h, ids = [dict() for _ in range(B)], None, [0] * B
        for step in range(2):
            sel = [2 * b + step for b in range(B)]
            h, warp, cls, frames, ids = net.forward_batch(t["pc1"][sel], t["pc2"][sel], t["ft1"][sel], t["ft2"][sel], h, prevs, max_ids=ids)
            for b in range(B):
                r = ref_frames[b][step]
                assert torch.equal(cls[b:b + 1], r[2]) and torch.equal(warp[b:b + 1], r[1]) and torch.equal(h[:, b:b + 1], r[0])
                aff_list, aff_mat, idx1, confs, objects, _, objs_curr = frames[b]
                assert list(objects.keys()) == list(r[7].keys()) and len(objs_curr) == len(r[9])
                assert all(torch.equal(x, y) for x, y in zip(objs_curr, r[9]))
                assert tuple(aff_mat.shape) == tuple(r[4].shape) and (aff_mat.numel() == 0 or torch.allclose(aff_mat, r[4], atol=1e-6))
            prevs = [{k: v.clone() for k, v in frames[b][4].items()} for b in range(B)]
