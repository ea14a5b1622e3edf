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
