"""Generate tests/golden/track4d_two_frames.npz from the UNMODIFIED reference `Track4D.forward` (dev container only):
`python -m oracle.gen_golden_track`.  TEST INFRASTRUCTURE ONLY.

Two consecutive frame pairs (synthetic.make_batch(2, 512, seed 1234), weights synthetic.make_state_dict seed 1234) through
the reference's forward (src/models/track4d.py:49-65: backbone, DBSCAN clustering with scikit-learn, affinity MLP in a
Python double loop, 500-iteration Sinkhorn, id bookkeeping), the second frame seeing the first frame's objects.  Stored per
frame: the backbone outputs the association path consumes (flow, cls, prop, h), the moving-point labels, every object as
its point indices, aff_mat, indices1, confidences and the object ids in dict order."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    from oracle import ref_harness
    from ratrack_b200 import synthetic

    net = ref_harness.make_track4d(npoints=512)
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=True)
    net.eval()
    d = synthetic.make_batch(2, 512, seed=1234)
    t = {k: torch.from_numpy(v) for k, v in d.items()}
    save = {}
    prev, h = dict(), None
    with torch.no_grad():
        for fr in range(2):
            a = {k: v[fr:fr + 1] for k, v in t.items()}
            out = net.backbone(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h if h is not None else torch.zeros(5, 1, 128))
            h_in = h
            h2, warp, cls, aff_list, aff_mat, idx1, confs, objects, _, objs_curr = net(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h_in, prev)
            mov = np.nonzero((cls > 0.5).squeeze(0).numpy())[0]
            feats = torch.cat((warp, a["pc1"], out[0], a["ft1"], out[6]), dim=1)[0].numpy()      # (139, N)
            save[f"f{fr}_flow"], save[f"f{fr}_cls"], save[f"f{fr}_prop"], save[f"f{fr}_h"] = out[0].numpy(), cls.numpy(), out[6].numpy(), h2.numpy()
            save[f"f{fr}_nobj"] = np.int64(len(objs_curr))
            for i, o in enumerate(objs_curr):      # recover every object's point indices from its xyz columns
                cols = o[0].numpy()
                ids = [int(np.nonzero((feats[:, mov] == cols[:, [c]]).all(0))[0][0]) for c in range(cols.shape[1])]
                save[f"f{fr}_obj{i}"] = mov[np.array(ids)]
            save[f"f{fr}_aff_mat"] = aff_mat.numpy()
            save[f"f{fr}_idx1"] = idx1.numpy() if idx1 is not None else np.zeros((1, 0), np.int64)
            save[f"f{fr}_confs"] = np.array([float(c) for c in confs], np.float32)
            save[f"f{fr}_ids"] = np.array(list(objects.keys()), np.int64)
            print("frame", fr, "moving", len(mov), "objects", len(objs_curr), "aff", tuple(aff_mat.shape), "ids", list(objects.keys()))
            prev = {k: v.clone().detach() for k, v in objects.items()}
            h = h2
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "track4d_two_frames.npz"), **save)
