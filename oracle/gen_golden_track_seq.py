"""Generate tests/golden/track4d_sequence.npz from the UNMODIFIED reference `Track4D.forward` (dev container only):
`python -m oracle.gen_golden_track_seq`.  TEST INFRASTRUCTURE ONLY.

Ten consecutive frame pairs (synthetic.make_batch(10, 384, seed 77), weights synthetic.make_state_dict seed 1234) through the
reference's forward (src/models/track4d.py:49-65) with the recurrent state (h, objects of the previous frame, max_id) carried
from frame to frame, exactly as main_utils.epoch does (:127-157).  Stored per frame: cls, the warped cloud, every object as
its point indices, indices1, the object ids in dict order, and two robustness margins: the smallest |cls - 0.5| and the
smallest gap of a pairwise DBSCAN feature distance from eps = 1.5 (how far the frame is from a different hard decision)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
FRAMES, POINTS, SEED = 10, 384, 77

if __name__ == "__main__":
    from oracle import ref_harness
    from ratrack_b200 import synthetic

    net = ref_harness.make_track4d(npoints=512)
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=True)
    net.eval()
    d = synthetic.make_batch(FRAMES, POINTS, seed=SEED)
    t = {k: torch.from_numpy(v) for k, v in d.items()}
    save = {}
    prev, h = dict(), None
    with torch.no_grad():
        for fr in range(FRAMES):
            a = {k: v[fr:fr + 1] for k, v in t.items()}
            out = net.backbone(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h if h is not None else torch.zeros(5, 1, 128))
            h2, warp, cls, aff_list, aff_mat, idx1, confs, objects, _, objs_curr = net(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h, prev)
            mov = np.nonzero((cls > 0.5).squeeze(0).numpy())[0]
            feats = torch.cat((warp, a["pc1"], out[0], a["ft1"], out[6]), dim=1)[0].numpy()      # (139, N)
            save[f"f{fr}_cls"], save[f"f{fr}_warp"], save[f"f{fr}_h"] = cls.numpy(), warp.numpy(), h2.numpy()
            save[f"f{fr}_nobj"] = np.int64(len(objs_curr))
            for i, o in enumerate(objs_curr):
                cols = o[0].numpy()
                ids = [int(np.nonzero((feats[:, mov] == cols[:, [c]]).all(0))[0][0]) for c in range(cols.shape[1])]
                save[f"f{fr}_obj{i}"] = mov[np.array(ids)]
            save[f"f{fr}_idx1"] = idx1.numpy() if idx1 is not None else np.zeros((1, 0), np.int64)
            save[f"f{fr}_ids"] = np.array(list(objects.keys()), np.int64)
            save[f"f{fr}_aff_mat"] = aff_mat.numpy()
            f8 = np.concatenate([feats[3:9, mov], feats[10:12, mov]]).T.astype(np.float64)
            dist = np.sqrt(((f8[:, None] - f8[None]) ** 2).sum(-1)) if len(mov) else np.zeros((0, 0))
            save[f"f{fr}_margin_cls"] = np.float64(np.abs(cls.numpy() - 0.5).min())
            save[f"f{fr}_margin_eps"] = np.float64(np.abs(dist - 1.5).min()) if dist.size else np.float64(np.inf)
            print("frame", fr, "moving", len(mov), "objects", len(objs_curr), "ids", list(objects.keys()),
                  "margins cls %.2e eps %.2e" % (save[f"f{fr}_margin_cls"], save[f"f{fr}_margin_eps"]))
            prev = {k: v.clone().detach() for k, v in objects.items()}
            h = h2.detach()
    save["max_id"] = np.int64(net.max_id)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "track4d_sequence.npz"), **save)
