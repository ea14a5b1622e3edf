import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic
from ratrack_b200.track4d import Track4D
from oracle.gen_golden_track_seq import FRAMES, POINTS, SEED
class Args:
    npoints = 512
    min_obj_points = 2
g = np.load("tests/golden/track4d_sequence.npz")
net = Track4D(Args()); net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False); net = net.cuda().eval()
d = synthetic.make_batch(FRAMES, POINTS, seed=SEED)
prev, h = dict(), None
with torch.no_grad():
    for fr in range(3):
        a = {k: torch.from_numpy(v[fr:fr + 1]).cuda() for k, v in d.items()}
        out = net.backbone(a["pc1"], a["pc2"], a["ft1"], a["ft2"], h)
        r = net.track(a["pc1"], a["ft1"], out, prev)
        h = r[0]
        cls_ref = g[f"f{fr}_cls"]
        mov = np.nonzero(cls_ref[0] > 0.5)[0]
        print("frame", fr, "objs", len(r[9]), "golden", int(g[f"f{fr}_nobj"]), "sizes", [o.shape[2] for o in r[9]], "golden sizes", [len(g[f"f{fr}_obj{i}"]) for i in range(int(g[f"f{fr}_nobj"]))])
        print("  cls err", np.abs(out[2].cpu().numpy() - cls_ref).max(), "warp err", np.abs(r[1].cpu().numpy() - g[f"f{fr}_warp"]).max(), "mask equal", np.array_equal(out[2].cpu().numpy() > 0.5, cls_ref > 0.5))
        prev = {k: v.clone() for k, v in r[7].items()}
        if fr == 0:
            warp = r[1]
            feats = torch.cat((warp, a["pc1"], warp - a["pc1"], a["ft1"]), dim=1)[0]
            for i, o in enumerate(r[9][:4]):
                cols = o[0, :6]
                got = np.array([int(torch.nonzero((feats[:6, mov] == cols[:, [c]]).all(0))[0, 0]) for c in range(cols.shape[1])])
                print("   obj", i, "ours", mov[got], "golden", g[f"f0_obj{i}"])
