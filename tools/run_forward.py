"""Minimal driver: a few fused forwards at (B, N) -- the command ncu wraps for single-kernel captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402


class Args:
    npoints = 512


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
net = Track4DBackbone(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda().eval()
d = synthetic.make_batch(B, 1024, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
h = torch.zeros(5, B, 128, device="cuda")
with torch.no_grad():
    for _ in range(steps):
        out = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h)
torch.cuda.synchronize()
print("ok", float(out[0].abs().max()))
