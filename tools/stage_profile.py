"""Critical path of one fused forward, measured with CUDA events inside the engine (no profiler attached):
milliseconds between the stage boundaries of the feature-path stream, averaged over a few forwards.
    python tools/stage_profile.py [batch] [reps]     -> gpurun_out/stage_profile.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402


class Args:
    npoints = 512


B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
net = Track4DBackbone(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda().eval()
d = synthetic.make_batch(B, 1024, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
h = torch.zeros(5, B, 128, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
acc = {}
with torch.no_grad():
    for _ in range(3):
        net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h)
    net._engine.stage_profile(True)
    for _ in range(reps):
        flush.zero_()
        net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h)
        for name, ms in net._engine.stage_times():
            acc.setdefault(name, []).append(ms)
lines = [f"stage profile (CUDA events on the feature-path stream, B={B}, N=1024, mean of {reps} forwards, L2 flushed before each)"]
tot = 0.0
for name, v in acc.items():
    m = sum(v) / len(v)
    tot += m
    lines.append(f"{1e3 * m:9.1f} us  {name}")
lines.append(f"{1e3 * tot:9.1f} us  total (first mark to last)")
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/stage_profile.txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
