"""torch.profiler kernel table of one training step (modular path): where the 0.74 s at batch 256 go.
    python tools/profile_train.py [batch] -> gpurun_out/train_profile.txt"""
import os
import sys

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ratrack_b200 import synthetic, train  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402


class Args:
    npoints = 512


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = 1024
net = Track4DBackbone(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda()
opt = train.make_optimizer(net, lr=1e-4)
d = synthetic.make_batch(B, N, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
rng = np.random.default_rng(9)
gt_flow = t["pc1"] + torch.from_numpy(rng.normal(0, 0.4, (B, 3, N)).astype(np.float32)).cuda()
gt_cls = torch.from_numpy(rng.random((B, N)) < 0.3).cuda()
h0 = torch.zeros(5, B, 128, device="cuda")
for _ in range(2):
    train.train_step(net, opt, t["pc1"], t["pc2"], t["ft1"], t["ft2"], gt_flow, gt_cls, h0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train.train_step(net, opt, t["pc1"], t["pc2"], t["ft1"], t["ft2"], gt_flow, gt_cls, h0)
    torch.cuda.synchronize()
txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "train_profile.txt"), "w").write(f"batch {B}\n" + txt)
print(txt[-6000:])
