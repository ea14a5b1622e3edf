"""One small training step (batch 1, N=256) through every kernel of the training path -- for compute-sanitizer runs.
    python tools/run_train_small.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ratrack_b200 import synthetic, train  # noqa: E402
from ratrack_b200.lib import dense_tc  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402


class Args:
    npoints = 512


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dense_tc.MIN_ROWS = 0          # every dense layer on the tensor-core kernels, whatever its row count
B, N = 1, 256
net = Track4DBackbone(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda()
opt = train.make_optimizer(net, lr=1e-4)
d = synthetic.make_batch(B, N, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
rng = np.random.default_rng(9)
gt_flow = t["pc1"] + torch.from_numpy(rng.normal(0, 0.4, (B, 3, N)).astype(np.float32)).cuda()
gt_cls = torch.from_numpy(rng.random((B, N)) < 0.3).cuda()
loss = train.train_step(net, opt, t["pc1"], t["pc2"], t["ft1"], t["ft2"], gt_flow, gt_cls, torch.zeros(5, B, 128, device="cuda"))[0]
torch.cuda.synchronize()
print("loss", float(loss))
