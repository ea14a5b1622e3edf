"""One forward + backward of the fused cost-volume training path (FeatureCorrelator on the rows kernels) -- for ncu captures.
    python tools/run_costvol_train.py [batch] [points]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.model_utils import FeatureCorrelator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
fc = FeatureCorrelator(16, in_channel=515, mlp=[256, 256, 256]).cuda()
d = synthetic.make_batch(B, N, seed=1)
pc1, pc2 = torch.from_numpy(d["pc1"]).cuda(), torch.from_numpy(d["pc2"]).cuda()
f1 = torch.randn(B, 256, N, device="cuda", requires_grad=True)
f2 = torch.randn(B, 256, N, device="cuda", requires_grad=True)
for _ in range(2):
    out = fc(pc1, pc2, f1, f2)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
