"""Dump torch-GPU results of the reference's expanded-form square_distance (model_utils.py:17-39) so the
rounding order of cuBLAS' K=3 inner product and torch's 3-term sum can be identified offline
(SURVEY.md App. A.8 / E.4).  Writes gpurun_out/probe_sqdist.npz."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ratrack_b200 import synthetic  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
d = synthetic.make_batch(2, 1024, seed=1234)
src = torch.from_numpy(np.ascontiguousarray(d["pc1"].transpose(0, 2, 1))).cuda()
dst = torch.from_numpy(np.ascontiguousarray(d["pc2"].transpose(0, 2, 1))).cuda()
mm = torch.matmul(src, dst.permute(0, 2, 1))
s2 = torch.sum(src ** 2, -1)
d2 = torch.sum(dst ** 2, -1)
dist = -2 * mm
dist += s2.view(2, 1024, 1)
dist += d2.view(2, 1, 1024)
dist = torch.maximum(dist, torch.zeros_like(dist))
# also the permuted (non-contiguous) path exactly as the reference calls it: pc (B,3,N).permute(0,2,1)
srcp = torch.from_numpy(d["pc1"]).cuda().permute(0, 2, 1)
dstp = torch.from_numpy(d["pc2"]).cuda().permute(0, 2, 1)
mm_p = torch.matmul(srcp, dstp.permute(0, 2, 1))
s2_p = torch.sum(srcp ** 2, -1)
idx = torch.topk(dist, 16, dim=-1, largest=False, sorted=False)[1]
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed("gpurun_out/probe_sqdist.npz", mm=mm.cpu().numpy(), s2=s2.cpu().numpy(), d2=d2.cpu().numpy(),
                    dist=dist.cpu().numpy(), mm_p=mm_p.cpu().numpy(), s2_p=s2_p.cpu().numpy(), topk=idx.cpu().numpy())
print("ok", torch.cuda.get_device_name(0), torch.version.cuda)
