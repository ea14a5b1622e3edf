"""Per-tensor parity report of Track4D.backbone on the GPU: modular path (torch fp32 dense layers + our ops) and
fused engine vs the CPU oracle replaying the same cost-volume neighbour sets.  Writes gpurun_out/parity_report.txt."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import backbone_oracle  # noqa: E402
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.model_utils import Track4DBackbone  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


class Args:
    npoints = 512


lines = []
for batch, n in ((1, 256), (2, 1024), (4, 1024)):
    d = synthetic.make_batch(batch, n, seed=1234)
    c = {k: torch.from_numpy(v) for k, v in d.items()}
    for fused in (False, 'simt', 'tc-cv', 'tc'):
        net = Track4DBackbone(Args())
        sd = synthetic.make_state_dict(net, seed=1234)
        net.load_state_dict(sd, strict=False)
        net.use_fused, net.capture_knn = bool(fused), True
        net = net.cuda().eval()
        t = {k: v.cuda() for k, v in c.items()}
        if fused:
            from ratrack_b200.engine import FusedBackbone
            net._engine = FusedBackbone(net)
            net._engine.set_flags(costvol_tc=fused != 'simt', mlp_tc=fused == 'tc')
        with torch.no_grad():
            out = net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], torch.zeros(5, batch, 128, device="cuda"))
            knn = net.cost_volume_neighbours(t["pc1"], t["pc2"])
        ref = backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], torch.zeros(5, batch, 128),
                                       knn_override=tuple(k.cpu().long() for k in knn))
        ref_free = backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], torch.zeros(5, batch, 128))
        if fused:
            torch.cuda.synchronize()
            net._engine.check_status()
        row = [f"B={batch} N={n} {('fused-' + fused) if fused else 'modular'}"]
        for nm, a, r in zip(["flow", "h", "cls", "cor", "f1", "f2", "prop"], out, ref):
            err = float((a.cpu() - r).abs().max())
            scale = max(1.0, float(r.abs().max()))
            row.append(f"{nm}: {err:.2e} (/{scale:.1f} = {err / scale:.1e})")
        # how many rows of the neighbour sets differ from the oracle's own torch.topk choice
        o12 = backbone_oracle.knn_point(16, c["pc2"].permute(0, 2, 1), c["pc1"].permute(0, 2, 1))
        nd = int((torch.sort(o12, -1)[0] != torch.sort(knn[0].cpu().long(), -1)[0]).any(-1).sum())
        row.append(f"knn12 rows differing from torch-CPU topk: {nd}")
        row.append(f"flow vs free-running oracle: {float((out[0].cpu() - ref_free[0]).abs().max()):.2e}")
        lines.append(" | ".join(row))
        print(lines[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
open("gpurun_out/parity_report.txt", "w").write("\n".join(lines) + "\n")
