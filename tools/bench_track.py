"""Per-frame latency of the whole `Track4D.forward` (backbone + clustering + affinity + Sinkhorn + ids) at the reference's
operating point (batch 1), a sequence of frames each seeing the previous frame's objects.
    python tools/bench_track.py [points]  -> gpurun_out/track_bench.txt"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ratrack_b200 import synthetic  # noqa: E402
from ratrack_b200.track4d import Track4D  # noqa: E402


class Args:
    npoints = 512
    min_obj_points = 2


N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
net = Track4D(Args())
net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
net = net.cuda().eval()
frames = 24
d = synthetic.make_batch(frames, N, seed=1234)
t = {k: torch.from_numpy(v).cuda() for k, v in d.items()}
lines = []
with torch.no_grad():
    for rep in range(2):          # second pass is the timed one (first warms kernels, allocator, cuBLAS)
        prev, h = dict(), None
        net.max_id = 0
        tb = tt = 0.0
        nobj = 0
        for f in range(frames):
            a = [t[k][f:f + 1] for k in ("pc1", "pc2", "ft1", "ft2")]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = net.backbone(a[0], a[1], a[2], a[3], h)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            r = net.track(a[0], a[2], out, prev)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            tb += t1 - t0
            tt += t2 - t1
            h = r[0]
            prev = {k: v.clone().detach() for k, v in r[7].items()}
            nobj += len(r[9])
    lines.append(f"Track4D.forward, batch 1, N={N}, {frames} consecutive frames (mean objects per frame {nobj / frames:.1f}):")
    lines.append(f"  backbone (fused engine)                          {1e3 * tb / frames:8.3f} ms / frame")
    lines.append(f"  clustering + affinity + Sinkhorn + ids           {1e3 * tt / frames:8.3f} ms / frame")
    lines.append(f"  whole forward                                    {1e3 * (tb + tt) / frames:8.3f} ms / frame = {frames / (tb + tt):.0f} frames/s (wall clock, synchronised per stage)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "track_bench.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
