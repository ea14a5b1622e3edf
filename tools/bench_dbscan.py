"""DBSCAN timing: rt_dbscan (one kernel, labels stay on the device) against the reference's path
(device -> host copy + sklearn.cluster.DBSCAN.fit_predict on the CPU, src/models/track4d.py:111-118).
    python tools/bench_dbscan.py  -> gpurun_out/dbscan_bench.txt"""
import os
import sys
import time

import numpy as np
import torch
from sklearn.cluster import DBSCAN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ratrack_b200 import association  # noqa: E402

rng = np.random.default_rng(3)
lines = ["DBSCAN(eps=1.5, min_samples=2) on (n, 8) feature rows; microseconds per set",
         f"{'n':>6s} {'ours (device)':>14s} {'reference: D2H + sklearn':>26s} {'speed-up':>9s}"]
for n in (64, 300, 1024):
    k = max(2, n // 25)
    centres = rng.uniform(-40, 40, (k, 8))
    x = np.concatenate([centres[rng.integers(0, k, n // 2)] + rng.normal(0, 0.45, (n // 2, 8)), rng.uniform(-45, 45, (n - n // 2, 8))]).astype(np.float32)
    xg = torch.from_numpy(x).cuda()
    for _ in range(3):
        association.dbscan_labels(xg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        lab = association.dbscan_labels(xg)
    e1.record()
    torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / 20 * 1e3
    db = DBSCAN(eps=1.5, min_samples=2)
    db.fit_predict(x)
    t0 = time.perf_counter()
    for _ in range(5):
        want = db.fit_predict(xg.detach().cpu().numpy())
    ref = (time.perf_counter() - t0) / 5 * 1e6
    assert np.array_equal(lab.cpu().numpy(), want)
    lines.append(f"{n:6d} {ours:14.1f} {ref:26.1f} {ref / ours:9.1f}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "dbscan_bench.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
