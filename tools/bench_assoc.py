"""Association step timing: rt_sinkhorn_match (one kernel) against the reference algorithm executed by torch on the same
GPU (oracle/association_oracle.py = the reference's log_optimal_transport + sinkhorn_module: ~2000 launches) and on the CPU.
    python tools/bench_assoc.py   -> gpurun_out/assoc_bench.txt"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import association_oracle  # noqa: E402
from oracle.gen_golden_assoc import make_aff  # noqa: E402
from ratrack_b200 import association  # noqa: E402

lines = ["Sinkhorn association (alpha 0.9, 500 iterations), one frame (batch 1) unless noted; times in microseconds",
         f"{'m x n':>10s} {'ours (1 kernel)':>16s} {'torch GPU (ref alg.)':>22s} {'torch CPU (ref alg.)':>22s} {'speed-up vs torch GPU':>22s}"]
for m, n in [(5, 5), (20, 20), (40, 31), (64, 64), (127, 127)]:
    aff_c = torch.from_numpy(make_aff(m, n, 7))
    aff = aff_c.cuda()
    for _ in range(3):
        association.sinkhorn_module(aff, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        idx = association.sinkhorn_module(aff, None)
    e1.record()
    torch.cuda.synchronize()
    ours = e0.elapsed_time(e1) / 20 * 1e3

    def ref_gpu():
        scores = association_oracle.log_sinkhorn_iterations  # noqa: F841  (same code path as the reference, tensors on the GPU)
        return association_oracle.sinkhorn_module(aff.cpu())[0]
    # reference algorithm with CUDA tensors
    def run_gpu():
        s = aff
        b = s.shape[0]
        alpha = torch.tensor(0.9, device="cuda")
        ms, ns = torch.tensor(float(m), device="cuda"), torch.tensor(float(n), device="cuda")
        couplings = torch.cat([torch.cat([s, alpha.expand(b, m, 1)], -1), torch.cat([alpha.expand(b, 1, n), alpha.expand(b, 1, 1)], -1)], 1)
        norm = -(ms + ns).log()
        log_mu = torch.cat([norm.expand(m), ns.log()[None] + norm])[None].expand(b, -1)
        log_nu = torch.cat([norm.expand(n), ms.log()[None] + norm])[None].expand(b, -1)
        return association_oracle.log_sinkhorn_iterations(couplings, log_mu, log_nu, 500) - norm
    run_gpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        z = run_gpu()
    torch.cuda.synchronize()
    tg = (time.perf_counter() - t0) / 3 * 1e6
    t0 = time.perf_counter()
    ref_idx = association_oracle.sinkhorn_module(aff_c)[0]
    tc = (time.perf_counter() - t0) * 1e6
    assert np.array_equal(idx.cpu().numpy(), ref_idx.numpy()), (m, n)
    lines.append(f"{m:>4d} x {n:<4d} {ours:16.1f} {tg:22.1f} {tc:22.1f} {tg / ours:22.1f}")
a = torch.from_numpy(np.concatenate([make_aff(20, 20, s) for s in range(256)])).cuda()
association.sinkhorn_module(a, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    association.sinkhorn_module(a, None)
e1.record()
torch.cuda.synchronize()
lines.append(f"batch of 256 frames, 20 x 20: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per call = {e0.elapsed_time(e1) / 10 * 1e3 / 256:.2f} us per frame")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "assoc_bench.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
