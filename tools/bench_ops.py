"""Per-op roofline table of the Seam-A kernels (the `pointnet2_cuda` entry points) at BASELINE configs[1]
geometry: 64 clouds (32 pairs), N=1024, S=512.  For every op: CUDA-event time of OUR kernel and of the
reference's own kernel (oracle/_ref, unmodified sources compiled for sm_100) on the same buffers, the
algorithmic HBM bytes of SURVEY.md section 8(d), achieved GB/s and the fraction of the measured HBM peak.
L2 is flushed before every timed launch (256 MiB memset), so the numbers are cold-cache.

    python tools/bench_ops.py [--once]      # --once: one launch per op (for an ncu capture)
Writes gpurun_out/ops_roofline.json and prints a table.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_gpu  # noqa: E402  (reference kernels, comparison only)
from ratrack_b200 import pointnet2_cuda as ours  # noqa: E402
from ratrack_b200 import _cabi, synthetic  # noqa: E402

ONCE = "--once" in sys.argv
REPS = 1 if ONCE else 20
B, N, S = 64, 1024, 512


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p))["hbm_gbs"] if os.path.exists(p) else 6650.0


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    ref = ref_gpu.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    d = synthetic.make_batch(B // 2, N, seed=1234)
    xyz = torch.from_numpy(np.ascontiguousarray(np.concatenate([d["pc1"], d["pc2"]]).transpose(0, 2, 1))).to(dev)  # (B,N,3)
    g = torch.Generator(device="cpu").manual_seed(1234)

    def rnd(*shape):
        return torch.randn(*shape, generator=g).to(dev)

    def timed(fn):
        ts = []
        for _ in range(0 if ONCE else 3):   # --once: exactly one launch per op reaches the profiler
            fn()
        for _ in range(REPS):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    rows = []

    def report(name, kernel, bytes_, fn_ours, fn_ref, note="", evals=None):
        t_o = timed(fn_ours)
        t_r = timed(fn_ref) if (ref is not None and fn_ref is not None and not ONCE) else None
        gbs = bytes_ / (t_o * 1e-3) / 1e9
        rows.append({"op": name, "kernel": kernel, "algorithmic_bytes": bytes_, "ours_us": t_o * 1e3,
                     "ref_kernel_us": None if t_r is None else t_r * 1e3, "achieved_gbs": gbs,
                     "frac_of_hbm_peak": gbs / peaks(), "speedup_vs_ref_kernel": None if t_r is None else t_r / t_o,
                     # search ops are bound by point-pair distance evaluations, not HBM: algorithmic pair evaluations / time
                     "pair_distance_evals": evals, "gevals_per_s": None if evals is None else evals / (t_o * 1e-3) / 1e9,
                     "note": note})

    # ---- FPS (SA1: N -> S; SA2/3: S -> S) ----------------------------------------------------------------
    for n_in in (N, S):
        x = xyz[:, :n_in].contiguous()
        temp = torch.empty(B, n_in, device=dev)
        idx = torch.empty(B, S, dtype=torch.int32, device=dev)

        def f(mod):
            temp.fill_(1e10)
            mod.furthest_point_sampling_wrapper(B, n_in, S, x, temp, idx)
        report(f"furthest_point_sample n={n_in} m={S}", "fps_reg_kernel", B * (12 * n_in + 4 * S),
               lambda: f(ours), lambda: f(ref), "latency-bound: m-1 dependent rounds; includes the temp fill", evals=B * (S - 1) * n_in)
    fps_idx = idx.clone()
    new_xyz = torch.gather(xyz[:, :S], 1, fps_idx.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()   # (B,S,3)

    # ---- ball_query: the six (radius, nsample) of one PNHead ----------------------------------------------
    for n_in, r, ns in ((N, 2.0, 4), (N, 4.0, 8), (S, 4.0, 8), (S, 8.0, 16), (S, 8.0, 16), (S, 16.0, 32)):
        x = xyz[:, :n_in].contiguous()
        q = new_xyz if n_in == S else torch.gather(xyz, 1, torch.randint(0, N, (B, S, 1), generator=g).to(dev).expand(-1, -1, 3)).contiguous()
        idx = torch.zeros(B, S, ns, dtype=torch.int32, device=dev)
        report(f"ball_query n={n_in} r={r} ns={ns}", "ball_query_kernel", B * (12 * n_in + 12 * S + 4 * S * ns),
               lambda: ours.ball_query_wrapper(B, n_in, S, r, ns, q, x, idx),
               lambda: ref.ball_query_wrapper(B, n_in, S, r, ns, q, x, idx), evals=B * S * n_in)
    bq_idx = idx.clone()   # (B,S,32) over S points

    # ---- group_points ----------------------------------------------------------------------------------
    for c, n_in, ns in ((64, S, 32), (514, N, 8), (3, N, 8)):
        pts = rnd(B, c, n_in)
        gi = torch.randint(0, n_in, (B, S, ns), generator=g).int().to(dev) if ns != 32 else bq_idx
        out = torch.empty(B, c, S, ns, device=dev)
        report(f"group_points C={c} n={n_in} ns={ns}", "group_points_kernel",
               B * (4 * S * ns + 4 * c * min(n_in, S * ns) + 4 * c * S * ns),
               lambda: ours.group_points_wrapper(B, c, n_in, S, ns, pts, gi, out),
               lambda: ref.group_points_wrapper(B, c, n_in, S, ns, pts, gi, out))
        if c == 64:
            gp = torch.zeros(B, c, n_in, device=dev)
            go = rnd(B, c, S, ns)

            def fg(mod):
                gp.zero_()
                mod.group_points_grad_wrapper(B, c, n_in, S, ns, go, gi, gp)
            report(f"group_points_grad C={c} n={n_in} ns={ns}", "group_points_grad_kernel",
                   B * (4 * S * ns + 4 * c * S * ns + 4 * c * n_in), lambda: fg(ours), lambda: fg(ref), "includes zero fill")

    # ---- gather_points (new_xyz) ------------------------------------------------------------------------
    xt = xyz.transpose(1, 2).contiguous()
    out = torch.empty(B, 3, S, device=dev)
    report(f"gather_points C=3 n={N} m={S}", "gather_points_kernel", B * (4 * S + 12 * S + 12 * S),
           lambda: ours.gather_points_wrapper(B, 3, N, S, xt, fps_idx, out),
           lambda: ref.gather_points_wrapper(B, 3, N, S, xt, fps_idx, out), "12 KB per cloud: launch-latency bound")

    # ---- three_nn / three_interpolate -------------------------------------------------------------------
    for n_u, m_k in ((S, S), (N, S)):
        unk = xyz[:, :n_u].contiguous()
        kn = new_xyz
        d2 = torch.empty(B, n_u, 3, device=dev)
        ni = torch.empty(B, n_u, 3, dtype=torch.int32, device=dev)
        report(f"three_nn n={n_u} m={m_k}", "three_nn_kernel", B * (12 * n_u + 12 * m_k + 24 * n_u),
               lambda: ours.three_nn_wrapper(B, n_u, m_k, unk, kn, d2, ni),
               lambda: ref.three_nn_wrapper(B, n_u, m_k, unk, kn, d2, ni), evals=B * n_u * m_k)
    w = torch.softmax(rnd(B, N, 3), -1).contiguous()
    for c, n_u in ((64, S), (128, N)):
        pts = rnd(B, c, S)
        nidx = ni[:, :n_u].contiguous()
        ww = w[:, :n_u].contiguous()
        out = torch.empty(B, c, n_u, device=dev)
        report(f"three_interpolate C={c} m={S} n={n_u}", "three_interpolate_kernel", B * (4 * c * S + 24 * n_u + 4 * c * n_u),
               lambda: ours.three_interpolate_wrapper(B, c, S, n_u, pts, nidx, ww, out),
               lambda: ref.three_interpolate_wrapper(B, c, S, n_u, pts, nidx, ww, out))
        if c == 128:
            gp = torch.zeros(B, c, S, device=dev)
            go = rnd(B, c, n_u)

            def fi(mod):
                gp.zero_()
                mod.three_interpolate_grad_wrapper(B, c, n_u, S, go, nidx, ww, gp)
            report(f"three_interpolate_grad C={c} m={S} n={n_u}", "three_interpolate_grad_kernel",
                   B * (4 * c * n_u + 24 * n_u + 4 * c * S), lambda: fi(ours), lambda: fi(ref), "includes zero fill")

    # ---- knn (Seam A, k=16) and the cost-volume kNN (expanded form; the reference uses matmul + topk) -----
    d2 = torch.empty(B, N, 16, device=dev)
    ki = torch.empty(B, N, 16, dtype=torch.int32, device=dev)
    report(f"knn k=16 n={N} m={N}", "knn_kernel", B * (12 * N + 12 * N + 8 * N * 16),
           lambda: ours.knn_wrapper(B, N, N, 16, xyz, xyz, d2, ki), lambda: ref.knn_wrapper(B, N, N, 16, xyz, xyz, d2, ki), evals=B * N * N)
    st = torch.cuda.current_stream().cuda_stream

    def torch_knn():
        dist = -2 * torch.matmul(xyz, xyz.permute(0, 2, 1))
        dist += torch.sum(xyz ** 2, -1).view(B, N, 1)
        dist += torch.sum(xyz ** 2, -1).view(B, 1, N)
        torch.topk(torch.clamp_min(dist, 0.0), 16, dim=-1, largest=False, sorted=False)
    report(f"knn_point (cost volume) k=16 n={N} m={N}", "knn_expanded_warp_kernel", B * (12 * N + 12 * N + 4 * N * 16),
           lambda: _cabi.call("rt_knn_expanded", B, N, N, 16, xyz.data_ptr(), xyz.data_ptr(), ki.data_ptr(), st),
           None if ONCE else torch_knn, "reference column = torch matmul + topk (model_utils.py:85-99), materialises (B,N,N)", evals=B * N * N)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if not ONCE:
        json.dump({"geometry": {"clouds": B, "N": N, "S": S}, "hbm_peak_gbs": peaks(), "l2": "flushed before every launch",
                   "reps": REPS, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "ops_roofline.json"), "w"), indent=1)
    print(f"{'op':52s} {'ours us':>9s} {'ref us':>9s} {'x':>6s} {'GB/s':>8s} {'frac':>6s} {'Gevals/s':>9s}")
    for r in rows:
        ru = f"{r['ref_kernel_us']:9.1f}" if r["ref_kernel_us"] else "        -"
        sp = f"{r['speedup_vs_ref_kernel']:6.1f}" if r["speedup_vs_ref_kernel"] else "     -"
        ge = f"{r['gevals_per_s']:9.1f}" if r["gevals_per_s"] else "        -"
        print(f"{r['op']:52s} {r['ours_us']:9.1f} {ru} {sp} {r['achieved_gbs']:8.1f} {r['frac_of_hbm_peak']:6.3f} {ge}")


if __name__ == "__main__":
    main()
