#!/usr/bin/env python
"""bench.py -- frames/s of Track4D.backbone on synthetic B x 1024-point radar frame pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--points N]

One "step" = one pass of the hot path (PNHead x3 + cost volume + flow decoder, eval mode) over one batch
of synthetic frame pairs.  Workload at N=1 is BASELINE.json configs[1]: N=1024 points, batch 32 per GPU
(weak scaling: every rank runs its own 32-pair shard, no data-path collective -- DESIGN.md "Multi-GPU").

Printed JSON (one line, rank 0):
  value        frames(pairs)/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public host-buffer API (pinned host -> H2D -> backbone -> D2H flow+cls)
  roofline     the dominant kernel of the step, timed live with CUDA events on its launch stream
  cpu_baseline the CPU oracle port (torch-CPU dense layers + C/OpenMP pointnet2 ops) on a bounded sample
  ref_gpu      (informational) the reference's own CUDA kernels (oracle/_ref) under the same torch modules
--impl reference times the reference's CPU path (the oracle port: the reference has no CPU implementation of
its native ops, and its Python cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def init_dist(dev):
    """NCCL process group for the N-GPU launch.  NCCL prints its version banner (and NCCL_DEBUG output) on the process's
    stdout while the communicator is created; stdout is reserved for the ONE JSON line of the contract, so file
    descriptor 1 points at stderr until the first collective has run."""
    import torch.distributed as dist

    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        import torch
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)

METRIC = "frames/sec on Bx1024-pt radar pairs (Track4D.backbone forward)"
UNIT = "frames/s"


class Args:
    npoints = 512


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_rate(pairs, micro, points, threads=None):
    """Oracle port of the reference path on the host cores: returns (frames/s, cores, seconds)."""
    import torch

    from oracle import backbone_oracle
    from ratrack_b200 import synthetic
    from ratrack_b200.model_utils import Track4DBackbone

    cores = threads or os.cpu_count()
    torch.set_num_threads(cores)
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    sd = synthetic.make_state_dict(Track4DBackbone(Args()), seed=1234)
    d = synthetic.make_batch(micro, points, seed=1234)
    c = {k: torch.from_numpy(v) for k, v in d.items()}
    h = torch.zeros(5, micro, 128)
    backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], h)  # warm-up (thread pools, mkldnn primitives)
    t0 = time.perf_counter()
    done = 0
    while done < pairs:
        backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], h)
        done += micro
    dt = time.perf_counter() - t0
    return done / dt, cores, dt


def run_reference(a):
    """--impl reference: the reference's CPU path (oracle port) on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import backbone_oracle
    from ratrack_b200 import synthetic
    from ratrack_b200.model_utils import Track4DBackbone

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    micro = 8
    sd = synthetic.make_state_dict(Track4DBackbone(Args()), seed=1234)
    d = synthetic.make_batch(micro, a.points, seed=1234)
    c = {k: torch.from_numpy(v) for k, v in d.items()}
    h = torch.zeros(5, micro, 128)
    for _ in range(a.warmup):
        backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], h)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], h)
    dt = time.perf_counter() - t0
    v = micro * a.steps / dt
    sample = f"{micro} of the {a.batch} pairs of the step per timed step (micro-batch {micro}), N={a.points}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic N={a.points} pts, batch={a.batch} per GPU, full backbone+scene-flow forward (configs[1])",
                   "batch_per_gpu": a.batch, "points": a.points, "npoints": 512, "path": "reference CPU path (oracle port)",
                   "l2": "n/a (host run)", "parallelism": f"dp{a.gpus}"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_train(a):
    """Training step (ratrack_b200/train.py) on synthetic frame pairs and synthetic targets: frames/s of
    forward (train-mode BatchNorm) + track_4d_loss + backward + gradient all-reduce (ranks > 1) + Adam."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from ratrack_b200 import _cabi, sharding, synthetic, train
    from ratrack_b200.model_utils import Track4DBackbone

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_dist(dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = a.batch if a.batch != 32 else 256
    N = a.points
    net = Track4DBackbone(Args())
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    net = net.to(dev)
    opt = train.make_optimizer(net, lr=1e-4)
    d = synthetic.make_batch(B, N, seed=1234 + rank)
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    rng = np.random.default_rng(99 + rank)
    gt_flow = t["pc1"] + torch.from_numpy(rng.normal(0, 0.4, (B, 3, N)).astype(np.float32)).to(dev)
    gt_cls = torch.from_numpy(rng.random((B, N)) < 0.3).to(dev)
    aff_gt = torch.from_numpy((rng.random(B * 12) < 0.25).astype(np.float32)).to(dev)
    h0 = torch.zeros(5, B, 128, device=dev)

    def aff_fn(out):   # stand-in affinity entries (the association module is outside this path): 12 per frame pair
        return torch.sigmoid(out[6][:, :12].mean(dim=2)).reshape(-1)

    def step():
        return train.train_step(net, opt, t["pc1"], t["pc2"], t["ft1"], t["ft2"], gt_flow, gt_cls, h0, aff_fn, aff_gt)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        loss = step()[0]
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    _cabi.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()[0]
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    pairs, ms = sharding.job_throughput(B * a.steps, e0.elapsed_time(e1), device=dev)
    if rank == 0:
        print(json.dumps({
            "metric": "frames/sec on Bx1024-pt radar pairs (training step: forward + multi-task loss + backward + Adam)",
            "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"synthetic N={N} pts, batch={B} per GPU, forward+backward multi-task loss (configs[2])",
                       "batch_per_gpu": B, "points": N, "npoints": 512, "path": "modular (CUDA pointnet2 ops + grad kernels under autograd)",
                       "l2": "working set per step (> 10 GB of activations) exceeds L2", "parallelism": f"dp{world}"},
            "gpu_launches": _cabi.launch_count, "clocks": clk, "final_loss": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frame pairs per GPU per step")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--modular", action="store_true", help="time the modular (unfused) path instead of the fused engine")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / ref_gpu legs")
    ap.add_argument("--train", action="store_true",
                    help="time the TRAINING step instead (BASELINE configs[2]: forward + multi-task loss + backward + Adam, "
                         "default batch 256 per GPU); not the headline metric")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a)
    if a.train:
        return run_train(a)

    import numpy as np
    import torch
    import torch.distributed as dist

    from ratrack_b200 import _cabi, synthetic
    from ratrack_b200.model_utils import Track4DBackbone

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        init_dist(dev)
    torch.backends.cudnn.allow_tf32 = False          # fp32 math everywhere (SURVEY.md hard part 3)
    torch.backends.cuda.matmul.allow_tf32 = False

    B, N = a.batch, a.points
    net = Track4DBackbone(Args())
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    net = net.to(dev).eval()
    fused = (not a.modular) and net.fused_available()
    net.use_fused = fused
    # every rank owns a different shard of the synthetic job (seed offset by rank)
    d = synthetic.make_batch(B, N, seed=1234 + rank)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
    t = {k: v.to(dev) for k, v in host.items()}
    h0 = torch.zeros(5, B, 128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        return net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h0)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dom = "rt_group_points"
    with torch.no_grad():
        for _ in range(a.warmup):
            step()
        eng = net._engine if fused else None
        dom_ev = []
        # ---- timed region 1: device-resident inputs ------------------------------------------------
        clocks = ClockSampler(local)
        barrier()
        if rank == 0:
            clocks.start()
        _cabi.launch_count = 0
        _cabi.profile = {"name": dom, "events": []}
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        eng_l0 = eng.launch_count() if eng else 0
        for e0, e1 in ev:
            flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
            if eng:                            # fresh event pair per step around the dominant kernel
                dom_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
                eng.set_profile_events(*dom_ev[-1])
            e0.record()
            step()
            e1.record()
        barrier()
        launches = (eng.launch_count() - eng_l0) if eng else _cabi.launch_count
        if eng:
            eng.set_profile_events(None, None)
        prof = _cabi.profile
        _cabi.profile = None
        clk = clocks.stop() if rank == 0 else None
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        dom_ms = [e0.elapsed_time(e1) for e0, e1 in (dom_ev if eng else prof["events"])]

        # ---- timed region 2: end to end through the host-buffer API -------------------------------
        for _ in range(2):
            net.infer_host(host["pc1"], host["pc2"], host["ft1"], host["ft2"])
        barrier()
        t0 = time.perf_counter()
        ev2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev2[0].record()
        for _ in range(a.steps):
            flow, cls = net.infer_host(host["pc1"], host["pc2"], host["ft1"], host["ft2"])
        ev2[1].record()
        barrier()
        ms_e2e = ev2[0].elapsed_time(ev2[1])
    from ratrack_b200 import sharding
    pairs_total, ms = sharding.job_throughput(B * a.steps, ms, device=dev)          # SUM of pairs, MAX of device time
    _, ms_e2e = sharding.job_throughput(B * a.steps, ms_e2e, device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    value = pairs_total / (ms * 1e-3)
    e2e = pairs_total / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * 4 for v in host.values())
    d2h = B * 3 * N * 4 + B * N * 4
    # ---- roofline of the dominant kernel (definitions: DESIGN.md "Kernels and rooflines") -----------
    if fused:
        from ratrack_b200 import engine
        roof = engine.roofline_of_dominant(B, N, dom_ms, peaks, launch_pairs=(B + 1) // 2 if eng.num_lanes(B) == 2 else B)
        roof["lanes"] = eng.num_lanes(B)
    else:
        # modular path: the largest launch of ours is group_points of the 514-channel embedding
        # (mse SA1, C=514, ns=8): algorithmic bytes 4*S*ns + 4*C*min(N,S*ns) + 4*C*S*ns per cloud
        per_call = sorted(dom_ms)[-max(1, len(dom_ms) // 36):]   # the biggest launches = that configuration
        avg = sum(per_call) / len(per_call)
        S, ns, C = 512, 8, 514
        bytes_ = B * (4 * S * ns + 4 * C * min(N, S * ns) + 4 * C * S * ns)
        ach = bytes_ / (avg * 1e-3) / 1e9
        roof = {"kernel": "gather_rows_kernel (group_points C=514 ns=8)", "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm"], "peak_source": peaks["src"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                "traffic": None, "avg_launch_ms": avg}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"synthetic N={N} pts, batch={B} per GPU, full backbone+scene-flow forward (configs[1])",
                   "batch_per_gpu": B, "points": N, "npoints": 512, "path": "fused" if fused else "modular",
                   "l2": "flushed between timed iterations (256 MiB memset)", "parallelism": f"dp{world}"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clk, "roofline": roof,
    }
    if world == 1 and not a.no_cpu:
        v, cores, dt = cpu_reference_rate(pairs=32, micro=8, points=N)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"32 pairs (4 micro-batches of 8) of the same N={N} workload, {dt:.1f} s"}
        try:
            out["ref_gpu"] = ref_gpu_rate(net, t, h0, a.steps, B)
        except Exception as e:  # informational only
            out["ref_gpu"] = {"unavailable": str(e)[:200]}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ref_gpu_rate(net, t, h0, steps, B):
    """Informational: the reference's own CUDA kernels (oracle/_ref, unmodified sources compiled for sm_100)
    behind the same torch modules = the reference's GPU pipeline on this B200."""
    import torch

    from oracle import ref_gpu
    from ratrack_b200.lib import pointnet2_utils as U

    ref = ref_gpu.load()
    if ref is None:
        return {"unavailable": "oracle/_ref/pointnet2_cuda.so not built"}
    from ratrack_b200.lib.pytorch_utils import PointwiseConv2d

    ours, fused = U.pointnet2, net.use_fused
    U.pointnet2, net.use_fused = ref, False
    PointwiseConv2d.use_gemm = False      # 1x1 convolutions through nn.Conv2d / cuDNN, as the reference's modules run them
    try:
        with torch.no_grad():
            for _ in range(3):
                net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h0)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                net.backbone(t["pc1"], t["pc2"], t["ft1"], t["ft2"], h0)
            e1.record()
            torch.cuda.synchronize()
        return {"value": B * steps / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
                "what": "reference CUDA kernels (oracle/_ref) + torch fp32 modules, same B200, same batch"}
    finally:
        U.pointnet2, net.use_fused = ours, fused
        PointwiseConv2d.use_gemm = True


if __name__ == "__main__":
    main()
