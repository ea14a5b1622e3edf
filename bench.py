#!/usr/bin/env python
"""bench.py -- frames/s of Track4D.backbone on synthetic 1024-point radar frame pairs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--total T] [--batch B] [--points P]

Workload (BASELINE.json configs[3], strong scaling): a job of T = 1024 frame pairs of P = 1024 points is sharded over
the N ranks (sharding.shard_range, contiguous, no data-path collective -- DESIGN.md "Multi-GPU"); every rank walks its
shard in micro-batches of B = 32 pairs, which is exactly BASELINE configs[1] (N=1024, batch 32, full backbone +
scene-flow forward, eval mode), so the N=1 line is configs[1] repeated over 1024 distinct pairs.
One "step" = one pass of the hot path over the rank's whole shard.

Printed JSON (one line, rank 0):
  value        frames(pairs)/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e          same metric through the public host-buffer API (pinned host -> H2D -> backbone -> D2H flow+cls per
               micro-batch, copies of neighbouring micro-batches overlapped with compute: infer_host_stream)
  roofline     the dominant kernel of the step, timed live with CUDA events on its launch stream
  cpu_baseline the CPU oracle port (torch-CPU dense layers + C/OpenMP pointnet2 ops) on a bounded sample: micro-batches of 8 pairs
               of the same workload for about 12 s (at most 1024 pairs); the sample is named in the record
  ref_gpu      (informational) the reference's own CUDA kernels (oracle/_ref) under torch fp32 modules evaluated op by op as the
               reference's modules do (model_utils.reference_dataflow: nn.Conv2d / cuDNN, none of this package's dense kernels)
  train        (informational, BASELINE configs[2]) the training step at batch 256 per GPU -- autograd over the package's tcgen05
               dense kernels and fused cost-volume / grouping kernels -- with the time and bytes of its two
               collectives (gradient all_reduce, affinity all_gather)
  train_cfg4   (informational, BASELINE configs[4]) the same step at N=3000 points, 16 pairs per GPU (= batch 128 on 8 GPUs)
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


def job_config(a, world, path, l2):
    return {"workload": f"synthetic N={a.points} pts, {a.total} frame pairs sharded {world} way(s) (configs[3]), micro-batch "
                        f"{a.batch} pairs = configs[1], full backbone+scene-flow forward",
            "total_pairs": a.total, "micro_batch": a.batch, "points": a.points, "npoints": 512, "path": path, "l2": l2,
            "parallelism": f"dp{world}"}


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


def cpu_reference_rate(pairs, micro, points, threads=None, min_seconds=0.0, max_pairs=None):
    """Oracle port of the reference path on the host cores: at least `pairs` pairs and `min_seconds` of work (bounded by
    `max_pairs`); returns (frames/s, cores, seconds, pairs done)."""
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
    while done < pairs or (time.perf_counter() - t0 < min_seconds and (max_pairs is None or done < max_pairs)):
        backbone_oracle.backbone(sd, c["pc1"], c["pc2"], c["ft1"], c["ft2"], h)
        done += micro
    dt = time.perf_counter() - t0
    return done / dt, cores, dt, done


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
    sample = f"{micro} of the {a.total} pairs of the step per timed step (micro-batch {micro}), N={a.points}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": job_config(a, a.gpus, "reference CPU path (oracle port)", "n/a (host run)"),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def _collective_ms(fn, reps=10):
    """Mean device time of one collective issued alone (CUDA events on the current stream, after 2 warm-up calls)."""
    import torch

    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def train_record(dev, rank, world, B, N, steps, warmup, label="configs[2] geometry"):
    """Training step (ratrack_b200/train.py) on synthetic frame pairs and synthetic targets: forward (train-mode BatchNorm)
    + track_4d_loss + backward + gradient all-reduce (overlapped with the tail of backward, ranks > 1) + Adam; weak scaling
    (B pairs per GPU).  -> dict (rank 0) with frames/s, step time, peak memory and the two collectives' time / bytes."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from ratrack_b200 import _cabi, losses, sharding, synthetic, train
    from ratrack_b200.model_utils import Track4DBackbone

    net = Track4DBackbone(Args())
    net.load_state_dict(synthetic.make_state_dict(net, seed=1234), strict=False)
    net = net.to(dev)
    opt = train.make_optimizer(net, lr=1e-4)
    d = synthetic.make_batch(B, N, seed=1234 + rank)
    t = {k: torch.from_numpy(v).to(dev) for k, v in d.items()}
    rng = np.random.default_rng(99 + rank)
    gt_flow = t["pc1"] + torch.from_numpy(rng.normal(0, 0.4, (B, 3, N)).astype(np.float32)).to(dev)
    gt_cls = torch.from_numpy(rng.random((B, N)) < 0.3).to(dev)
    n_aff = 12                                   # stand-in affinity entries per frame pair (association is outside this path)
    aff_gt = torch.from_numpy((rng.random(B * n_aff) < 0.25).astype(np.float32)).to(dev)
    h0 = torch.zeros(5, B, 128, device=dev)
    buckets = sharding.GradBuckets(net) if world > 1 else None

    def aff_fn(out):
        return torch.sigmoid(out[6][:, :n_aff].mean(dim=2)).reshape(-1)

    def step():
        return train.train_step(net, opt, t["pc1"], t["pc2"], t["ft1"], t["ft2"], gt_flow, gt_cls, h0, aff_fn, aff_gt,
                                buckets=buckets, max_entries=B * n_aff)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    torch.cuda.reset_peak_memory_stats()
    for _ in range(warmup):
        loss = step()[0]
    clocks = ClockSampler(dev.index)
    barrier()
    if rank == 0:
        clocks.start()
    _cabi.launch_count = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()[0]
    e1.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    pairs, ms = sharding.job_throughput(B * steps, e0.elapsed_time(e1), device=dev)
    # the two collectives of the step, each issued alone at its real size (0 at one rank: there is nothing to exchange)
    nparam = sum(p.numel() for p in net.parameters() if p.requires_grad)
    coll = {"grad_all_reduce": {"bytes": 4 * nparam, "ms": 0.0}, "affinity_all_gather": {"bytes_per_rank": 3 * 4 * B * n_aff, "ms": 0.0}}
    if world > 1:
        flat = torch.zeros(nparam, device=dev)
        block = torch.zeros(3, B * n_aff, device=dev)
        out = torch.empty(world, 3, B * n_aff, device=dev)
        coll["grad_all_reduce"]["ms"] = _collective_ms(lambda: dist.all_reduce(flat))
        coll["affinity_all_gather"]["ms"] = _collective_ms(lambda: dist.all_gather(list(out.unbind(0)), block))
        coll["grad_all_reduce"]["bus_gbs"] = 2 * (world - 1) / world * 4 * nparam / (coll["grad_all_reduce"]["ms"] * 1e-3) / 1e9
        coll["grad_all_reduce"]["overlap"] = "decoder/cost-volume bucket is reduced while pn_head is still in backward (sharding.GradBuckets)"
    if buckets is not None:
        buckets.remove()
    rec = {"metric": "frames/sec on Bx%d-pt radar pairs (training step: forward + multi-task loss + backward + Adam)" % N,
           "value": pairs / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms / steps, "scaling": "weak", "dtype": "f32",
           "config": {"workload": f"synthetic N={N} pts, batch={B} per GPU, forward+backward multi-task loss ({label})",
                      "batch_per_gpu": B, "points": N, "npoints": 512,
                      "path": "autograd over the package's own kernels: tcgen05 split-fp16 forward / dgrad / wgrad for the dense layers, fused channels-innermost cost volume and grouping, deterministic gradient kernels; BatchNorm / max-pool / GRU / Adam are torch", "parallelism": f"dp{world}"},
           "collectives": coll, "gpu_launches": _cabi.launch_count, "clocks": clk, "final_loss": float(loss),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del net, opt, t
    torch.cuda.empty_cache()
    return rec


def run_train(a):
    import torch
    import torch.distributed as dist

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
    rec = train_record(dev, rank, world, a.batch if a.batch != 32 else 256, a.points, a.steps, a.warmup)
    if rank == 0:
        rec.update({"higher_is_better": True, "vs_baseline": None, "data": "synthetic"})
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--total", type=int, default=1024, help="frame pairs of the whole job, sharded over the ranks (configs[3])")
    ap.add_argument("--batch", type=int, default=32, help="frame pairs per micro-batch (configs[1])")
    ap.add_argument("--points", type=int, default=1024)
    ap.add_argument("--modular", action="store_true", help="time the modular (unfused) path instead of the fused engine")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / ref_gpu legs")
    ap.add_argument("--no-train", action="store_true", help="skip the informational training sub-record")
    ap.add_argument("--train", action="store_true",
                    help="time ONLY the training step (BASELINE configs[2]: forward + multi-task loss + backward + Adam, "
                         "default batch 256 per GPU; --points 3000 --batch 16 = configs[4] on 8 GPUs); not the headline metric")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference(a)
    if a.train:
        return run_train(a)

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from ratrack_b200 import _cabi, sharding, synthetic
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
    # this rank's shard of the job: `count` distinct synthetic pairs, walked in micro-batches of B (the last one may be short)
    start, count = sharding.shard_range(a.total, rank, world)
    d = synthetic.make_batch(count, N, seed=1234 + rank)
    host = {k: torch.from_numpy(v).pin_memory() for k, v in d.items()}
    t = {k: v.to(dev) for k, v in host.items()}
    spans = [(o, min(B, count - o)) for o in range(0, count, B)]
    h0 = torch.zeros(5, B, 128, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step(profile_first=None):
        out = None
        for i, (o, c) in enumerate(spans):
            if profile_first is not None:
                eng.set_profile_events(*(profile_first if i == 0 else (None, None)))
            out = net.backbone(t["pc1"][o:o + c], t["pc2"][o:o + c], t["ft1"][o:o + c], t["ft2"][o:o + c], h0[:, :c])
        return out

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
            flush.zero_()                      # L2 flush between timed steps (outside the event pair)
            if eng:                            # fresh event pair per step around the dominant kernel of its first micro-batch
                dom_ev.append((torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)))
            e0.record()
            step(dom_ev[-1] if eng else None)
            e1.record()
        barrier()
        launches = (eng.launch_count() - eng_l0) if eng else _cabi.launch_count
        if eng:
            eng.set_profile_events(None, None)
            eng.check_status()
        prof = _cabi.profile
        _cabi.profile = None
        clk = clocks.stop() if rank == 0 else None
        ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
        dom_ms = [e0.elapsed_time(e1) for e0, e1 in (dom_ev if eng else prof["events"])]

        # ---- timed region 2: end to end through the host-buffer API -------------------------------
        def host_batches(reps):
            for _ in range(reps):
                for o, c in spans:
                    yield tuple(host[k][o:o + c] for k in ("pc1", "pc2", "ft1", "ft2"))

        for _ in net.infer_host_stream(host_batches(1)):
            pass
        barrier()
        ev2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev2[0].record()
        checksum = 0.0
        for flow, cls in net.infer_host_stream(host_batches(a.steps)):
            checksum += float(flow[0, 0, 0]) + float(cls[0, 0])      # the results are read on the host
        ev2[1].record()
        barrier()
        ms_e2e = ev2[0].elapsed_time(ev2[1])
        assert checksum == checksum, "NaN in the end-to-end results"
    pairs_total, ms = sharding.job_throughput(count * a.steps, ms, device=dev)          # SUM of pairs, MAX of device time
    _, ms_e2e = sharding.job_throughput(count * a.steps, ms_e2e, device=dev)

    train = train4 = None
    if not a.no_train and not a.modular:
        del t, flush
        net._engine = None
        torch.cuda.empty_cache()
        try:
            train = train_record(dev, rank, world, 256, N, steps=3, warmup=2)
        except Exception as e:  # informational only
            train = {"unavailable": str(e)[:200]}
        try:
            # BASELINE configs[4]: N~3000 (3 accumulated VoD frames), batch 128 over 8 GPUs = 16 pairs per GPU (weak scaling:
            # the same 16-pair shard per GPU at every N, so N=8 is the configuration itself)
            train4 = train_record(dev, rank, world, 16, 3000, steps=2, warmup=1,
                                  label="configs[4] geometry: N~3000, batch 128 / 8 GPUs = 16 per GPU")
        except Exception as e:  # informational only
            train4 = {"unavailable": str(e)[:200]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = _peaks()
    value = pairs_total / (ms * 1e-3)
    e2e = pairs_total / (ms_e2e * 1e-3)
    h2d = sum(v.numel() * 4 for v in host.values())
    d2h = count * 3 * N * 4 + count * N * 4
    # ---- roofline of the dominant kernel (definitions: DESIGN.md "Kernels and rooflines") -----------
    b0 = spans[0][1]
    if fused:
        from ratrack_b200 import engine
        roof = engine.roofline_of_dominant(b0, N, dom_ms, peaks, launch_pairs=(b0 + 1) // 2 if eng.num_lanes(b0) == 2 else b0)
        roof["lanes"] = eng.num_lanes(b0)
    else:
        # modular path: the largest launch of ours is group_points of the 514-channel embedding
        # (mse SA1, C=514, ns=8): algorithmic bytes 4*S*ns + 4*C*min(N,S*ns) + 4*C*S*ns per cloud
        per_call = sorted(dom_ms)[-max(1, len(dom_ms) // 36):]   # the biggest launches = that configuration
        avg = sum(per_call) / len(per_call)
        S, ns, C = 512, 8, 514
        bytes_ = b0 * (4 * S * ns + 4 * C * min(N, S * ns) + 4 * C * S * ns)
        ach = bytes_ / (avg * 1e-3) / 1e9
        roof = {"kernel": "gather_rows_kernel (group_points C=514 ns=8)", "bound": "hbm", "achieved": ach,
                "peak": peaks["hbm"], "peak_source": peaks["src"], "unit": "GB/s", "frac": ach / peaks["hbm"],
                "traffic": None, "avg_launch_ms": avg}
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "ms_per_micro_batch": ms / a.steps / len(spans), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": job_config(a, world, "fused" if fused else "modular",
                             "every step streams the rank's whole shard of distinct pairs (working set of one micro-batch "
                             "~150 MB > 126 MB L2); 256 MiB flush between timed steps"),
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clk, "roofline": roof,
    }
    if train is not None:
        out["train"] = train
    if train4 is not None:
        out["train_cfg4"] = train4
    if world == 1 and not a.no_cpu:
        v, cores, dt, done = cpu_reference_rate(pairs=32, micro=8, points=N, min_seconds=12.0, max_pairs=1024)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{done} pairs ({done // 8} micro-batches of 8) of the same N={N} workload, {dt:.1f} s"}
        try:
            t1 = {k: v[:B].to(dev) for k, v in host.items()}
            out["ref_gpu"] = ref_gpu_rate(net, t1, torch.zeros(5, t1["pc1"].size(0), 128, device=dev), 5, t1["pc1"].size(0))
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
    from ratrack_b200.model_utils import reference_dataflow

    ours, fused = U.pointnet2, net.use_fused
    U.pointnet2, net.use_fused = ref, False
    try:
        # op-by-op dataflow of the reference's modules: nn.Conv2d / cuDNN, torch Linear / BatchNorm, channel-major grouping
        with torch.no_grad(), reference_dataflow():
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


if __name__ == "__main__":
    main()
