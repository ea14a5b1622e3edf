"""Host-side multi-rank logic on CPU: world_size 2, gloo backend (the N>1 path of bench.py minus the kernels)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ratrack_b200 import sharding


def test_shard_range_is_a_partition():
    for total in (0, 1, 7, 32, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == total
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            pos = 0
            for s, c in spans:
                assert s == pos
                pos += c
    with pytest.raises(ValueError):
        sharding.shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sharding.shard_range(1027, rank, world)
    pairs, ms = sharding.job_throughput(count, 10.0 + 5.0 * rank)
    counts = [sharding.shard_range(1027, r, world)[1] for r in range(world)]
    flow = torch.full((count, 3, 4), float(rank))
    parts = sharding.gather_flow(flow, counts)
    ok = pairs == 1027 and ms == 10.0 + 5.0 * (world - 1) and all(
        p.shape[0] == c and bool((p == float(r)).all()) for r, (p, c) in enumerate(zip(parts, counts)))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_rank_gloo_throughput_and_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _train_worker(rank, world, port, q):
    """Two ranks, each with half of the frame pairs and a different number of affinity entries: after the sharded
    loss + gradient all-reduce every rank must hold the loss value and the gradients of the single-process step."""
    from ratrack_b200 import losses

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    B, N = 4, 16
    lin = torch.nn.Linear(5, 4)                       # stand-in for the backbone: per-point flow (3) + cls logit (1)
    x = torch.randn(B, N, 5)
    gt_flow = torch.randn(B, 3, N)
    gt_cls = torch.rand(B, N) < 0.4
    n_aff = [5, 2]                                    # entries per rank (ragged)
    aff_in = torch.randn(sum(n_aff), 5)
    aff_gt = (torch.rand(sum(n_aff)) < 0.5).float()

    def terms(xs, gts, gcs, ai, ag, group_loss):
        y = lin(xs)
        flow = y[..., :3].permute(0, 2, 1)
        cls = torch.sigmoid(y[..., 3])
        aff = torch.sigmoid(lin(ai)[:, 0])
        sf = losses.flow_loss(flow, gts)
        seg = losses.motion_seg_loss(cls, gcs)
        trk = group_loss(aff, ag)
        return 0.5 * sf + 0.5 * trk + seg, trk

    # single-process reference over the whole batch (seg: equal positives per shard is not guaranteed, so the
    # reference value is the mean of the per-shard segmentation terms, which is what the all-reduce average yields)
    start, count = sharding.shard_range(B, rank, world)
    a0 = sum(n_aff[:rank])
    total, trk = terms(x[start:start + count], gt_flow[start:start + count], gt_cls[start:start + count],
                       aff_in[a0:a0 + n_aff[rank]], aff_gt[a0:a0 + n_aff[rank]], losses.sharded_affinity_loss)
    total.backward()
    n = sharding.allreduce_gradients(lin.parameters())
    got = [p.grad.clone() for p in lin.parameters()]
    trk_val = float(trk)

    # the same step with the synchronisation-free affinity all_gather (fixed-width blocks) and the bucketed gradient
    # reduction that starts inside backward: must give the same loss value and the same gradients
    lin.zero_grad()
    holder = torch.nn.Module()
    holder.pn_head = lin                                # "late" bucket
    holder.extra = torch.nn.Linear(2, 2)               # "early" bucket, one parameter of which never gets a gradient
    buckets = sharding.GradBuckets(holder)
    total2, trk2 = terms(x[start:start + count], gt_flow[start:start + count], gt_cls[start:start + count],
                         aff_in[a0:a0 + n_aff[rank]], aff_gt[a0:a0 + n_aff[rank]],
                         lambda a, g: losses.sharded_affinity_loss(a, g, max_entries=8))
    (total2 + 0.0 * holder.extra.weight.sum()).backward()
    n2 = buckets.finish()
    buckets.remove()
    same = abs(float(trk2) - trk_val) < 1e-6 and n2 == n + 6 and all(
        bool(torch.allclose(g, p.grad, rtol=1e-6, atol=1e-7)) for g, p in zip(got, lin.parameters()))

    lin.zero_grad()
    ref_total = 0.0
    for r in range(world):
        s, c = sharding.shard_range(B, r, world)
        y = lin(x[s:s + c])
        ref_total = ref_total + (0.5 * losses.flow_loss(y[..., :3].permute(0, 2, 1), gt_flow[s:s + c]) +
                                 losses.motion_seg_loss(torch.sigmoid(y[..., 3]), gt_cls[s:s + c])) / world
    ref_trk = losses.affinity_loss(torch.sigmoid(lin(aff_in)[:, 0]), aff_gt)   # ONE mean over all entries
    (ref_total + 0.5 * ref_trk).backward()
    ok = same and n == sum(p.numel() for p in lin.parameters()) and abs(trk_val - float(ref_trk)) < 1e-6
    for g, p in zip(got, lin.parameters()):
        ok = ok and bool(torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, ok))


def test_two_rank_gloo_sharded_loss_and_gradient_allreduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
