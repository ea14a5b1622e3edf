"""Frame-pair sharding across ranks (one process per GPU).  The hot path has no data-path collective:
every frame pair is independent in the forward pass, so a global batch is split by
pairs and the only communication is the bookkeeping below (barrier, max-over-ranks time, pair counts).
Works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced split of `total` frame pairs: -> (start, count); counts differ by at most 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(total), world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def job_throughput(local_pairs: int, local_ms: float, device=None):
    """Whole-job throughput: pairs of all ranks / MAX over ranks of the device time.  -> (pairs_total, ms_max)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(local_pairs), float(local_ms)
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    p = torch.tensor([int(local_pairs)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(p, op=dist.ReduceOp.SUM)
    return int(p.item()), float(t.item())


def gather_flow(flow_local: torch.Tensor, counts):
    """Optional result collection: all ranks' (b_r, 3, N) flow tensors -> list on every rank (all_gather,
    padded to the largest shard)."""
    world = dist.get_world_size()
    bmax = max(counts)
    pad = torch.zeros((bmax,) + tuple(flow_local.shape[1:]), dtype=flow_local.dtype, device=flow_local.device)
    pad[: flow_local.shape[0]] = flow_local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return [o[:c] for o, c in zip(out, counts)]


def allreduce_gradients(params, average: bool = True, group=None):
    """Data-parallel gradient reduction of the training step: ONE all_reduce over a flat fp32 bucket of every
    parameter gradient (1.8 M floats = 7.3 MB for the whole model: latency-bound on NVLink, so a single bucket),
    then scatter back.  Replaces the reference's single-process nn.DataParallel gradient gather
    (reference: src/models/model.py:38-40).  Parameters without a gradient contribute zeros so that every rank
    reduces the same layout.  Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return sum(p.numel() for p in params)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= world
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off: off + n].view_as(p).to(p.dtype)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off
