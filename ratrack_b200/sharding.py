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
