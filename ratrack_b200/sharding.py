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


class GradBuckets:
    """Gradient reduction that starts before backward has finished.  The parameters are split into flat fp32 buckets in
    the order their gradients become final during backward (the decoder / cost-volume parameters first, pn_head -- the
    first module of the forward pass -- last); a post-accumulate-grad hook counts arrivals and issues the bucket's
    all_reduce (async) the moment its last gradient lands, so the decoder bucket travels while pn_head is still being
    differentiated.  `finish()` issues what is left, waits, averages and writes the reduced values back to `.grad`.
    Same result as `allreduce_gradients` (one bucket after backward); replaces the reference's nn.DataParallel gather
    (reference: src/models/model.py:38-40)."""

    def __init__(self, module, late_prefixes=("pn_head.",), group=None, average=True):
        self.group, self.average = group, average
        named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        late = [p for n, p in named if n.startswith(tuple(late_prefixes))]
        early = [p for n, p in named if not n.startswith(tuple(late_prefixes))]
        self.buckets = [b for b in (early, late) if b]
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self._owner = {}
        self._hooks = []
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._owner[p] = bi
                self._hooks.append(p.register_post_accumulate_grad_hook(self._arrived))
        self.reset()

    def reset(self):
        self._seen = [0] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)

    def _launch(self, bi):
        bucket = self.buckets[bi]
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        self._flat[bi] = flat
        if self.world > 1:
            self._work[bi] = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _arrived(self, p):
        bi = self._owner[p]
        self._seen[bi] += 1
        if self._seen[bi] == len(self.buckets[bi]) and self._flat[bi] is None:
            self._launch(bi)

    def finish(self):
        """-> number of gradient elements reduced."""
        total = 0
        for bi, bucket in enumerate(self.buckets):
            if self._flat[bi] is None:        # some parameter of the bucket got no gradient this step
                self._launch(bi)
            if self._work[bi] is not None:
                self._work[bi].wait()
            flat = self._flat[bi]
            if self.average and self.world > 1:
                flat /= self.world
            off = 0
            for p in bucket:
                n = p.numel()
                g = flat[off: off + n].view_as(p).to(p.dtype)
                if p.grad is None:
                    p.grad = g.clone()
                else:
                    p.grad.copy_(g)
                off += n
            total += off
        self.reset()
        return total

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
