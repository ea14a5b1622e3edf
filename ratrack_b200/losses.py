"""Loss functions of the training step, with the reference's names and arithmetic
(reference: src/losses/loss.py:8-31 track_4d_loss, :48-72 affinity_loss, :85-89 flow_loss, :124-146 motion_seg_loss).

The reference is written for batch 1 (`sc_loss[0]`, a 1-D `gt_cls` mask); these mirrors give bit-for-bit the
same value at batch 1 and generalise to a batch by averaging over its frame pairs.  They are tensor-only: the
label plumbing that produces `gt_flow`, `gt_cls` and the ground-truth affinity entries (open3d boxes, dict
look-ups: models/utils/track4d_utils.py) is outside this package's path.

Multi-GPU: flow and segmentation terms are per-shard means, made global by the gradient all-reduce (equal shard
sizes); the affinity term is a mean over the entries of every rank, so its entries are all-gathered
(`sharded_affinity_loss`, the one collective the loss needs -- SURVEY.md section 8e).
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F


def flow_loss(pc1_wrap, gt_flow, reduction="mean"):
    """Mean over the points of |pc1_wrap - gt_flow|_2; pc1_wrap, gt_flow (B,3,N).  reference :85-89.
    reduction "first" = the reference literally (`sc_loss[0]`: frame pair 0 only), "mean" = mean over the batch
    (identical at batch 1), "none" = (B,)."""
    sc = (pc1_wrap - gt_flow).pow(2).sum(dim=1).sqrt().mean(dim=1)
    if reduction == "first":
        return sc[0]
    if reduction == "none":
        return sc
    return sc.mean()


def motion_seg_loss(pred_cls, gt_cls, nan_to_zero=False):
    """0.4 * BCE over the moving points + 0.6 * BCE over the static ones.  reference :124-146.
    pred_cls (B,N) probabilities, gt_cls (N,) or (B,N) bool.  Mask-weighted sums instead of the reference's boolean
    indexing (no device -> host sync).  An empty class makes the reference's mean-over-nothing NaN, and
    track_4d_loss then replaces the whole term by 0 (reference :14-19): `nan_to_zero` does that here without a NaN
    ever entering the autograd graph."""
    gt = gt_cls.to(device=pred_cls.device)
    if gt.dim() == 1:
        gt = gt.unsqueeze(0).expand_as(pred_cls)
    pos = (gt == True).to(torch.float32)  # noqa: E712  (the reference's own comparison)
    neg = (gt == False).to(torch.float32)  # noqa: E712
    # element-wise nn.BCELoss (log clamped at -100, analytic backward that stays finite at p = 0 or 1 exactly -- a sigmoid
    # output saturates to 1.0 in fp32 -- which a hand-written -log(p) / -log1p(-p) would not)
    bce = F.binary_cross_entropy(pred_cls.float(), pos, reduction="none")
    bce_pos = bce_neg = bce
    n_pos, n_neg = pos.sum(), neg.sum()
    loss = 0.4 * (bce_pos * pos).sum() / n_pos.clamp_min(1.0) + 0.6 * (bce_neg * neg).sum() / n_neg.clamp_min(1.0)
    valid = (n_pos > 0) & (n_neg > 0)
    return torch.where(valid, loss, torch.full_like(loss, 0.0 if nan_to_zero else float("nan")))


def affinity_ground_truth(ids_prev, ids_curr, device=None):
    """Row-major 0/1 vector: entry (i,j) = 1 iff ids_prev[i] == ids_curr[j].  reference :49-66 (the double loop over
    mappings_prev.keys() x mappings_curr.keys())."""
    a = torch.as_tensor(list(ids_prev), device=device).view(-1, 1)
    b = torch.as_tensor(list(ids_curr), device=device).view(1, -1)
    return (a == b).to(torch.float32).reshape(-1)


def affinity_loss(aff_list, aff_gt_list):
    """F.binary_cross_entropy(aff_list, aff_gt_list) (mean over the entries); 0 when there are none.  reference :67-72."""
    if aff_list.numel() == 0:
        return torch.zeros((), device=aff_list.device)
    return F.binary_cross_entropy(aff_list.float(), aff_gt_list.float())


class _AllGatherEntries(torch.autograd.Function):
    """all_gather of one variable-length fp32 vector per rank (padded to the longest), concatenated in rank order.
    Backward hands every rank the gradient slice of its own entries."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        n = torch.tensor([x.numel()], device=x.device, dtype=torch.int64)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n, group=group)
        counts = [int(c) for c in counts]
        width = max(max(counts), 1)
        pad = torch.zeros(width, device=x.device, dtype=x.dtype)
        pad[: x.numel()] = x.detach().reshape(-1)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        ctx.offset = sum(counts[:rank])
        ctx.count = counts[rank]
        return torch.cat([p[:c] for p, c in zip(parts, counts)])

    @staticmethod
    def backward(ctx, g):
        return g[ctx.offset: ctx.offset + ctx.count].clone(), None


def all_gather_entries(x, group=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x.reshape(-1)
    return _AllGatherEntries.apply(x.reshape(-1), group)


class _AllGatherFixed(torch.autograd.Function):
    """all_gather of a (rows, width) block of the same shape on every rank -> (world * rows... ) stacked as (world, rows, width).
    No size exchange, hence no device -> host synchronisation.  Backward returns the rank's own slice."""

    @staticmethod
    def forward(ctx, x, group):
        world = dist.get_world_size(group)
        ctx.rank = dist.get_rank(group)
        out = torch.empty((world,) + tuple(x.shape), device=x.device, dtype=x.dtype)
        dist.all_gather(list(out.unbind(0)), x.detach().contiguous(), group=group)   # views of `out`: gathered in place
        return out

    @staticmethod
    def backward(ctx, g):
        return g[ctx.rank].clone(), None


def sharded_affinity_loss(aff_list, aff_gt_list, group=None, grad_average=True, max_entries=None):
    """Affinity BCE as ONE mean over the entries of all ranks (what the single-process reference computes over its
    whole batch).  Every rank returns the same value.  With `grad_average` the local gradient is scaled by the world
    size, so that the usual gradient all-reduce *average* yields exactly the gradient of the global mean.

    `max_entries` (an upper bound on one rank's entry count, e.g. 20 x 20 objects per frame x frames per shard) selects
    the synchronisation-free form: every rank contributes one fixed-width block [prediction | target | validity] to a
    single all_gather and the mean is a masked sum -- no counts travel to the host.  Without it the ragged vectors are
    gathered after one exchange of their lengths (a host synchronisation per step)."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return affinity_loss(aff_list, aff_gt_list)
    if max_entries is None:
        pred = all_gather_entries(aff_list.float(), group)
        gt = all_gather_entries(aff_gt_list.float().detach(), group)
        loss = affinity_loss(pred, gt)
    else:
        n = aff_list.numel()                       # a shape: known on the host without touching the device
        if n > max_entries:
            raise ValueError(f"sharded_affinity_loss: {n} entries on this rank exceed max_entries={max_entries}")
        block = torch.zeros(3, max_entries, device=aff_list.device, dtype=torch.float32)
        block[0] = 0.5                             # padding predictions: any value inside (0,1); masked out below
        if n:
            block = torch.cat([torch.cat([aff_list.float().reshape(-1), block[0, n:]])[None],
                               torch.cat([aff_gt_list.float().detach().reshape(-1), block[1, n:]])[None],
                               torch.cat([torch.ones(n, device=block.device), block[2, n:]])[None]])
        allb = _AllGatherFixed.apply(block, group)                     # (world, 3, max_entries)
        pred, gt, mask = allb[:, 0], allb[:, 1].detach(), allb[:, 2].detach()
        bce = F.binary_cross_entropy(pred, gt, reduction="none")
        cnt = mask.sum()
        loss = torch.where(cnt > 0, (bce * mask).sum() / cnt.clamp_min(1.0), torch.zeros((), device=block.device))
    if grad_average:
        loss = loss.detach() + (loss - loss.detach()) * world
    return loss


def track_4d_loss(pc1_wrap, cls, gt_flow, gt_cls, aff_list=None, aff_gt_list=None, pretrain=False, group=None,
                  flow_reduction="mean", max_entries=None):
    """total = 0.5 * flow + 0.5 * affinity + 1.0 * segmentation (segmentation only while `pretrain`); NaN terms are
    replaced by 0 as in the reference.  reference :8-31.  Returns (total, items)."""
    sf = flow_loss(pc1_wrap, gt_flow, reduction=flow_reduction)
    if aff_list is None:
        trk = torch.zeros((), device=pc1_wrap.device)
    else:
        trk = sharded_affinity_loss(aff_list, aff_gt_list, group, max_entries=max_entries)
    seg = motion_seg_loss(cls, gt_cls, nan_to_zero=True)
    zero = torch.zeros((), device=pc1_wrap.device)
    sf = torch.where(torch.isnan(sf), zero, sf)
    trk = torch.where(torch.isnan(trk), zero, trk)
    total = 0.5 * sf + 0.5 * trk + 1.0 * seg
    if pretrain:
        total = seg
    return total, {"Loss": total, "SceneFlowLoss": sf, "TrackingLoss": trk, "SegLoss": seg}
