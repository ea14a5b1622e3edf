"""One training step of the hot path, as the reference's loop runs it (reference: src/main_utils.py:127-157 forward +
loss, src/main.py:61 Adam(lr, weight_decay=1e-10), src/models/model.py:38-40 multi-GPU replication), with the batch
dimension the reference hard-codes to 1 generalised and nn.DataParallel replaced by one process per GPU:

    zero_grad -> Track4DBackbone (train mode: BatchNorm on batch statistics, CUDA grad kernels under autograd)
              -> pc1_warp = pc1 + flow -> track_4d_loss (0.5 flow + 0.5 affinity + 1.0 segmentation)
              -> backward -> ONE all_reduce of the flat gradient bucket (ranks > 1) -> Adam step

The affinity entries come from the association module, which is outside this package's path; callers pass them in
(`aff_fn`), and their loss is a global mean over all ranks' entries (losses.sharded_affinity_loss: all_gather).
"""
import torch

from . import losses, sharding


def make_optimizer(net, lr=1e-3):
    return torch.optim.Adam(net.parameters(), lr=lr, weight_decay=1e-10)   # reference: src/main.py:61


def train_step(net, opt, pc1, pc2, ft1, ft2, gt_flow, gt_cls, h=None, aff_fn=None, aff_gt=None, pretrain=False, group=None,
               buckets=None, max_entries=None):
    """-> (total loss (detached), items dict, h (detached)).  `aff_fn(outputs) -> aff_list` builds the affinity
    entries from the backbone outputs (None: no tracking term).  `buckets` (sharding.GradBuckets over `net`) overlaps the
    gradient all-reduce with the tail of backward; without it ONE all_reduce runs after backward.  `max_entries` selects
    the synchronisation-free affinity all_gather (losses.sharded_affinity_loss)."""
    net.train()
    opt.zero_grad(set_to_none=True)
    out = net.backbone(pc1, pc2, ft1, ft2, h)
    flow, h_new, cls = out[0], out[1], out[2]
    pc1_wrap = pc1 + flow                                                # reference: src/models/track4d.py:52
    aff = aff_fn(out) if aff_fn is not None else None
    total, items = losses.track_4d_loss(pc1_wrap, cls, gt_flow, gt_cls, aff, aff_gt, pretrain=pretrain, group=group,
                                        max_entries=max_entries)
    total.backward()
    if buckets is not None:
        buckets.finish()
    else:
        sharding.allreduce_gradients(net.parameters(), average=True, group=group)
    opt.step()
    return total.detach(), {k: v.detach() for k, v in items.items()}, h_new.detach()
