"""Evaluation metrics of the reference's epoch loop on the device, with the reference's names, arguments and arithmetic
(reference: src/main_utils.py:342-374 eval_scene_flow, :377-389 eval_motion_seg, :260-298 get_carterian_res).

The reference copies every tensor to the host and evaluates with numpy once per frame (four device->host copies and a
synchronisation per frame inside the training / evaluation loop).  Here every metric is a 0-dim DEVICE tensor: nothing is
read back until the caller asks (`as_floats`), so the per-frame loop stays asynchronous, and a batch of frames is handled
in one call (the reference is written for batch 1: `mask.squeeze(0)`).  numpy promotes the resolution terms to float64
(`res = np.array([...])` is float64); the same promotion is made here so the values agree to float64 round-off.
"""
import math

import torch

_RES = {"radar": (0.2, 1 * math.pi / 180, 1.6 * math.pi / 180),      # LRR30    (reference :262-265)
        "lidar": (0.04, 0.4 * math.pi / 180, 0.08 * math.pi / 180)}  # HDL-64E  (reference :267-270)


def get_carterian_res(pc, sensor):
    """pc (B,3,N) -> (B,N,3) float64 per-point x/y/z measurement resolution of a range/elevation/azimuth sensor.
    reference :260-298 (name kept, typo included: callers import it by this name)."""
    res = torch.tensor(_RES[sensor], dtype=torch.float64, device=pc.device)
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]                      # float32, as the reference computes r / theta / phi
    r = torch.sqrt(x ** 2 + y ** 2 + z ** 2)
    theta = torch.arcsin(z / r)
    phi = torch.arctan2(y, x)
    ct, st, cp, sp = torch.cos(theta), torch.sin(theta), torch.cos(phi), torch.sin(phi)
    grad_x = torch.stack((cp * ct, -r * st * cp, -r * ct * sp), dim=2)
    grad_y = torch.stack((sp * ct, -r * sp * st, r * ct * cp), dim=2)
    grad_z = torch.stack((st, r * ct, torch.zeros_like(x)), dim=2)
    return torch.stack([(g.abs().double() * res).sum(dim=2) for g in (grad_x, grad_y, grad_z)], dim=2)


def eval_scene_flow(pc, pred, labels, mask):
    """pc, pred, labels (B,3,N); mask (B,N) (the reference passes the raw `cls` probabilities, :146, and compares them
    with 0 and 1 exactly -- kept as is).  -> dict of 0-dim device tensors: rne, 50-50 rne, mov_rne, stat_rne, sas, ras, epe.
    reference :342-374."""
    mask = mask.reshape(mask.shape[-2], mask.shape[-1]) if mask.dim() > 2 else mask
    if mask.dim() == 1:
        mask = mask.unsqueeze(0)
    pc, pred, labels = pc.detach().float(), pred.detach().float(), labels.detach().float()
    error = torch.sqrt(((pred - labels) ** 2).sum(dim=1) + 1e-20)                 # (B,N) float32
    epe = error.double().mean()
    gtflow_len = torch.sqrt((labels * labels).sum(dim=1) + 1e-20)
    res_r = torch.sqrt(get_carterian_res(pc, "radar").sum(dim=2) + 1e-20)
    res_l = torch.sqrt(get_carterian_res(pc, "lidar").sum(dim=2) + 1e-20)
    rn_error = error.double() / (res_r / res_l)
    rne = rn_error.mean()
    m0 = (mask.detach() == 0)
    m1 = (mask.detach() == 1)
    # the reference indexes columns with the batch-1 mask (rn_error[:, mask == 0]); per-frame masks generalise it
    mov_rne = (rn_error * m0).sum() / (m0.sum() + 1e-6)
    stat_rne = (rn_error * m1).sum() / m1.sum()                                   # 0/0 = nan when no entry equals 1: np.mean([])
    avg_rne = (mov_rne + stat_rne) / 2
    total = pred.shape[0] * pred.shape[2]
    rel = rn_error / gtflow_len.double()
    sas = ((rn_error <= 0.10) | (rel <= 0.10)).sum().double() / total
    ras = ((rn_error <= 0.20) | (rel <= 0.20)).sum().double() / total
    return {"rne": rne, "50-50 rne": avg_rne, "mov_rne": mov_rne, "stat_rne": stat_rne, "sas": sas, "ras": ras, "epe": epe}


def eval_motion_seg(pre, gt):
    """pre, gt: 0/1 tensors of one shape -> dict of 0-dim device tensors acc, miou, sen.  reference :377-389."""
    pre, gt = pre.detach(), gt.detach()
    tp = ((pre == 1) & (gt == 1)).sum().double() + 1e-20
    tn = ((pre == 0) & (gt == 0)).sum().double() + 1e-20
    fp = ((pre == 1) & (gt == 0)).sum().double() + 1e-20
    fn = ((pre == 0) & (gt == 1)).sum().double() + 1e-20
    acc = (tp + tn) / (tp + tn + fp + fn)
    sen = tp / (tp + fn)
    miou = 0.5 * (tp / (tp + fp + fn + 1e-4) + tn / (tn + fp + fn + 1e-4))
    return {"acc": acc, "miou": miou, "sen": sen}


def accumulate(total, frame):
    """total[key] += frame[key] on the device (reference :141-150 does this with host floats)."""
    for k, v in frame.items():
        total[k] = v if k not in total else total[k] + v
    return total


def as_floats(metrics):
    """One device->host read for a whole dict of metrics."""
    keys = list(metrics)
    if not keys:
        return {}
    vals = torch.stack([metrics[k].double() for k in keys]).cpu().tolist()
    return dict(zip(keys, vals))
