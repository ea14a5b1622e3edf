"""numpy restatement of the reference's evaluation metrics.  TEST INFRASTRUCTURE ONLY.

eval_scene_flow: /root/reference/src/main_utils.py:342-374; eval_motion_seg: :377-389; get_carterian_res: :260-298.
Pinned by tests/golden/metrics.npz, which oracle/gen_golden_metrics.py produced by calling the reference's own functions."""
import numpy as np


def get_carterian_res(pc, sensor):
    res = np.array([0.2, 1 * np.pi / 180, 1.6 * np.pi / 180] if sensor == "radar" else [0.04, 0.4 * np.pi / 180, 0.08 * np.pi / 180])
    x, y, z = pc[:, 0], pc[:, 1], pc[:, 2]
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    theta = np.arcsin(z / r)
    phi = np.arctan2(y, x)
    gx = np.stack((np.cos(phi) * np.cos(theta), -r * np.sin(theta) * np.cos(phi), -r * np.cos(theta) * np.sin(phi)), axis=2)
    gy = np.stack((np.sin(phi) * np.cos(theta), -r * np.sin(phi) * np.sin(theta), r * np.cos(theta) * np.cos(phi)), axis=2)
    gz = np.stack((np.sin(theta), r * np.cos(theta), np.zeros((np.size(x, 0), np.size(x, 1)))), axis=2)
    return np.stack((np.sum(abs(gx) * res, axis=2), np.sum(abs(gy) * res, axis=2), np.sum(abs(gz) * res, axis=2)), axis=2)


def eval_scene_flow(pc, pred, labels, mask):
    """batch-1 arrays as the reference takes them: pc, pred, labels (1,3,N); mask (1,N)."""
    mask = mask[0]
    error = np.sqrt(np.sum((pred - labels) ** 2, 1) + 1e-20)
    epe = np.mean(error)
    gtflow_len = np.sqrt(np.sum(labels * labels, 1) + 1e-20)
    res_r = np.sqrt(np.sum(get_carterian_res(pc, "radar"), 2) + 1e-20)
    res_l = np.sqrt(np.sum(get_carterian_res(pc, "lidar"), 2) + 1e-20)
    rn_error = error / (res_r / res_l)
    rne = np.mean(rn_error)
    mov_rne = np.sum(rn_error[:, mask == 0]) / (np.sum(mask == 0) + 1e-6)
    with np.errstate(all="ignore"):
        stat_rne = np.mean(rn_error[:, mask == 1])
    avg_rne = (mov_rne + stat_rne) / 2
    n = np.size(pred, 0) * np.size(pred, 2)
    sas = np.sum(np.logical_or(rn_error <= 0.10, rn_error / gtflow_len <= 0.10)) / n
    ras = np.sum(np.logical_or(rn_error <= 0.20, rn_error / gtflow_len <= 0.20)) / n
    return {"rne": rne, "50-50 rne": avg_rne, "mov_rne": mov_rne, "stat_rne": stat_rne, "sas": sas, "ras": ras, "epe": epe}


def eval_motion_seg(pre, gt):
    tp = np.logical_and(pre == 1, gt == 1).sum() + 1e-20
    tn = np.logical_and(pre == 0, gt == 0).sum() + 1e-20
    fp = np.logical_and(pre == 1, gt == 0).sum() + 1e-20
    fn = np.logical_and(pre == 0, gt == 1).sum() + 1e-20
    return {"acc": (tp + tn) / (tp + tn + fp + fn), "miou": 0.5 * (tp / (tp + fp + fn + 1e-4) + tn / (tn + fp + fn + 1e-4)),
            "sen": tp / (tp + fn)}
