"""Generate tests/golden/metrics.npz from the UNMODIFIED reference (dev container only): `python -m oracle.gen_golden_metrics`.
TEST INFRASTRUCTURE ONLY.  Stores what the reference's own eval_scene_flow / eval_motion_seg (src/main_utils.py:342-389)
return for seeded radar-shaped inputs: a frame whose mask is the raw sigmoid output (what main_utils.py:146 passes), a frame
with a hard 0/1 mask, and segmentation cases including an empty class."""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

SF_KEYS = ["rne", "50-50 rne", "mov_rne", "stat_rne", "sas", "ras", "epe"]
SEG_KEYS = ["acc", "miou", "sen"]


def make_case(n, seed, hard_mask):
    from ratrack_b200 import synthetic

    d = synthetic.make_batch(1, n, seed=seed)
    rng = np.random.default_rng(seed)
    pc = d["pc1"].astype(np.float32)
    if seed == 4:
        pc = pc * np.float32(0.02)        # short target vectors: the relative-error branches of sas / ras decide
    gt = pc + rng.normal(0, 0.3, pc.shape).astype(np.float32)
    pred = gt + rng.normal(0, 0.15, pc.shape).astype(np.float32)
    mask = (rng.random((1, n)) < 0.3).astype(np.float32) if hard_mask else (1 / (1 + np.exp(-rng.normal(0, 2, (1, n))))).astype(np.float32)
    return pc, pred, gt, mask


if __name__ == "__main__":
    from oracle import ref_harness

    ref_harness.install()
    import main_utils as M

    save = {}
    cases = [(256, 1, False), (1024, 2, True), (333, 3, True), (512, 4, True)]
    save["sf_cases"] = np.array([(n, s, int(h)) for n, s, h in cases])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for n, s, h in cases:
            pc, pred, gt, mask = make_case(n, s, h)
            out = M.eval_scene_flow(torch.from_numpy(pc), torch.from_numpy(pred), torch.from_numpy(gt), torch.from_numpy(mask))
            save[f"sf_{n}_{s}"] = np.array([float(out[k]) for k in SF_KEYS], dtype=np.float64)
            print(n, s, {k: float(out[k]) for k in SF_KEYS})
        rng = np.random.default_rng(5)
        seg = [(rng.random(300) < 0.4, rng.random(300) < 0.3), (np.zeros(64, bool), rng.random(64) < 0.5), (np.ones(10, bool), np.ones(10, bool))]
        for i, (pre, gt) in enumerate(seg):
            out = M.eval_motion_seg(torch.from_numpy(pre.astype(np.float32)), torch.from_numpy(gt.astype(np.float32)))
            save[f"seg_pre_{i}"], save[f"seg_gt_{i}"] = pre, gt
            save[f"seg_{i}"] = np.array([float(out[k]) for k in SEG_KEYS], dtype=np.float64)
            print(i, {k: float(out[k]) for k in SEG_KEYS})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **save)
