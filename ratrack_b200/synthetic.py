"""Deterministic synthetic radar frame pairs shaped like View-of-Delft radar records.

Recipe: SURVEY.md section 8(d).  Statistics follow the three example frames the reference
ships (src/dataset_classes/vod/example_set/radar/training/velodyne/*.bin: x in [0,98],
sigma_y ~ 9, sigma_z ~ 2.5, RCS ~ N(-13, 12.5), v_r ~ N(-2, 1.6), ~1 % duplicate xyz rows).
A record is the 7-column VoD row [x, y, z, RCS, v_r, v_r_comp, time]
(src/vod/frame/data_loader.py:71,174); the model consumes pc = cols 0:3 as (B,3,N) and
ft = cols 3:5 as (B,2,N) (src/main_utils.py:76-79).
"""
import os

import numpy as np


def make_frame_pair(n_points: int, rng: np.random.Generator):
    """One (frame1, frame2) pair of (N,7) float32 records."""
    n = int(n_points)
    n_clu = 24
    centres = np.stack([rng.uniform(2, 80, n_clu), rng.uniform(-25, 25, n_clu), rng.uniform(-1, 2, n_clu)], 1)
    vel = rng.uniform(-1, 1, (n_clu, 3)) * np.array([10.0, 10.0, 0.5]) / np.sqrt(3.0)
    n_fg = int(round(0.6 * n))
    cid = rng.integers(0, n_clu, n_fg)
    fg = centres[cid] + rng.normal(0, 1.0, (n_fg, 3))
    bg = np.stack([rng.uniform(0, 100, n - n_fg), rng.uniform(-40, 40, n - n_fg), rng.uniform(-3, 5, n - n_fg)], 1)
    xyz1 = np.concatenate([fg, bg], 0)
    cl = np.concatenate([cid, -np.ones(n - n_fg, dtype=np.int64)])
    perm = rng.permutation(n)
    xyz1, cl = xyz1[perm], cl[perm]
    # ~1 % exact duplicates (forces FPS / ball-query tie-breaks, as in the real frames)
    n_dup = int(np.ceil(0.01 * n))
    dst = rng.choice(n, n_dup, replace=False)
    src = rng.integers(0, n, n_dup)
    xyz1[dst] = xyz1[src]
    cl[dst] = cl[src]

    def attrs(m, clv):
        rcs = rng.normal(-13, 12.5, m)
        v = np.where(clv >= 0, vel[np.maximum(clv, 0), 0], 0.0)
        v_r = rng.normal(-2, 1.6, m) + 0.2 * v
        v_rc = v_r + rng.normal(0, 0.3, m)
        return np.stack([rcs, v_r, v_rc, np.zeros(m)], 1)

    f1 = np.concatenate([xyz1, attrs(n, cl)], 1)
    # second frame: ego shift + per-cluster motion over 0.1 s + jitter, 10 % rows redrawn, shuffled
    motion = np.where(cl[:, None] >= 0, vel[np.maximum(cl, 0)] * 0.1, 0.0)
    xyz2 = xyz1 + np.array([-0.8, 0.0, 0.0]) + motion + rng.normal(0, 0.05, (n, 3))
    n_new = int(round(0.1 * n))
    redo = rng.choice(n, n_new, replace=False)
    xyz2[redo] = np.stack([rng.uniform(0, 100, n_new), rng.uniform(-40, 40, n_new), rng.uniform(-3, 5, n_new)], 1)
    cl2 = cl.copy()
    cl2[redo] = -1
    f2 = np.concatenate([xyz2, attrs(n, cl2)], 1)
    f2 = f2[rng.permutation(n)]
    return f1.astype(np.float32), f2.astype(np.float32)


def make_batch(batch: int, n_points: int, seed: int = 1234):
    """-> dict of numpy arrays: pc1, pc2 (B,3,N), ft1, ft2 (B,2,N), float32, C-contiguous."""
    rng = np.random.default_rng(seed)
    pc1 = np.empty((batch, 3, n_points), np.float32)
    pc2 = np.empty_like(pc1)
    ft1 = np.empty((batch, 2, n_points), np.float32)
    ft2 = np.empty_like(ft1)
    for b in range(batch):
        a, c = make_frame_pair(n_points, rng)
        pc1[b], ft1[b] = a[:, 0:3].T, a[:, 3:5].T
        pc2[b], ft2[b] = c[:, 0:3].T, c[:, 3:5].T
    return dict(pc1=pc1, pc2=pc2, ft1=ft1, ft2=ft2)


BN_CALIB_FILE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "bn_calibration_seed1234.npz")


def make_state_dict(module, seed: int = 1234, calibrated: bool = True):
    """Deterministic weights for any module exposing the reference's state_dict surface.

    There is no network for checkpoints, so benchmarks and parity fixtures use this recipe: every
    tensor is drawn from a numpy Generator seeded by (seed, crc32(key)) -- independent of module
    construction order, so the reference model and this package get identical values for
    identical keys.  Conv / Linear weights ~ N(0, 2/fan_in) (Kaiming-normal, as the reference
    initialises its SharedMLP convs: src/lib/pytorch_utils.py:138,176), GRU weights ~ N(0, 1/fan_in),
    biases ~ N(0, 0.05), BN weight ~ U(0.8, 1.2), BN bias ~ N(0, 0.1).  The last WeightNet conv
    (8 -> 256, no normalisation follows it) is scaled by 1/16 so the two sums over 16 neighbours in
    the cost volume (src/utils/model_utils/model_utils.py:236,248) keep activations O(1), as a trained
    network's would be.

    BN running statistics: with `calibrated` (seed 1234 only) they come from
    data/bn_calibration_seed1234.npz -- batch statistics of one train-mode pass (momentum 1) of the
    reference model over synthetic.make_batch(4, 1024, seed=4321), written by oracle/gen_golden.py --
    so eval-mode activations are normalised like a trained model's.  Otherwise
    running_mean ~ N(0, 0.2), running_var ~ U(0.6, 1.6).
    Returns a dict of torch tensors (CPU, fp32 / int64) to pass to load_state_dict(strict=False).
    """
    import os
    import zlib

    import torch

    calib = None
    if calibrated:
        if seed != 1234 or not os.path.exists(BN_CALIB_FILE):
            raise RuntimeError("BN calibration exists only for seed 1234 (ratrack_b200/data/); pass calibrated=False")
        calib = np.load(BN_CALIB_FILE)
    out = {}
    for key, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        rng = np.random.default_rng([seed, zlib.crc32(key.encode())])
        parts = key.split(".")
        is_bn = ".bn." in key or "mlp_bns" in key or (len(parts) >= 2 and parts[-2] == "1")
        if key.endswith("num_batches_tracked"):
            out[key] = torch.tensor(1, dtype=torch.int64)
            continue
        if key.endswith("running_mean") or key.endswith("running_var"):
            if calib is not None and key in calib.files:
                v = calib[key]
            elif calib is not None:
                v = np.zeros(shape) if key.endswith("running_mean") else np.ones(shape)
            elif key.endswith("running_mean"):
                v = rng.normal(0, 0.2, shape)
            else:
                v = rng.uniform(0.6, 1.6, shape)
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            gain = 1.0 if "GRU" in key or "gru" in key else 2.0
            v = rng.normal(0, np.sqrt(gain / fan_in), shape)
            if "weightnet" in key and key.endswith("mlp_convs.2.weight"):
                v = v / 16.0
        elif len(shape) == 1 and is_bn and key.endswith("weight"):
            v = rng.uniform(0.8, 1.2, shape)
        elif len(shape) == 1 and is_bn:
            v = rng.normal(0, 0.1, shape)
        elif len(shape) == 0:
            v = np.asarray(1.0)
        else:
            v = rng.normal(0, 0.05, shape)
            if "weightnet" in key and key.endswith("mlp_convs.2.bias"):
                v = v / 16.0
        out[key] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return out
