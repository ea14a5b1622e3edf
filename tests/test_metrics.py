"""Evaluation metrics: the numpy oracle and the device implementation against golden outputs of
the reference's own eval_scene_flow / eval_motion_seg (tests/golden/metrics.npz, generator oracle/gen_golden_metrics.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import metrics_oracle
from oracle.gen_golden_metrics import SEG_KEYS, SF_KEYS, make_case

G = np.load(os.path.join(GOLDEN, "metrics.npz"))
SF_CASES = [tuple(int(x) for x in c) for c in G["sf_cases"]]


def _close(a, b, tol):
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= tol * max(1.0, abs(b))


@pytest.mark.parametrize("n,seed,hard", SF_CASES)
def test_oracle_matches_reference_scene_flow_metrics(n, seed, hard):
    pc, pred, gt, mask = make_case(n, seed, bool(hard))
    out = metrics_oracle.eval_scene_flow(pc, pred, gt, mask)
    for k, want in zip(SF_KEYS, G[f"sf_{n}_{seed}"]):
        assert _close(float(out[k]), float(want), 1e-12), (k, float(out[k]), float(want))


def test_oracle_matches_reference_motion_seg_metrics():
    for i in range(3):
        out = metrics_oracle.eval_motion_seg(G[f"seg_pre_{i}"].astype(np.float32), G[f"seg_gt_{i}"].astype(np.float32))
        for k, want in zip(SEG_KEYS, G[f"seg_{i}"]):
            assert _close(float(out[k]), float(want), 1e-12), (i, k)


def _device_checks(dev):
    from ratrack_b200 import metrics

    for n, seed, hard in SF_CASES:
        pc, pred, gt, mask = (torch.from_numpy(x).to(dev) for x in make_case(n, seed, bool(hard)))
        out = metrics.as_floats(metrics.eval_scene_flow(pc, pred, gt, mask))
        for k, want in zip(SF_KEYS, G[f"sf_{n}_{seed}"]):
            # float32 sin/cos/sqrt of the device's libm differ from numpy's in the last bit: 1e-6 relative
            assert _close(out[k], float(want), 1e-6), (n, k, out[k], float(want))
    for i in range(3):
        pre, gt = (torch.from_numpy(G[f"seg_{w}_{i}"].astype(np.float32)).to(dev) for w in ("pre", "gt"))
        out = metrics.as_floats(metrics.eval_motion_seg(pre, gt))
        for k, want in zip(SEG_KEYS, G[f"seg_{i}"]):
            assert _close(out[k], float(want), 1e-12), (i, k)
    # a batch of frames = the mean of the per-frame values for the frame-averaged metrics
    cases = [make_case(512, 10 + j, True) for j in range(3)]
    b = [torch.from_numpy(np.concatenate([c[i] for c in cases])).to(dev) for i in range(4)]
    whole = metrics.as_floats(metrics.eval_scene_flow(*b))
    parts = [metrics.as_floats(metrics.eval_scene_flow(*[torch.from_numpy(x).to(dev) for x in c])) for c in cases]
    for k in ("rne", "epe", "sas", "ras"):
        assert abs(whole[k] - np.mean([p[k] for p in parts])) <= 1e-9
    tot = {}
    for p in cases:
        metrics.accumulate(tot, metrics.eval_scene_flow(*[torch.from_numpy(x).to(dev) for x in p]))
    assert abs(metrics.as_floats(tot)["epe"] - sum(p["epe"] for p in parts)) <= 1e-9


def test_metrics_module_on_cpu_tensors():
    """The metric functions are plain tensor programs: the same code is checked here on CPU tensors (host logic) ..."""
    _device_checks(torch.device("cpu"))


@pytest.mark.gpu
def test_metrics_module_on_the_device():
    """... and on the device, where the epoch loop calls them without a host round trip per frame."""
    _device_checks(torch.device("cuda"))
